cd /root/repo
T0=$(date +%s)
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gag" 2>&1 | tail -3
EEGAN_GAG_RPT=4 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gag" 2>&1 | tail -3
echo "RPT=8 (default)"; timeout 120 python scratch/gag_time.py 2>&1 | tail -3
echo "RPT=4"; EEGAN_GAG_RPT=4 timeout 120 python scratch/gag_time.py 2>&1 | tail -3
echo "old rowsum"; EEGAN_GAG_ROWSUM=0 timeout 120 python scratch/gag_time.py 2>&1 | tail -3
echo "total $(( $(date +%s) - T0 )) s"
