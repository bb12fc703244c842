cd /root/repo
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gag" 2>&1 | tail -2
EEGAN_GAG_RPT=4 EEGAN_GAG_NS=3 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gag" 2>&1 | tail -2
echo "RPT=8 NS=4 (default)"; timeout 120 python scratch/gag_time.py 2>&1 | tail -3
echo "RPT=8 NS=3"; EEGAN_GAG_NS=3 timeout 120 python scratch/gag_time.py 2>&1 | tail -3
echo "RPT=4 NS=4"; EEGAN_GAG_RPT=4 timeout 120 python scratch/gag_time.py 2>&1 | tail -3
echo "RPT=4 NS=3"; EEGAN_GAG_RPT=4 EEGAN_GAG_NS=3 timeout 120 python scratch/gag_time.py 2>&1 | tail -3
