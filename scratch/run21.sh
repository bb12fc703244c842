cd /root/repo
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gag or attention" > gpurun_out/t_gag.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_gag.log | cut -c1-250
timeout 100 python scratch/gag_time.py
EEGAN_GAG_TC=0 timeout 100 python scratch/gag_time.py
