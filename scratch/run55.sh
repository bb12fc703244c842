cd /root/repo
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gag" 2>&1 | tail -2
timeout 60 python scratch/gag_time.py 2>&1 | tail -3
