cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gag" 2>&1 | tail -6
timeout 300 python scratch/gag_time.py 2>&1 | tail -4
