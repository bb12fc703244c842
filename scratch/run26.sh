cd /root/repo
timeout 300 python -m pytest tests/test_gpu_gemm_h.py -x -q 2>&1 | tail -15
timeout 120 python scratch/h_probe.py 2>&1 | tail -5
