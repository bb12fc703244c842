cd /root/repo
timeout 300 python -m pytest tests/test_gpu_ssa.py -m gpu -x -q > gpurun_out/t_ssa.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_ssa.log | cut -c1-250
