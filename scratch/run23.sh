cd /root/repo
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gag" > gpurun_out/t_gag.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_gag.log | cut -c1-300
