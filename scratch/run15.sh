cd /root/repo
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gag or attention" > gpurun_out/t_gag.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_gag.log
timeout 200 python scratch/gag_time.py
