cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v3.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_v3.log
