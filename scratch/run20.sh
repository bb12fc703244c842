cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1_v3.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_r1_v3.log
timeout 900 python bench.py > gpurun_out/bench_r1_v3.json 2> gpurun_out/bench_r1_v3.err; echo "rc=$?"; tail -3 gpurun_out/bench_r1_v3.err
