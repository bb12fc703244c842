set -x
cd /root/repo
timeout 120 python scratch/trunc_test.py > gpurun_out/trunc.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1_v2.log 2>&1; tail -3 gpurun_out/pytest_r1_v2.log
timeout 600 python bench.py > gpurun_out/bench_r1_v2.json 2> gpurun_out/bench_r1_v2.err; tail -c 600 gpurun_out/bench_r1_v2.err
cat gpurun_out/trunc.log
