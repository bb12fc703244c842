cd /root/repo
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ffma_engine" > gpurun_out/t1.log 2>&1; tail -15 gpurun_out/t1.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v3.log 2>&1; tail -15 gpurun_out/pytest_v3.log
timeout 300 python bench.py --no-extra --steps 50 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; tail -c 1500 gpurun_out/bench_v3.json; tail -5 gpurun_out/bench_v3.err
