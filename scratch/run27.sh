cd /root/repo
export EEGAN_ENGINE=3
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 300 python bench.py --no-extra --steps 30 --warmup 5 > gpurun_out/bench_h_v0.json 2> gpurun_out/bench_h_v0.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_h_v0.json
