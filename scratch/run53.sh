cd /root/repo
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
echo "ffma2"; timeout 120 python scratch/gag_time.py 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_r1_ffma2.json 2> gpurun_out/bench_r1_ffma2.err; echo "bench rc=$?"
