cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s)
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "tests done $(( $(date +%s) - T0 )) s"
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$? $(( $(date +%s) - T0 )) s"
for c in 3 4 9 12; do echo "CPS=$c"; EEGAN_GAG_CPS=$c timeout 120 python scratch/gag_time.py 2>&1 | tail -3; done
echo "CPS=6 (default)"; timeout 120 python scratch/gag_time.py 2>&1 | tail -3
echo "total $(( $(date +%s) - T0 )) s"
