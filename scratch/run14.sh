cd /root/repo
timeout 900 python bench.py > gpurun_out/bench_r1_v3.json 2> gpurun_out/bench_r1_v3.err; echo "rc=$?"; tail -3 gpurun_out/bench_r1_v3.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; echo "ref rc=$?"; tail -c 700 gpurun_out/bench_r1_ref.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke()"
