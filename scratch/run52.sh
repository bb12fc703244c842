cd /root/repo
T0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; echo "bench rc=$? $(( $(date +%s) - T0 )) s"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_final_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_final_launches.csv python bench.py --no-extra --steps 3 --warmup 3 --eager > /dev/null 2>&1; echo "ncu rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
echo "total $(( $(date +%s) - T0 )) s"
