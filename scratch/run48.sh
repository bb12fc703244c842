cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_final.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['frac_of_engine_ceiling'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --no-extra --steps 3 --warmup 3 --eager > gpurun_out/ncu_final_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:h_gemm -s 15 -c 5 -o gpurun_out/prof_r1_final_gemm -f python bench.py --no-extra --steps 2 --warmup 3 --eager > gpurun_out/ncu_final_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gag_bwd -c 4 -o gpurun_out/prof_r1_final_gag -f python scratch/gag_one.py > gpurun_out/ncu_final_gag.log 2>&1; echo "ncu gag rc=$?"
python __graft_entry__.py smoke 2>&1 | tail -1
