cd /root/repo
timeout 900 python bench.py > gpurun_out/bench_r1_h2.json 2> gpurun_out/bench_r1_h2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_h2.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac_of_engine_ceiling'])
print(json.dumps(d['extra']['aux_rows_8f']))
print(json.dumps(d['extra']['global_attention_general'][0]))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:h_gemm -s 15 -c 5 -o gpurun_out/prof_r1_h_v3 -f python bench.py --no-extra --steps 2 --warmup 3 --eager > gpurun_out/ncu_h_v3.log 2>&1; echo "ncu rc=$?"
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref2.json 2>/dev/null; echo "ref rc=$?"; head -c 400 gpurun_out/bench_r1_ref2.json
