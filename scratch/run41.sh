cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_r1_h3.json 2> gpurun_out/bench_r1_h3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_h3.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac_of_engine_ceiling'])
print(json.dumps(d['extra']['aux_rows_8f']))
for g in d['extra']['global_attention_general']: print(g['res'], g['ms_fwd_bwd'], g['hbm_frac_fwd_bwd'], g['ms_fwd'], g['hbm_frac_fwd'])
PY
