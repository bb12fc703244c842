cd /root/repo
for f in 1 0; do
EEGAN_CARVEOUT=$f timeout 300 python bench.py --no-extra --steps 50 --warmup 10 > gpurun_out/bench_h_carve$f.json 2>/dev/null; echo "carveout=$f rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_h_carve$f.json'))
s=d['roofline']['stage_ms_per_step']
print(round(d['ms_per_step']*1e3,1), d['value'], {k[:5]:round(v*1e3,1) for k,v in s.items() if v>0})
PY
done
