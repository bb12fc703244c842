cd /root/repo
export EEGAN_ENGINE=3
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for pdl in 1 0; do
EEGAN_PDL=$pdl timeout 300 python bench.py --no-extra --steps 30 --warmup 5 > gpurun_out/bench_h_v2_pdl$pdl.json 2> gpurun_out/bench_h_v2_pdl$pdl.err; echo "pdl=$pdl rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_h_v2_pdl$pdl.json'))
s=d['roofline']['stage_ms_per_step']
print(round(d['ms_per_step']*1e3,1), d['value'], d['e2e']['value'], {k[:5]:round(v*1e3,1) for k,v in s.items() if v>0})
PY
done
tail -3 gpurun_out/bench_h_v2_pdl1.err
