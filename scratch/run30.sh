cd /root/repo
export EEGAN_ENGINE=3
for dbg in 0 1 2 4 6; do
EEGAN_H_DBG=$dbg timeout 300 python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err; echo "dbg=$dbg rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_dbg.json'))
s=d['roofline']['stage_ms_per_step']
print(round(d['ms_per_step']*1e3,1), {k[:5]:round(v*1e3,1) for k,v in s.items() if v>0})
PY
done
