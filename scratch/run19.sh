cd /root/repo
for v in 0 1; do export EEGAN_V3_PRESPLIT_C=$v;
timeout 300 python bench.py --no-extra --steps 50 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; tail -2 gpurun_out/bench_v3.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_v3.json').read().strip().splitlines()[-1])
print("presplit_c=$v", d['value'], d['ms_per_step'])
print({k: round(v*1e3,1) for k,v in d['roofline']['stage_ms_per_step'].items() if v})
PY
done
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -2
