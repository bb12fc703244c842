cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v3.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_v3.log
timeout 300 python bench.py --no-extra --steps 50 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; tail -3 gpurun_out/bench_v3.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_v3.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print({k: round(v*1e3,1) for k,v in d['roofline']['stage_ms_per_step'].items()})
PY
