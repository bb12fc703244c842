cd /root/repo
export EEGAN_ENGINE=3
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --no-extra --steps 30 --warmup 5 > gpurun_out/bench_h_v1.json 2> gpurun_out/bench_h_v1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_h_v1.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print(d['roofline']['stage_ms_per_step'])
PY
