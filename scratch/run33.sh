cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "graphed or fixture" 2>&1 | tail -3
timeout 300 python bench.py --no-extra --steps 50 --warmup 10 > gpurun_out/bench_h_v3.json 2> gpurun_out/bench_h_v3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_h_v3.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
PY
tail -2 gpurun_out/bench_h_v3.err
