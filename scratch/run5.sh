cd /root/repo
timeout 120 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -x -q > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"; tail -8 gpurun_out/t_gemm.log
for ts in 0 1; do
EEGAN_V3_TS=$ts timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q > gpurun_out/pytest_v3_ts$ts.log 2>&1; echo "ts=$ts rc=$?"; tail -4 gpurun_out/pytest_v3_ts$ts.log
EEGAN_V3_TS=$ts timeout 300 python bench.py --no-extra --steps 50 > gpurun_out/bench_v3_ts$ts.json 2> gpurun_out/bench_v3.err; tail -3 gpurun_out/bench_v3.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_v3_ts$ts.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print({k: round(v*1e3,1) for k,v in d['roofline']['stage_ms_per_step'].items()})
PY
done
