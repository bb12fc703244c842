cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/bench_r1_h.json 2> gpurun_out/bench_r1_h.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1_h.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac_of_engine_ceiling'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_h.csv python bench.py --no-extra --steps 3 --warmup 3 --eager > gpurun_out/ncu_h_list.log 2>&1; echo "ncu list rc=$?"
