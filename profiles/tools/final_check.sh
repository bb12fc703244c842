#!/bin/bash
# One-box end-of-round check (run under gpurun, 1 GPU): GPU test suite, smoke(), the default bench line, the launch list of the
# same command under ncu, and the reference arm.  Outputs under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-.}"
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/final_pytest.log; cat gpurun_out/final_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_final_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["drop_in_api"]["ms_per_step"], d["parity"]["ok"])
for r in d["extra"]["global_attention_general"]:
    print(r["res"], round(r["ms_fwd_bwd"] * 1e3, 1), round(r["hbm_frac_fwd_bwd"], 3), round(r["ms_fwd"] * 1e3, 1), round(r["hbm_frac_fwd"], 3))
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --no-extra --steps 3 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
wc -l gpurun_out/r2_final_launches.csv
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2_final_reference_arm_bench.json
cut -c1-300 gpurun_out/r2_final_reference_arm_bench.json
