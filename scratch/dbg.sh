for d in 0 1 2 4 3 5 6 7; do
  EEGAN_TC_DBG=$d EEGAN_TC_TRUNC_HI=1 python bench.py --steps 30 --warmup 5 --no-extra 2>/dev/null | tail -1 > /tmp/o.json
  python - <<PY
import json
d=json.load(open('/tmp/o.json')); s=d["roofline"]["stage_ms_per_step"]
print("dbg=$d", [round(1e3*s[k]) for k in s if k.startswith("gemm")], round(1e3*d["ms_per_step"]))
PY
done
