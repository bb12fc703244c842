set -x
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $1 "${@:2}" 2>> gpurun_out/r2n.err | grep "^{" ; }
for m in serial overlap graph; do run 8 --steps 30 --warmup 5 --sharded-mode $m > gpurun_out/r2n_c2_n8_$m.json; done
run 8 --config c3 --steps 30 --warmup 5 > gpurun_out/r2n_c3_n8.json
run 8 --config c5 --steps 10 --warmup 3 > gpurun_out/r2n_c5_n8.json
run 8 --config c4 > gpurun_out/r2n_c4_n8.json
run 4 --steps 30 --warmup 5 > gpurun_out/r2n_c2_n4_serial.json
run 4 --steps 30 --warmup 5 --sharded-mode overlap > gpurun_out/r2n_c2_n4_overlap.json
run 4 --config c3 --steps 30 --warmup 5 > gpurun_out/r2n_c3_n4.json
run 2 --config c3 --steps 30 --warmup 5 > gpurun_out/r2n_c3_n2.json
run 2 --config c4 > gpurun_out/r2n_c4_n2.json
python -m pytest tests/test_gpu_multigpu.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r2n_multigpu_pytest_n8.log
python - <<EOF2
import json,glob
for f in sorted(glob.glob("gpurun_out/r2n_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "%.3g"%d["value"], "%.4f ms"%d["ms_per_step"], d["parity"]["ok"], "e2e %.3g"%d["e2e"]["value"], (d.get("comm_free_same_shape") or {}).get("ms_per_step"), d.get("nccl_ms_per_step"))
        for r in d.get("sweep",[]): print("    B=%d %.3f ms %.3g pairs/s cf=%s nccl=%s"%(r["global_B"], r["ms_per_step"], r["pairs_per_s"], r["comm_free_ms"], r["nccl_ms"]))
    except Exception as e: print(f, "ERR", e)
EOF2
cat gpurun_out/r2n_multigpu_pytest_n8.log; tail -5 gpurun_out/r2n.err
