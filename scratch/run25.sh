cd /root/repo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_check.py > gpurun_out/multigpu_check_n8.log 2>&1; echo "check rc=$?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/multigpu_check_n8.log | tail -4
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/bench_v3_n$n.json 2> gpurun_out/bench_v3_n$n.err; echo "bench n=$n rc=$?"; tail -c 700 gpurun_out/bench_v3_n$n.json | head -c 400; echo
done
