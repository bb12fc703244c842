cd /root/repo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_check.py > gpurun_out/multigpu_check_h_n2.log 2>&1; echo "check rc=$?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/multigpu_check_h_n2.log | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_h_n2.json 2> gpurun_out/bench_h_n2.err; echo "bench n=2 rc=$?"; head -c 300 gpurun_out/bench_h_n2.json; echo
