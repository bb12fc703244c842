cd /root/repo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_check.py > gpurun_out/multigpu_check.log 2>&1; echo "check rc=$?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/multigpu_check.log | tail -5
