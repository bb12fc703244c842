cd /root/repo
timeout 300 python bench.py --no-extra --steps 3 --warmup 3 --eager > gpurun_out/plain_v3.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v3.csv python bench.py --no-extra --steps 3 --warmup 3 --eager > gpurun_out/ncu_v3_list.log 2>&1
timeout 300 python bench.py --no-extra --steps 2 --warmup 3 --eager > gpurun_out/plain_v3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|v3_" -s 42 -c 14 -o gpurun_out/prof_r1_v3_step -f python bench.py --no-extra --steps 2 --warmup 3 --eager > gpurun_out/ncu_v3_full.log 2>&1
tail -2 gpurun_out/ncu_v3_full.log | cut -c1-200
