cd /root/repo
timeout 300 python bench.py --no-extra --steps 2 --warmup 3 --eager > gpurun_out/plain_v3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 15 -c 5 -o gpurun_out/prof_r1_v3b_gemm -f python bench.py --no-extra --steps 2 --warmup 3 --eager > gpurun_out/ncu_v3.log 2>&1
tail -2 gpurun_out/ncu_v3.log
