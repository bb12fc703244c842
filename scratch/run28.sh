cd /root/repo
export EEGAN_ENGINE=3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:h_gemm -s 15 -c 5 -o gpurun_out/prof_r1_h_v0 -f python bench.py --no-extra --steps 2 --warmup 3 --eager > gpurun_out/ncu_h_v0.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_h_v0.log
