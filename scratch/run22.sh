cd /root/repo
timeout 120 python scratch/gag_one.py > gpurun_out/gag_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gag_tc -s 1 -c 1 -o gpurun_out/prof_r1_gag_tc -f python scratch/gag_one.py > gpurun_out/ncu_gag.log 2>&1
tail -2 gpurun_out/ncu_gag.log
