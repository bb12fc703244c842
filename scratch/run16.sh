cd /root/repo
timeout 120 python scratch/gag_one.py > gpurun_out/gag_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gag_ -s 2 -c 2 -o gpurun_out/prof_r1_gag3 -f python scratch/gag_one.py > gpurun_out/ncu_gag.log 2>&1
tail -2 gpurun_out/ncu_gag.log
