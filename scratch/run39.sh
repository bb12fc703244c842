cd /root/repo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gag_bwd -c 6 -o gpurun_out/prof_r1_gag_bwd2 -f python scratch/gag_one.py > gpurun_out/ncu_gag_bwd2.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_gag_bwd2.log
