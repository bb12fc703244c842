cd /root/repo
timeout 300 python scratch/gag_time2.py 2>&1 | tail -4
