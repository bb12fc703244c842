cd /root/repo
EEGAN_H_WIDE=0 timeout 120 python scratch/h_probe3.py 2>&1 | tail -3
EEGAN_H_WIDE=1 timeout 120 python scratch/h_probe3.py 2>&1 | tail -3
