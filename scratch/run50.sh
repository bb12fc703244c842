cd /root/repo
for ns in 2 3 4 5 6; do
EEGAN_H_STAGES=$ns EEGAN_H_WIDE=0 timeout 120 python scratch/h_probe3.py 2>&1 | tail -3 | sed "s/^/stages=$ns /"
done
