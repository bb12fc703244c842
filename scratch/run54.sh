cd /root/repo
for sh in "64 128" "256 32"; do
  set -- $sh
  timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gag_launch2_$1.csv python scratch/gag_one.py $1 $2 > /dev/null 2>&1
  echo "== $sh"; python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/gag_launch2_$1.csv')) if len(r)>5 and r[0].isdigit()]
out=[(r[4][:60], r[-1]) for r in rows if 'gag' in r[4]]
for k,v in out[len(out)//2:]: print(k, v)
PY
done
