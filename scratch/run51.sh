cd /root/repo
for sh in "64 128" "256 32"; do
  set -- $sh
  timeout 280 ncu --set full --clock-control none --import-source on -k regex:gag -s 5 -c 5 -o gpurun_out/gag_full_$1 -f python scratch/gag_one.py $1 $2 > gpurun_out/gag_full_$1.log 2>&1
  echo "ncu rc=$? $sh"
done
ls -la gpurun_out/*.ncu-rep
