NCU="ncu --set full --clock-control none --import-source on"
REPS=1 $NCU -k regex:^k_merkle_level$ -s 1 -c 1 -o gpurun_out/prof_merkle_r2a python tools/prof_kernels.py merkle 22 2 > gpurun_out/prof_merkle.log 2>&1
tail -3 gpurun_out/prof_merkle.log
for f in gpurun_out/prof_merkle_r2a.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
  ncu -i $f --page source --csv > ${f%.ncu-rep}_source.csv 2>/dev/null
done
ls -la gpurun_out
