#!/bin/bash
# ncu captures of the hot kernels (run under gpurun; keeps gpurun_out/ below the 64 MiB copy-back limit)
set -x
TAG=${TAG:-r3a}
NCU="ncu --set full --clock-control none --import-source on"
REPS=1 $NCU -k regex:^k_merkle_level$ -s 1 -c 1 -o gpurun_out/prof_merkle_$TAG python tools/prof_kernels.py merkle 22 2 > gpurun_out/prof_merkle.log 2>&1
REPS=1 $NCU -k regex:k_linearhash -c 1 -o gpurun_out/prof_lh_$TAG python tools/prof_kernels.py merkle 20 48 > gpurun_out/prof_lh.log 2>&1
REPS=1 $NCU -k regex:k_ntt3 -c 3 -o gpurun_out/prof_ntt_$TAG python tools/prof_kernels.py ntt 24 2 > gpurun_out/prof_ntt.log 2>&1
if [ -n "$WITH_MSM" ]; then REPS=1 $NCU -k regex:k_msm_accumulate -c 1 -o gpurun_out/prof_msm_$TAG python tools/prof_kernels.py msm 22 > gpurun_out/prof_msm.log 2>&1; fi
if [ -n "$WITH_EVAL" ]; then REPS=1 $NCU -k regex:k_eval -c 2 -o gpurun_out/prof_eval_$TAG python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-msm --no-wide --log-n 22 > gpurun_out/prof_eval.log 2>&1; fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${TAG}_bench_fib24.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-msm --no-wide --no-big-hash --no-agg --no-verify > gpurun_out/launches.log 2>&1
for f in gpurun_out/prof_*_$TAG.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
  ncu -i $f --page source --csv > ${f%.ncu-rep}_source.csv 2>/dev/null
done
du -sh gpurun_out; ls -la gpurun_out
if [ $(du -sm gpurun_out | cut -f1) -gt 60 ]; then rm -f gpurun_out/*.ncu-rep; fi
