#!/bin/bash
# r3: full GPU suite + bench + ncu of the TMA NTT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r3_a.json 2> gpurun_out/bench_r3_a.err; tail -c 600 gpurun_out/bench_r3_a.err
NCU="ncu --set full --clock-control none --import-source on"
REPS=1 timeout 600 $NCU -k regex:k_ntt3 -c 3 -o gpurun_out/prof_ntt_r3a python tools/prof_kernels.py ntt 24 2 > gpurun_out/prof_ntt.log 2>&1
ncu -i gpurun_out/prof_ntt_r3a.ncu-rep --page raw --csv > gpurun_out/prof_ntt_r3a_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_ntt_r3a.ncu-rep --page source --csv > gpurun_out/prof_ntt_r3a_source.csv 2>/dev/null
ls -la gpurun_out | grep r3
