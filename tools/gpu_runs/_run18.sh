#!/bin/bash
# r3: sanitizers on the small walk, then the ncu round (merkle, linearhash, ntt, msm) + launch list
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/sanitizer_memcheck_r3.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck_r3.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/sanitizer_racecheck_r3.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck_r3.log
WITH_MSM=1 TAG=r3a bash tools/prof_round.sh > gpurun_out/prof_round_r3a.log 2>&1; tail -5 gpurun_out/prof_round_r3a.log
