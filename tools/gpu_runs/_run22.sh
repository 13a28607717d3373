#!/bin/bash
# ncu of one small warp-per-node Merkle level (latency case): 2^14 leaves -> the 64-node level is the 7th warp launch
REPS=1 ncu --set full --clock-control none --import-source on -k regex:k_merkle_level_warp -s 6 -c 1 -o gpurun_out/prof_mwarp_r3b python tools/prof_kernels.py merkle 14 2 > gpurun_out/prof_mwarp.log 2>&1
ncu -i gpurun_out/prof_mwarp_r3b.ncu-rep --page raw --csv > gpurun_out/prof_mwarp_r3b_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_mwarp_r3b.ncu-rep --page source --csv > gpurun_out/prof_mwarp_r3b_source.csv 2>/dev/null
tail -3 gpurun_out/prof_mwarp.log; ls -la gpurun_out/prof_mwarp*
