#!/bin/bash
# window choice with the fitted per-bucket cost: per-rank sizes of 1 / 2 / 4 / 8 GPUs
for lg in 22 21 20 19; do REPS=5 timeout 300 python tools/prof_kernels.py msm $lg 2>&1 | grep -E "msm_|table" | awk '{print $1, $2, $3, $4, $5, $8}'; done
timeout 600 python -m pytest tests/test_gpu_msm.py -x -q -m gpu -k "table" 2>&1 | tail -1
