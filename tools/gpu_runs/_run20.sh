#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_errors.py tests/test_gpu_groth16.py -x -q -m gpu 2>&1 | tail -4
for lg in 19 20 22; do REPS=5 timeout 300 python tools/prof_kernels.py msm $lg 2>&1 | grep -E "msm_|table"; done
