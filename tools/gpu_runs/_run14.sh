#!/bin/bash
# r3: TMA NTT parity + A/B timing
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ntt or interpolate or const_tree or sharded" 2>&1 | tail -15
for v in 0 1; do
  echo "== B200_NTT_TMA=$v"
  B200_NTT_TMA=$v REPS=5 timeout 300 python tools/prof_kernels.py ntt 24 2 2>&1 | tail -4
  B200_NTT_TMA=$v REPS=5 timeout 300 python tools/prof_kernels.py lde 24 2 2>&1 | tail -4
  B200_NTT_TMA=$v REPS=3 timeout 300 python tools/prof_kernels.py lde 20 64 2>&1 | tail -4
done
