#!/bin/bash
# A/B: NTT tile shapes (default = 2048 elements)
for v in default T8 T4 TILE1024 TILE4096; do
  if [ $v = default ]; then lib=$PWD/eigen_zkvm_b200/libb200zk.so; else lib=$PWD/eigen_zkvm_b200/build/exp/libb200zk_$v.so; fi
  echo "== $v"
  for shape in "ntt 24 2" "lde 24 2" "lde 20 64"; do
    B200ZK_LIB=$lib REPS=5 timeout 300 python tools/prof_kernels.py $shape 2>&1 | grep -E "pass" | awk '{print $1, $8, $12, $13}'
  done
done
for v in T8 T4 TILE1024; do B200ZK_LIB=$PWD/eigen_zkvm_b200/build/exp/libb200zk_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ntt or interpolate" 2>&1 | tail -1; done
