#!/bin/bash
# A/B of Poseidon-GL build variants on a B200 (run under gpurun).  Each variant recompiles merkle.cu with extra -D flags and
# links its own library next to the default one; tools/prof_kernels.py times the tree and leaf kernels through the C-ABI.
# usage: tools/ab_poseidon.sh "name1:-DFOO=1 -DBAR=0" "name2:..."      (built HERE, before the gpurun call, with BUILD=1)
set -e
cd "$(dirname "$0")/.."
V=eigen_zkvm_b200/build/variants
mkdir -p $V
NVCC=/usr/local/cuda/bin/nvcc
if [ -n "$BUILD" ]; then
  for spec in "$@"; do
    name=${spec%%:*}; flags=${spec#*:}
    $NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3 -ccbin /usr/bin/g++ -x cu $flags -c eigen_zkvm_b200/csrc/merkle.cu -o $V/merkle_$name.o &
  done
  wait
  for spec in "$@"; do
    name=${spec%%:*}
    objs=$(ls eigen_zkvm_b200/build/*.o | grep -v merkle.cu.o)
    $NVCC -shared -o $V/lib_$name.so $objs $V/merkle_$name.o -lcudart -ldl -ccbin /usr/bin/g++
  done
  exit 0
fi
for spec in "$@"; do
  name=${spec%%:*}
  echo "== variant $name (${spec#*:})"
  B200ZK_LIB=$PWD/$V/lib_$name.so REPS=3 python tools/prof_kernels.py merkle 24 2 | grep -E "merkle_level|linearhash"
  B200ZK_LIB=$PWD/$V/lib_$name.so REPS=3 python tools/prof_kernels.py merkle 21 48 | grep -E "linearhash"
done
