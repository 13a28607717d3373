python -m pytest tests/test_gpu_parity.py tests/test_gpu_stark.py tests/test_gpu_big_hash.py -x -q -m gpu 2>&1 | tail -4
tools/ab_poseidon.sh "b0:x" "b1:x" 2>&1 | tail -12
