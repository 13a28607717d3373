set -x
python -m pytest tests/test_gpu_groth16.py tests/test_gpu_c12_exec.py tests/test_gpu_big_hash.py tests/test_gpu_stark.py tests/test_gpu_msm.py -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 3 --warmup 3 --no-wide --no-verify --log-n 20 --no-cpu-baseline > gpurun_out/bench_r2_c.json 2> gpurun_out/bench_r2_c.err; tail -3 gpurun_out/bench_r2_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_c.json'))
m=d['msm']
print(m['value'], m['ms_per_msm'], m['mode']); print(m['e2e']); print(m['kernels']); print(m['plain'])
print(d.get('msm_other_curves'))
print(d.get('big_hash_merkle'))
print(d.get('aggregation'))
PY
