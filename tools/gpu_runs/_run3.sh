set -x
python -m pytest tests/test_gpu_groth16.py -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 3 --warmup 3 --no-wide --no-big-hash --no-verify --log-n 20 > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err; tail -3 gpurun_out/bench_r2_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_b.json'))
m=d['msm']
print(m['value'], m['ms_per_msm'], m['mode']); print(m['e2e']); print(m['kernels']); print(m['plain'])
print(d.get('msm_other_curves'))
PY
