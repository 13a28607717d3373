#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_stark.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-curves --no-wide --no-big-hash --no-msm > gpurun_out/bench_r3_e.json 2> gpurun_out/bench_r3_e.err; tail -c 300 gpurun_out/bench_r3_e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r3_e.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'])
print([(k['name'], round(k['ms_per_step'],3)) for k in d['kernels'] if k['name'] in ('step_program','fri_fold','lagrange_row','xdivxsub','eval_dot')])
a=d['aggregation']; print(a['value'], [(k['name'], round(k['ms_per_proof'],3)) for k in a['kernels'] if k['name'] in ('step_program','fri_fold','lagrange_row','eval_dot')])
PY
