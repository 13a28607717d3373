#!/bin/bash
REPS=5 timeout 300 python tools/prof_kernels.py msm 22 2>&1 | grep -E "msm_|table" | awk '{print $1, $2, $3, $4, $5, $8}'
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-wide --no-big-hash --no-agg --no-verify --log-n 20 > gpurun_out/bench_r3_g.json 2> gpurun_out/bench_r3_g.err; tail -c 300 gpurun_out/bench_r3_g.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r3_g.json').read().strip().splitlines()[-1])
m=d['msm']; print(m['value'], m['ms_per_msm'], m['e2e']['value'], m['mode'][:90]); print([(k['name'], round(k['ms_per_step'],3)) for k in m['kernels']])
for k in ('other_curves','msm_other_curves'):
    if k in d: print(json.dumps(d[k])[:400])
PY
timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_groth16.py -x -q -m gpu 2>&1 | tail -1
