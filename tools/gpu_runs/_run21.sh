#!/bin/bash
for lg in 22 19; do REPS=5 timeout 300 python tools/prof_kernels.py msm $lg 2>&1 | grep -E "msm_acc|table"; done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-curves --no-wide --no-big-hash --no-agg > gpurun_out/bench_r3_c.json 2> gpurun_out/bench_r3_c.err; tail -c 300 gpurun_out/bench_r3_c.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r3_c.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value']); m=d['msm']; print(m['value'], m['ms_per_msm'], [(k['name'], round(k['ms_per_step'],3)) for k in m['kernels']])
PY
