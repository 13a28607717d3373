#!/bin/bash
# r3: warp-resident Poseidon with global per-lane constants, radix-8 Fr passes, NTT grid swap
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fr_domain.py tests/test_gpu_groth16.py tests/test_gpu_stark.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-curves > gpurun_out/bench_r3_b.json 2> gpurun_out/bench_r3_b.err; tail -c 400 gpurun_out/bench_r3_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r3_b.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'])
for k in d.get('kernels',[]): print(k['name'], k['launches_per_step'], round(k['ms_per_step'],3))
for b in ('groth16_h','fr_domain','msm','wide','agg'):
    if b in d: print(b, json.dumps(d[b])[:600])
PY
