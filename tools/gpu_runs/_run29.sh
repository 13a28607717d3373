#!/bin/bash
timeout 900 python bench.py > gpurun_out/bench_r3_h.json 2> gpurun_out/bench_r3_h.err; tail -c 300 gpurun_out/bench_r3_h.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r3_h.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
m=d['msm']; print('msm', m['value'], m['ms_per_msm'], m['e2e']['value'], m.get('cpu_baseline'))
print('lde', d['lde_merkle']['value'], 'agg', d['aggregation']['value'])
print(json.dumps(d.get('groth16_h', d.get('fr_domain', {})))[:300])
PY
