#!/bin/bash
# final state of the round: full GPU suite, smoke, default bench line, ncu round r3b
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_r3_f.json 2> gpurun_out/bench_r3_f.err; tail -c 300 gpurun_out/bench_r3_f.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r3_f.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
print([(r['kernel'], round(r['frac'],3)) for r in d['roofline_ntt']])
print([(k['name'], k['launches_per_step'], round(k['ms_per_step'],3)) for k in d['kernels']])
print('msm', d['msm']['value'], 'lde', d['lde_merkle']['value'], [(k['name'], round(k['ms_per_step'],2), round(k.get('algo_GBps',0))) for k in d['lde_merkle']['kernels'][:2]], 'agg', d['aggregation']['value'])
PY
WITH_MSM=1 TAG=r3b bash tools/prof_round.sh > gpurun_out/prof_round_r3b.log 2>&1; tail -3 gpurun_out/prof_round_r3b.log
