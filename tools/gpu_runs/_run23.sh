#!/bin/bash
# final-state validation: full GPU suite, smoke(), default bench line
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r3_d.json 2> gpurun_out/bench_r3_d.err; tail -c 300 gpurun_out/bench_r3_d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r3_d.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks'], d['cpu_baseline'])
print(d['roofline']['frac'], [ (r['kernel'], round(r['frac'],3), r['int_issue']['thread_instr_per_unit']) for r in d['roofline_ntt']])
PY
