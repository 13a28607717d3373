set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 --no-msm --no-wide --no-big-hash > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; tail -3 gpurun_out/bench_r2_a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_a.json'))
print(d['value'], d['e2e'], d['gpu_launches'], d.get('verification'))
for k in d['kernels']: print(k['name'], k['launches_per_step'], round(k['ms_per_step'],3), round(k['algo_GBps'],1))
print(d.get('cpu_baseline'))
PY
