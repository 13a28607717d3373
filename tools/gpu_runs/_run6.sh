export PYTHONFAULTHANDLER=1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_stark.py tests/test_gpu_c12_exec.py tests/test_gpu_errors.py -x -q -m gpu 2>&1 | tail -6
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-wide --no-big-hash --no-agg --no-msm > gpurun_out/t6.json 2> gpurun_out/t6.err; echo rc=$?; grep -v "^$" gpurun_out/t6.err | tail -5
python - <<'PY'
import json
d=json.load(open('gpurun_out/t6.json'))
print(d['value'], d['e2e'], d['gpu_launches'], d.get('verification'))
for k in d['kernels']: print(k['name'], k['launches_per_step'], round(k['ms_per_step'],3), round(k['algo_GBps'],1))
PY
