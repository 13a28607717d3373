python -m pytest tests/test_gpu_stark.py -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 2 --warmup 3 --no-msm --no-wide --no-cpu-baseline --no-big-hash --no-agg > gpurun_out/bench_r2_f.json 2> gpurun_out/bench_r2_f.err; echo rc=$?; tail -3 gpurun_out/bench_r2_f.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_f.json'))
print(d['value'], d['e2e']['value'], d['self_verify'], d['verification'])
PY
