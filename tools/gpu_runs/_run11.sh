free -g | head -2; nproc
/usr/bin/time -v python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/ref_full.json 2> gpurun_out/ref_full.err; echo rc=$?
grep -E "Maximum resident|Elapsed" gpurun_out/ref_full.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/ref_full.json'))
print(d['value'], d['steps_measured'], d['config']['workload']); print(d['config']['reference_run']); print(d['cpu_baseline']['phases_s'])
PY
