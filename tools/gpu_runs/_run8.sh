python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_d.json 2> gpurun_out/bench_r2_d.err; echo rc=$?; grep -v "^$" gpurun_out/bench_r2_d.err | tail -4
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_d.json'))
print(d['value'], d['e2e'], d['gpu_launches'], d.get('verification'))
for k in d['kernels']: print(k['name'], k['launches_per_step'], round(k['ms_per_step'],3), round(k['algo_GBps'],1))
m=d['msm']; print(m['value'], m['ms_per_msm'], m['kernels'], m['e2e'], m['plain']['value'])
print(d.get('msm_other_curves'))
print([ (x['hash'], x['seconds']) for x in d.get('big_hash_merkle',[])])
a=d.get('aggregation'); print({k:v for k,v in a.items() if k!='kernels'} if a else None)
print(d.get('lde_merkle',{}).get('value'), d.get('groth16_h'))
print(d.get('cpu_baseline'))
PY
