# big-field Poseidon with lazy dot products: parity + bench blocks
python -m pytest tests/test_gpu_big_hash.py -x -q -m gpu 2>&1 | tail -5
python -m pytest tests/test_gpu_stark.py -x -q -m gpu -k "bn128 or bls or BN128 or BLS or wide" 2>&1 | tail -5
python bench.py --steps 2 --warmup 3 --no-msm --no-wide --no-cpu-baseline --no-verify > gpurun_out/bench_r2_e.json 2> gpurun_out/bench_r2_e.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_e.json'))
print(d['value'], d['e2e'])
print(json.dumps(d.get('big_hash_merkle'))[:1500])
a=d.get('aggregation'); print(a['value'], a['seconds_per_proof'], [(k['name'],round(k['ms_per_proof'],2)) for k in a['kernels'] if k['ms_per_proof']>1])
PY
