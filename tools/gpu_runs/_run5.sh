set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --no-verify --log-n 22 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; tail -5 gpurun_out/bench_r2_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n2.json'))
print(d['value'], d['n_gpus'], d['e2e'])
m=d['msm']; print(m['value'], m['ms_per_msm'], m['mode']); print(m['e2e']); print(m['plain']['value'])
print(d.get('lde_merkle'))
a=d.get('aggregation'); print({k:v for k,v in a.items() if k!='kernels'} if a else None)
PY
