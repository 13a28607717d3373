#!/bin/bash
# r3: the 8-GPU bench (replica proofs, MSM chunks + all-gather, sharded wide LDE + Merkle, configs[4] sub-proofs spread over ranks)
N=${N:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r3_n$N.json 2> gpurun_out/bench_r3_n$N.err; tail -c 500 gpurun_out/bench_r3_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r3_n$N.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['n_gpus'])
print('msm', d['msm']['value'], d['msm']['e2e']['value'], d['msm'].get('ms_per_msm'))
print('lde_merkle', d['lde_merkle']['value'], [ (k['name'], round(k['ms_per_step'],2)) for k in d['lde_merkle']['kernels']])
a=d['aggregation']; print('agg', a['value'], a.get('seconds_per_proof'), a.get('placement','')[:80])
PY
