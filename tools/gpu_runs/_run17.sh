#!/bin/bash
# r3: ADVICE fixes (errors, msm range flag) on 1 GPU, then the 2-GPU bench
timeout 900 python -m pytest tests/test_gpu_errors.py tests/test_gpu_msm.py tests/test_gpu_groth16.py -x -q -m gpu 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r3_n2.json 2> gpurun_out/bench_r3_n2.err; tail -c 600 gpurun_out/bench_r3_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r3_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['n_gpus'])
for b in ('msm','wide','agg'):
    if b in d: print(b, json.dumps({k:v for k,v in d[b].items() if k in ('value','unit','ms_per_msm','e2e','seconds','s_per_batch','sharded_s','single_s','ms','root_equal','n_gpus')})[:500])
PY
