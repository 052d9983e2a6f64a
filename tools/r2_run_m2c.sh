#!/bin/bash
export PYTHONUNBUFFERED=1
N=2
O=gpurun_out/r2m2c
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -3
tail -5 gpurun_out/multi_check_patch_shard.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus $N --no-train > ${O}_bench.json 2> ${O}_bench.err; python - <<PY
import json
for line in open('${O}_bench.json'):
    if line.startswith('{'):
        d=json.loads(line); print('N=2', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], 'weak', round(d['weak']['value']), round(d['weak']['e2e']['value']))
PY
tail -3 ${O}_bench.err
