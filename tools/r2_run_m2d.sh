#!/bin/bash
export PYTHONUNBUFFERED=1
N=2
O=gpurun_out/r2m2d
mkdir -p gpurun_out
for i in 1 2; do timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -2; grep -E "^rank" gpurun_out/multi_check_overlap_allreduce.log | cut -c1-220; done
cp gpurun_out/multi_check_overlap_allreduce.log ${O}_overlap.log; cp gpurun_out/multi_check_patch_shard.log ${O}_patch_shard.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus $N > ${O}_bench.json 2> ${O}_bench.err; python - <<PY
import json
for line in open('${O}_bench.json'):
    if line.startswith('{'):
        d=json.loads(line); print('N=2', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], 'weak', round(d['weak']['value']), round(d['weak']['e2e']['value']), 'train', round(d['train']['value']))
PY
