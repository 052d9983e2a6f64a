#!/bin/bash
export PYTHONUNBUFFERED=1
N=2
O=gpurun_out/r2m2b
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -6 > ${O}_pytest.log; cat ${O}_pytest.log
tail -4 gpurun_out/multi_check_patch_shard.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus $N > ${O}_bench.json 2> ${O}_bench.err; python -c "
import json; d=json.load(open('${O}_bench.json')); print('N=2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'weak', d['weak']['value'], 'train', d['train']['value'], d['train']['ms_per_step'])"; tail -3 ${O}_bench.err
timeout 300 $TR --master-port 29543 bench.py --gpus $N --impl reference --steps 1 --warmup 0 > ${O}_bench_ref.json 2> ${O}_bench_ref.err; cut -c1-300 ${O}_bench_ref.json
CUDA_VISIBLE_DEVICES=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_train_launches.csv python tools/train_one_step.py bf16 8 4 > ${O}_train_ncu.log 2>&1
