#!/bin/bash
# Second GPU call of round 2 (two B200s: gpurun --gpus 2 --timeout 600 -- 'bash tools/round2_two_gpu.sh'): the patch-sharded
# exchanges side by side, the label exchange included (written after round 1's GPU budget was spent; pinned on CPU under gloo).
export PYTHONUNBUFFERED=1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
O=gpurun_out/r2_n2
mkdir -p gpurun_out
$T tools/check_patch_shard.py 2>&1 | tail -5 > ${O}_check.log; cat ${O}_check.log
for G in mask probs labels; do
  $T bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --shard patches --gather $G 2>${O}_${G}.err | tee ${O}_${G}.json | cut -c1-200
done
$T bench.py --gpus 2 --steps 5 --warmup 3 --task train --mode bf16 --no-cpu-baseline 2>${O}_train.err | tee ${O}_train.json | cut -c1-200
