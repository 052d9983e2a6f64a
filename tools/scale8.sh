# 8 x B200 runs for profiles/r01_scaling_notes.md (one process per GPU, torchrun, NCCL)
export PYTHONUNBUFFERED=1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
O=gpurun_out/s30
$T bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline 2>${O}_n8_cases.err | tee ${O}_n8_cases.json | cut -c1-200
$T bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --shard patches 2>${O}_n8_patches.err | tee ${O}_n8_patches.json | cut -c1-200
$T bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --shard patches --gather probs 2>${O}_n8_patches_probs.err | tee ${O}_n8_patches_probs.json | cut -c1-200
$T bench.py --gpus 8 --steps 5 --warmup 3 --task train --mode bf16 2>${O}_n8_train.err | tee ${O}_n8_train.json | cut -c1-200
$T bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --shard patches --arch vbnet --classes 5 2>${O}_n8_vbnet.err | tee ${O}_n8_vbnet.json | cut -c1-200
$T bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu-baseline --mode fp32x --volume 256,256,256 --batch 9 2>${O}_n8_fp32x_256.err | tee ${O}_n8_fp32x_256.json | cut -c1-200
