export PYTHONUNBUFFERED=1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/s12_n8_cases.err | tee gpurun_out/s12_n8_cases.json | cut -c1-260
$T bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --shard patches 2>gpurun_out/s12_n8_patches.err | tee gpurun_out/s12_n8_patches.json | cut -c1-260
$T bench.py --gpus 8 --steps 5 --warmup 3 --task train --mode bf16 2>gpurun_out/s12_n8_train.err | tee gpurun_out/s12_n8_train.json | cut -c1-260
$T bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --shard patches --arch vbnet --classes 5 2>gpurun_out/s12_n8_vbnet.err | tee gpurun_out/s12_n8_vbnet.json | cut -c1-260
