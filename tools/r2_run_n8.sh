#!/bin/bash
# 8-GPU validation of what the driver's scaling run will launch: bench.py --gpus 8 (strong-scaling headline + weak + train)
export PYTHONUNBUFFERED=1
N=${1:-8}
O=gpurun_out/r2n${N}
mkdir -p gpurun_out
nvidia-smi -L > ${O}_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 bench.py --gpus $N > ${O}_bench.json 2> ${O}_bench.err; cut -c1-4000 ${O}_bench.json; tail -5 ${O}_bench.err
