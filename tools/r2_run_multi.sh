#!/bin/bash
# Round-2 multi-GPU call: gpurun --gpus N -- 'bash tools/r2_run_multi.sh N'
export PYTHONUNBUFFERED=1
N=${1:-2}
O=gpurun_out/r2m${N}
mkdir -p gpurun_out
nvidia-smi -L > ${O}_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_fullsize.py -m gpu -q -rs 2>&1 | tail -30 > ${O}_pytest.log; tail -12 ${O}_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus $N > ${O}_bench.json 2> ${O}_bench.err; cut -c1-3500 ${O}_bench.json; tail -5 ${O}_bench.err
timeout 300 $TR --master-port 29542 bench.py --gpus $N --gather mask --no-train --no-weak > ${O}_bench_mask.json 2> ${O}_bench_mask.err; cut -c1-600 ${O}_bench_mask.json; tail -3 ${O}_bench_mask.err
timeout 300 $TR --master-port 29543 bench.py --gpus $N --impl reference --steps 1 --warmup 0 > ${O}_bench_ref.json 2> ${O}_bench_ref.err; cut -c1-900 ${O}_bench_ref.json; tail -3 ${O}_bench_ref.err
