#!/bin/bash
export PYTHONUNBUFFERED=1
for i in 1 2 3 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600+i)) tools/check_overlap_allreduce.py 2>&1 | grep -E "^rank" | cut -c1-400
done
echo "== CIN1_TOEPLITZ=0"
for i in 1 2 3; do
SEG3D_CIN1_TOEPLITZ=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29610+i)) tools/check_overlap_allreduce.py 2>&1 | grep -E "^rank" | cut -c1-400
done
