#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
for envs in "SEG3D_FUSE_UP=0" "SEG3D_CIN1_TOEPLITZ=0" "SEG3D_FUSE_IN=0" "SEG3D_ZM_SPLIT=0" "SEG3D_FUSE_TAIL=0"; do
  echo "== $envs"; env $envs timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_sliding.py tests/test_gpu_train.py -m gpu -q -x -k "forward or golden or sliding or train_step or gradients" 2>&1 | tail -2
done
