#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2s
mkdir -p gpurun_out
L=$PWD/medical-segmentation3d-toolkit_b200/lib
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "narrow or up_block" 2>&1 | tail -2
SEG3D_LIB=$L/variant_xw8.so timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "narrow or up_block" 2>&1 | tail -2
for cfg in "default(minb2,xw4):" "xw8_minb2:SEG3D_LIB=$L/variant_xw8.so" "xw8_minb1:SEG3D_LIB=$L/variant_xw8m1.so" "xw4_minb1:SEG3D_LIB=$L/variant_m1.so" "default_nofuseup:SEG3D_FUSE_UP=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 400 python bench.py --layers --no-train --no-cpu-baseline > "${O}_$name.json" 2> "${O}_$name.err"; python -c "
import json; d=json.load(open('${O}_$name.json')); print('$name', d['value'], d['e2e']['value'], d['ms_per_step'])"; grep -E "KIND conv_tc_narrow" "${O}_$name.err"
done
