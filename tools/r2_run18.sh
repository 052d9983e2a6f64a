#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2r
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "narrow or up_block" 2>&1 | tail -4
for i in 1 2; do
timeout 400 python bench.py --layers --no-train --no-cpu-baseline > ${O}_bench$i.json 2> ${O}_bench$i.err; python -c "
import json; d=json.load(open('${O}_bench$i.json')); print('default', d['value'], d['e2e']['value'], d['ms_per_step'])"; grep -E "KIND conv_tc_narrow|KIND conv_tc_k3|KIND gn" ${O}_bench$i.err
done
SEG3D_FUSE_UP=0 timeout 400 python bench.py --layers --no-train --no-cpu-baseline > ${O}_bench_nofuse.json 2> ${O}_bench_nofuse.err; python -c "
import json; d=json.load(open('${O}_bench_nofuse.json')); print('FUSE_UP=0', d['value'], d['e2e']['value'], d['ms_per_step'])"; grep -E "KIND conv_tc_narrow|KIND conv_tc_k3|KIND gn" ${O}_bench_nofuse.err
