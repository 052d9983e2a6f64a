#!/bin/bash
# experiments: input-block weights hi-only (parity + speed), transposed convs run twice with GroupNorm in the epilogue
export PYTHONUNBUFFERED=1
O=gpurun_out/r2j
mkdir -p gpurun_out
SEG3D_CIN1_LO=0 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_sliding.py tests/test_gpu_fullsize.py -m gpu -q -k "not cin1" 2>&1 | tail -15 > ${O}_pytest_lo0.log; cat ${O}_pytest_lo0.log
cp gpurun_out/bf16_parity_vnet_c2.json ${O}_parity_vnet_lo0.json; cat ${O}_parity_vnet_lo0.json
SEG3D_CIN1_LO=0 timeout 400 python bench.py --layers --no-train > ${O}_bench_lo0.json 2> ${O}_bench_lo0.err; python -c "
import json; d=json.load(open('${O}_bench_lo0.json')); print('LO=0', d['value'], d['e2e']['value'], d['parity'])"; grep -E "in_block|KIND" ${O}_bench_lo0.err
SEG3D_FUSE_S2=up timeout 400 python bench.py --layers --no-train > ${O}_bench_up.json 2> ${O}_bench_up.err; python -c "
import json; d=json.load(open('${O}_bench_up.json')); print('FUSE_S2=up', d['value'], d['e2e']['value'], d['parity'])"; grep -E "up_conv|up_gn|KIND" ${O}_bench_up.err
timeout 400 python bench.py --layers --no-train > ${O}_bench.json 2> ${O}_bench.err; python -c "
import json; d=json.load(open('${O}_bench.json')); print('default', d['value'], d['e2e']['value'], d['parity'])"
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "forward or 96" 2>&1 | tail -5; cat gpurun_out/bf16_parity_vnet_c2.json
