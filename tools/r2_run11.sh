#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2k
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cin1" 2>&1 | tail -5 > ${O}_pytest_cin1.log; cat ${O}_pytest_cin1.log
timeout 400 python bench.py --layers --no-train --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err; python -c "
import json; d=json.load(open('${O}_bench.json')); print('default', d['value'], d['e2e']['value'])"; grep -E "in_block|KIND" ${O}_bench.err
SEG3D_CIN1_LO=0 timeout 400 python bench.py --layers --no-train --no-cpu-baseline > ${O}_bench_lo0.json 2> ${O}_bench_lo0.err; grep -E "in_block" ${O}_bench_lo0.err
