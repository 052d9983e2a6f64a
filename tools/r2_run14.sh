#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2n
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "split" 2>&1 | tail -15 > ${O}_pytest_split.log; cat ${O}_pytest_split.log
timeout 400 python bench.py --layers --no-train --no-cpu-baseline --mode fp32x > ${O}_bench_fp32x.json 2> ${O}_bench_fp32x.err; python -c "
import json; d=json.load(open('${O}_bench_fp32x.json')); print('fp32x', d['value'], d['e2e']['value'], d['ms_per_step'])"; grep -E "KIND|up_32.rblock|out_block.conv1|down_32.rblock" ${O}_bench_fp32x.err
timeout 400 python bench.py --layers --no-train --arch vbnet --classes 5 --mode fp32x > ${O}_bench_vb_fp32x.json 2> ${O}_bench_vb_fp32x.err; python -c "
import json; d=json.load(open('${O}_bench_vb_fp32x.json')); print('vbnet fp32x', d['value'], d['e2e']['value'], d['ms_per_step'], d['parity'])"; grep -E "KIND" ${O}_bench_vb_fp32x.err
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > ${O}_pytest.log; cat ${O}_pytest.log
