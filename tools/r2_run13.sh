#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2m
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sliding.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -5 > ${O}_pytest.log; cat ${O}_pytest.log
timeout 400 python bench.py --layers --no-train --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err; python -c "
import json; d=json.load(open('${O}_bench.json')); print('default b36', d['value'], d['e2e']['value'], d['ms_per_step'])"; grep -E "KIND" ${O}_bench.err
timeout 400 python bench.py --layers --no-train --no-cpu-baseline --mode fp32x > ${O}_bench_fp32x.json 2> ${O}_bench_fp32x.err; python -c "
import json; d=json.load(open('${O}_bench_fp32x.json')); print('fp32x', d['value'], d['e2e']['value'], d['ms_per_step'])"; grep -vE "Warning|warn" ${O}_bench_fp32x.err | tail -70
timeout 400 python bench.py --layers --no-train --no-cpu-baseline --arch vbnet --classes 5 --mode fp32x > ${O}_bench_vb_fp32x.json 2> ${O}_bench_vb_fp32x.err; python -c "
import json; d=json.load(open('${O}_bench_vb_fp32x.json')); print('vbnet fp32x', d['value'], d['e2e']['value'], d['ms_per_step'])"; grep -E "KIND" ${O}_bench_vb_fp32x.err
