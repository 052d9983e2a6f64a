#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2v
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sliding.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do timeout 400 python bench.py --no-train --no-cpu-baseline > ${O}_bench$i.json 2> ${O}_bench$i.err; python -c "
import json; d=json.load(open('${O}_bench$i.json')); print('default', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2))"; done
