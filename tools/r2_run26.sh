#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2y
mkdir -p gpurun_out
for i in 1 2; do
for lo in 1 0; do
  SEG3D_CIN1_LO=$lo timeout 400 python bench.py --no-train --no-cpu-baseline > ${O}_lo${lo}_$i.json 2> ${O}_lo${lo}_$i.err; python -c "
import json; d=json.load(open('${O}_lo${lo}_$i.json')); print('LO=$lo', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2))"
done; done
SEG3D_CIN1_LO=0 timeout 900 python -m pytest tests -m gpu -q -x -k "not cin1_toeplitz" 2>&1 | tail -3
SEG3D_CIN1_LO=0 timeout 300 python bench.py --task train --no-cpu-baseline 2>/dev/null | cut -c1-200
