#!/bin/bash
# ncu --set full (with source counters) of the fused narrow-output kernel in one B=20 forward
export PYTHONUNBUFFERED=1
O=gpurun_out/r2q
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fold -c 2 -o ${O}_fold python tools/profile_forward.py 20 fp16 > ${O}_ncu.log 2>&1
tail -3 ${O}_ncu.log
ncu -i ${O}_fold.ncu-rep --page raw --csv > ${O}_fold_raw.csv 2>/dev/null
ncu -i ${O}_fold.ncu-rep --page source --csv --print-source cuda > ${O}_fold_source.csv 2>/dev/null
ls -la gpurun_out | tail -4
