#!/bin/bash
# ncu --set full of the Toeplitz input-block kernel (both passes) in one B=20 forward
export PYTHONUNBUFFERED=1
O=gpurun_out/r2l
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cin1_toeplitz -c 4 -o ${O}_cin1t python tools/profile_forward.py 20 fp16 > ${O}_ncu.log 2>&1
tail -3 ${O}_ncu.log
ncu -i ${O}_cin1t.ncu-rep --page raw --csv > ${O}_cin1t_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
