#!/bin/bash
# ncu launch lists of the round-2 training step and inference forward (shares only: a number under ncu is never a bench value)
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e
mkdir -p gpurun_out
timeout 200 python tools/train_one_step.py bf16 8 4 > ${O}_train_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_train_launches.csv python tools/train_one_step.py bf16 8 4 > ${O}_train_ncu.log 2>&1
python tools/ncu_launches.py ${O}_train_launches.csv 4 > ${O}_train_launches_summary.txt; head -45 ${O}_train_launches_summary.txt
