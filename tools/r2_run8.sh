#!/bin/bash
# ncu launch list of one whole sliding-window step (gather, forwards, blend, finalize): shares only
export PYTHONUNBUFFERED=1
O=gpurun_out/r2h
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_step_launches.csv python bench.py --steps 1 --warmup 1 --no-train --no-cpu-baseline > ${O}_step_ncu.log 2>&1
python tools/ncu_launches.py ${O}_step_launches.csv 1 > ${O}_step_launches_summary.txt; head -40 ${O}_step_launches_summary.txt
