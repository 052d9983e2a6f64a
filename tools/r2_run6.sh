#!/bin/bash
# Round-2 baseline call after the container was re-created (one B200): full -m gpu suite, headline bench with the per-layer
# table, VBNet bench, then the ncu launch lists of the inference forward and of the training step (shares only).
export PYTHONUNBUFFERED=1
O=gpurun_out/r2f
mkdir -p gpurun_out
nvidia-smi -L > ${O}_gpus.txt
timeout 900 python -m pytest tests -m gpu -q -rs 2>&1 | tail -40 > ${O}_pytest.log; tail -25 ${O}_pytest.log
timeout 500 python bench.py --layers > ${O}_bench.json 2> ${O}_bench.err; cut -c1-3000 ${O}_bench.json; tail -12 ${O}_bench.err
timeout 300 python bench.py --arch vbnet --classes 5 --mode auto --no-train --no-cpu-baseline > ${O}_bench_vbnet.json 2> ${O}_bench_vbnet.err; cut -c1-400 ${O}_bench_vbnet.json; tail -3 ${O}_bench_vbnet.err
timeout 200 python tools/profile_forward.py 20 fp16 > ${O}_fwd_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_infer_launches.csv python tools/profile_forward.py 20 fp16 > ${O}_fwd_ncu.log 2>&1
python tools/ncu_launches.py ${O}_infer_launches.csv 2 > ${O}_infer_launches_summary.txt; head -30 ${O}_infer_launches_summary.txt
timeout 200 python tools/train_one_step.py bf16 8 4 > ${O}_train_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_train_launches.csv python tools/train_one_step.py bf16 8 4 > ${O}_train_ncu.log 2>&1
python tools/ncu_launches.py ${O}_train_launches.csv 4 > ${O}_train_launches_summary.txt; head -45 ${O}_train_launches_summary.txt
