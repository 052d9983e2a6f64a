#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2d
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > ${O}_pytest.log; tail -12 ${O}_pytest.log
timeout 300 python bench.py --task train --no-cpu-baseline > ${O}_train.json 2> ${O}_train.err; cut -c1-400 ${O}_train.json; tail -3 ${O}_train.err
SEG3D_GN_BWD_REMASK=0 timeout 300 python bench.py --task train --no-cpu-baseline > ${O}_train_noremask.json 2> ${O}_train_noremask.err; cut -c1-300 ${O}_train_noremask.json
SEG3D_PACK_KERNEL=0 timeout 300 python bench.py --task train --no-cpu-baseline > ${O}_train_nopack.json 2> ${O}_train_nopack.err; cut -c1-300 ${O}_train_nopack.json
