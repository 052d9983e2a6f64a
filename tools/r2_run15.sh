#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2o
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "wgrad_cin1" 2>&1 | tail -15 > ${O}_pytest_wg.log; cat ${O}_pytest_wg.log
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_e2e.py -m gpu -q -x 2>&1 | tail -8 > ${O}_pytest_train.log; cat ${O}_pytest_train.log
timeout 300 python bench.py --task train --no-cpu-baseline > ${O}_train.json 2> ${O}_train.err; cut -c1-500 ${O}_train.json; tail -3 ${O}_train.err
