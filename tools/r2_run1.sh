#!/bin/bash
# Round-2 first GPU call: the full -m gpu suite with the formerly gated tests, headline bench, batch-size experiments.
export PYTHONUNBUFFERED=1
O=gpurun_out/r2
mkdir -p gpurun_out
nvidia-smi -L > ${O}_gpus.txt
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -60 > ${O}_pytest.log; tail -30 ${O}_pytest.log
SEG3D_DEVICE_CROPS=1 timeout 200 python -m pytest tests/test_gpu_e2e.py -m gpu -q 2>&1 | tail -5 > ${O}_e2e_device_crops.log; cat ${O}_e2e_device_crops.log
timeout 300 python bench.py > ${O}_bench_b20.json 2> ${O}_bench_b20.err; cut -c1-300 ${O}_bench_b20.json
for B in 36 60; do
  timeout 200 python bench.py --batch $B --no-cpu-baseline > ${O}_bench_b${B}.json 2> ${O}_bench_b${B}.err; cut -c1-200 ${O}_bench_b${B}.json
done
for B in 4 6 10; do
  SEG3D_GRAPH=1 timeout 200 python bench.py --batch $B --no-cpu-baseline > ${O}_bench_graph_b${B}.json 2> ${O}_bench_graph_b${B}.err; cut -c1-200 ${O}_bench_graph_b${B}.json
done
timeout 200 python bench.py --task train --no-cpu-baseline > ${O}_bench_train.json 2> ${O}_bench_train.err; cut -c1-300 ${O}_bench_train.json
