#!/bin/bash
# First GPU call of round 2 (one B200): everything that was written after round 1's GPU budget was spent, plus the two
# unrun batch-size experiments of DESIGN.md section 7.  Usage:
#   gpurun --timeout 900 -- 'bash tools/round2_first_run.sh'        (outputs under gpurun_out/r2_*)
export PYTHONUNBUFFERED=1
O=gpurun_out/r2
mkdir -p gpurun_out
# 1. the regular suite, then the gated tests (standalone network modules, device crops, segmentation_voi, DISABLE partition)
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > ${O}_pytest.log
SEG3D_TEST_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_blocks.py -m gpu -q -s 2>&1 | tail -40 > ${O}_pytest_unverified.log
cat ${O}_pytest.log; tail -15 ${O}_pytest_unverified.log
# 2. device-side training crops through the e2e training flow
SEG3D_DEVICE_CROPS=1 timeout 200 python -m pytest tests/test_gpu_e2e.py -m gpu -q 2>&1 | tail -5 > ${O}_e2e_device_crops.log; cat ${O}_e2e_device_crops.log
# 3. headline bench, then patch batches above 20 and small batches replayed as CUDA graphs
timeout 300 python bench.py > ${O}_bench_b20.json 2> ${O}_bench_b20.err; cut -c1-220 ${O}_bench_b20.json
for B in 36 45 60; do
  timeout 200 python bench.py --batch $B --no-cpu-baseline > ${O}_bench_b${B}.json 2> ${O}_bench_b${B}.err; cut -c1-220 ${O}_bench_b${B}.json
done
for B in 4 6 10; do
  SEG3D_GRAPH=1 timeout 200 python bench.py --batch $B --no-cpu-baseline > ${O}_bench_graph_b${B}.json 2> ${O}_bench_graph_b${B}.err; cut -c1-220 ${O}_bench_graph_b${B}.json
done
# 4. (two GPUs, separate call: gpurun --gpus 2) patch-sharded inference with the label exchange:
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
#       --shard patches --gather labels --no-cpu-baseline ;  same with --gather mask ;  tools/check_patch_shard.py for parity
