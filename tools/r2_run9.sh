#!/bin/bash
# Toeplitz input block: TMA overlapping-window probe, its unit test, the network-level suites, bench with the per-layer table
export PYTHONUNBUFFERED=1
O=gpurun_out/r2i
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/probes/tma_overlap_probe.cu -lcuda > ${O}_probe.log 2>&1 && timeout 60 /tmp/tma_probe >> ${O}_probe.log 2>&1; cat ${O}_probe.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cin1" 2>&1 | tail -15 > ${O}_pytest_cin1.log; cat ${O}_pytest_cin1.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > ${O}_pytest.log; cat ${O}_pytest.log
timeout 400 python bench.py --layers --no-train --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err; cut -c1-1500 ${O}_bench.json; grep -E "in_block|KIND" ${O}_bench.err
SEG3D_FUSE_IN=0 timeout 400 python bench.py --layers --no-train --no-cpu-baseline > ${O}_bench_nofuse.json 2> ${O}_bench_nofuse.err; cut -c1-300 ${O}_bench_nofuse.json; grep -E "in_block|KIND" ${O}_bench_nofuse.err
