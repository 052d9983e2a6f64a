#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2p
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "up_block_gn_folded" 2>&1 | tail -15 > ${O}_pytest_up.log; cat ${O}_pytest_up.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > ${O}_pytest.log; cat ${O}_pytest.log
timeout 400 python bench.py --layers --no-train > ${O}_bench.json 2> ${O}_bench.err; python -c "
import json; d=json.load(open('${O}_bench.json')); print('default', d['value'], d['e2e']['value'], d['ms_per_step'], d['parity'])"; grep -E "KIND|up_32|out_block" ${O}_bench.err
SEG3D_FUSE_UP=0 timeout 400 python bench.py --layers --no-train --no-cpu-baseline > ${O}_bench_nofuse.json 2> ${O}_bench_nofuse.err; python -c "
import json; d=json.load(open('${O}_bench_nofuse.json')); print('FUSE_UP=0', d['value'], d['e2e']['value'], d['ms_per_step'])"; grep -E "KIND|up_32|out_block" ${O}_bench_nofuse.err
