#!/bin/bash
# Round-2 GPU call (one B200): full -m gpu suite, headline bench (train + parity records), VBNet bench in its default mode.
export PYTHONUNBUFFERED=1
O=gpurun_out/r2b
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs 2>&1 | tail -40 > ${O}_pytest.log; tail -25 ${O}_pytest.log
timeout 400 python bench.py --layers > ${O}_bench.json 2> ${O}_bench.err; cut -c1-3000 ${O}_bench.json; tail -12 ${O}_bench.err
timeout 300 python bench.py --arch vbnet --classes 5 --mode auto --no-train --no-cpu-baseline > ${O}_bench_vbnet.json 2> ${O}_bench_vbnet.err; cut -c1-400 ${O}_bench_vbnet.json; tail -3 ${O}_bench_vbnet.err
