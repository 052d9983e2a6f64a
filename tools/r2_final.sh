#!/bin/bash
# Round-2 final single-GPU measurement: full -m gpu suite, headline bench (train + parity + cpu baseline), VBNet and fp32x lines,
# ncu launch lists (inference volume, training step) and one --set full capture of the tensor-core launches of a forward.
export PYTHONUNBUFFERED=1
O=gpurun_out/r2z
mkdir -p gpurun_out
nvidia-smi -L > ${O}_gpus.txt
timeout 900 python -m pytest tests -m gpu -q -rs 2>&1 | tail -12 > ${O}_pytest.log; cat ${O}_pytest.log
timeout 600 python bench.py --layers > ${O}_bench.json 2> ${O}_bench.err; cut -c1-1200 ${O}_bench.json; grep KIND ${O}_bench.err
timeout 300 python bench.py --arch vbnet --classes 5 --mode auto --no-train > ${O}_bench_vbnet.json 2> ${O}_bench_vbnet.err; cut -c1-300 ${O}_bench_vbnet.json
timeout 300 python bench.py --arch vbnet --classes 5 --mode fp16 --no-train --no-cpu-baseline > ${O}_bench_vbnet_fp16.json 2> ${O}_bench_vbnet_fp16.err; cut -c1-300 ${O}_bench_vbnet_fp16.json
timeout 300 python bench.py --mode fp32x --no-train --layers > ${O}_bench_fp32x.json 2> ${O}_bench_fp32x.err; cut -c1-300 ${O}_bench_fp32x.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > ${O}_bench_ref.json 2> ${O}_bench_ref.err; cut -c1-400 ${O}_bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file ${O}_infer_launches.csv python bench.py --steps 1 --warmup 3 --no-train --no-cpu-baseline > ${O}_infer_ncu.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_train_launches.csv python tools/train_one_step.py bf16 8 3 > ${O}_train_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"zmarch|persistent|fold|cin1" -s 29 -c 29 -o ${O}_fwd_full python tools/profile_forward.py 36 fp16 > ${O}_fwd_full.log 2>&1
ncu -i ${O}_fwd_full.ncu-rep --page raw --csv > ${O}_fwd_full_raw.csv 2>/dev/null; rm -f ${O}_fwd_full.ncu-rep
ls -la gpurun_out | tail -20
