#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2w
mkdir -p gpurun_out
for cfg in "default(partial rows: none):" "promo128:SEG3D_TC_L2PROMO=2" "promo64:SEG3D_TC_L2PROMO=1" "promo0_all:SEG3D_TC_L2PROMO=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 400 python bench.py --layers --no-train --no-cpu-baseline > "${O}_$name.json" 2> "${O}_$name.err"; python -c "
import json; d=json.load(open('${O}_$name.json')); print('$name', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2))"; grep -E "down_32.down_conv|down_64.down_conv|KIND conv_tc_k2s2|KIND conv_tc_k3|KIND conv_tc_t2s2" "${O}_$name.err"
done
timeout 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:persistent -c 1 python tools/profile_forward.py 36 fp16 2>&1 | grep -E "dram__|gpu__time|persistent" | head -8
