#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2t
mkdir -p gpurun_out
for cfg in "default:" "c1:SEG3D_ZM_CTAS_PER_SM=1" "c1r12:SEG3D_ZM_CTAS_PER_SM=1 SEG3D_ZM_RING=11" "c2r4:SEG3D_ZM_RING=4" "lseg24:SEG3D_ZM_LSEG=24" "lseg48:SEG3D_ZM_LSEG=48" "c1lseg32:SEG3D_ZM_CTAS_PER_SM=1 SEG3D_ZM_LSEG=32"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 400 python bench.py --layers --no-train --no-cpu-baseline > "${O}_$name.json" 2> "${O}_$name.err"; python -c "
import json; d=json.load(open('${O}_$name.json')); print('$name', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2))"; grep -E "up_32.rblock.ops.0.conv|down_32.rblock.ops.0.conv" "${O}_$name.err"
done
