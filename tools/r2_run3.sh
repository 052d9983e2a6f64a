#!/bin/bash
# sub-batched schedule sweep (L2 residency of the level 0/1 tensors) + full suite
export PYTHONUNBUFFERED=1
O=gpurun_out/r2c
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > ${O}_pytest.log; tail -8 ${O}_pytest.log
for MB in 0 60 120 240; do
 for G in 0 1; do
  SEG3D_SUBBATCH_MB=$MB SEG3D_GRAPH=$G timeout 200 python bench.py --no-train --no-cpu-baseline --layers > ${O}_sb${MB}_g${G}.json 2> ${O}_sb${MB}_g${G}.err
  python - <<PY
import json
d=json.load(open('${O}_sb${MB}_g${G}.json'))
print('SUBBATCH_MB=$MB GRAPH=$G value %.1f e2e %.1f launches %d shares %s' % (d['value'], d['e2e']['value'], d['gpu_launches'], d['kernel_shares']))
PY
  grep KIND ${O}_sb${MB}_g${G}.err | head -8
 done
done
