#!/bin/bash
# batch-size and CUDA-graph sweep of the headline inference step (device-resident value and e2e)
export PYTHONUNBUFFERED=1
O=gpurun_out/r2g
mkdir -p gpurun_out
for G in 0 1; do
 for B in 20 30 36 45 60; do
  SEG3D_GRAPH=$G timeout 200 python bench.py --batch $B --no-train --no-cpu-baseline > ${O}_g${G}_b${B}.json 2> ${O}_g${G}_b${B}.err
  python - <<PY
import json
d=json.load(open('${O}_g${G}_b${B}.json'))
print('GRAPH=$G batch=$B value %.1f ms %.2f e2e %.1f launches %d k3 %.1f TF/s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['achieved']))
PY
 done
done | tee ${O}_sweep.txt
