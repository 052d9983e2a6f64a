"""Forward of B patches in a loop under ONE event pair vs the sum of per-launch event pairs: how much of a forward is spent between
kernels (launch gaps, ramp-up, clocks under sustained load)?   python tools/forward_gap.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'medical-segmentation3d-toolkit_b200'))
sys.path.insert(0, ROOT)
import torch
from segmentation3d._b200.plan import NetPlan
from oracle import init as oinit

B = int(sys.argv[1]) if len(sys.argv) > 1 else 36
sd = oinit.init_state_dict('vnet', 1, 2, 0)
plan = NetPlan(sd, mode='fp16', device='cuda')
ws, ops = plan.plan(B, 96, 96, 96)
x = torch.randn((B, 1, 96, 96, 96), device='cuda')
plan.load_input(ws, x)
for graph in (False, True):
    plan.use_graph = graph
    ws.pop('graph', None)
    for _ in range(3):
        plan.run(ws, ops)
    torch.cuda.synchronize()
    n = 20
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        plan.run(ws, ops)
    b.record()
    torch.cuda.synchronize()
    print('B=%d %s: %.3f ms per forward over %d back-to-back forwards (%d launches each)' % (B, 'graph' if graph else 'eager', a.elapsed_time(b) / n, n, plan.launches(ws)))
plan.use_graph = False
tot = sum(ms for _, ms in plan.run_profiled(ws, ops))
tot2 = sum(ms for _, ms in plan.run_profiled(ws, ops))
print('sum of per-launch event pairs: %.3f / %.3f ms' % (tot, tot2))
