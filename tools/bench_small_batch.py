"""Forward latency of VNet on B patches of 96^3 (fp16), eager launches vs CUDA-graph replay (SEG3D_GRAPH)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'medical-segmentation3d-toolkit_b200'))
sys.path.insert(0, ROOT)
import torch
from segmentation3d._b200.plan import NetPlan
from oracle import init as oinit

sd = oinit.init_state_dict('vnet', 1, 2, 0)
for B in (1, 2, 4):
    x = torch.randn((B, 1, 96, 96, 96), device='cuda')
    for graph in (False, True):
        plan = NetPlan(sd, mode='fp16', device='cuda')
        plan.use_graph = graph
        for _ in range(3):
            plan.forward(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 50
        for _ in range(n):
            plan.forward(x)
        torch.cuda.synchronize()
        print('B=%d %-6s %.3f ms per forward (wall clock, %d launches)' % (B, 'graph' if graph else 'eager', (time.perf_counter() - t0) / n * 1e3,
                                                                            plan.launches_per_forward))
