"""Time one tensor-core wgrad launch. usage: python tools/bench_wgrad.py Cin Cout N D"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'medical-segmentation3d-toolkit_b200'))
import torch
from segmentation3d._b200 import lib as L

L.load()
cin, cout, n, d = [int(v) for v in sys.argv[1:5]]
x = torch.randn((n, d, d, d, cin), device='cuda').bfloat16()
dy = torch.randn((n, d, d, d, cout), device='cuda').bfloat16()
dw = torch.zeros((27 * cin * cout,), device='cuda')


def go():
    L.call('seg3d_conv3d_wgrad', L.CONV_K3, L.BF16, L.ptr(x), cin, cin, L.ptr(dy), cout, cout, L.ptr(dw), n, d, d, d, L.stream_ptr())


for _ in range(2):
    go()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    go()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print('wgrad %dx%d N=%d D=%d: %.3f ms %.0f TFLOP/s' % (cin, cout, n, d, ms, 2.0 * n * d ** 3 * 27 * cin * cout / ms / 1e9))
