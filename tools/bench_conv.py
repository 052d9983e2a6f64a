"""Time single convolution launches through the C ABI (CUDA events, L2 flushed by rotating buffers).
usage: python tools/bench_conv.py [Cin Cout N D [mode] [impl]] ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'medical-segmentation3d-toolkit_b200'))
import torch
from segmentation3d._b200 import lib as L

L.load()
CASES = [(32, 32, 6, 96), (32, 16, 6, 96), (64, 64, 6, 48), (128, 128, 6, 24), (256, 256, 6, 12), (256, 256, 6, 6), (32, 32, 6, 48)]


def run(cin, cout, n, d, reps=10):
    x = [torch.randn((n, d, d, d, cin), device='cuda').half() for _ in range(3)]
    w = (torch.randn((27, cout, cin), device='cuda') * 0.05).half()
    y = torch.empty((n, d, d, d, cout), device='cuda', dtype=torch.half)
    stats = torch.zeros((n, 2), dtype=torch.float64, device='cuda')
    def go(i):
        L.call('seg3d_conv3d_fwd', L.CONV_K3, L.F16, L.IMPL_TCGEN05, L.ptr(x[i % 3]), cin, cin, L.ptr(w), None, L.ptr(y), cout, cout,
               n, d, d, d, L.ptr(stats), L.stream_ptr())
    for i in range(3):
        go(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        go(i)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    fl = 2.0 * n * d ** 3 * 27 * cin * cout
    return ms, fl / ms / 1e9


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'one':
    c = tuple(int(v) for v in sys.argv[2:6])
    ms, tf = run(*c, reps=3)
    print('%dx%d N=%d D=%d: %.3f ms %.0f TFLOP/s' % (c[0], c[1], c[2], c[3], ms, tf))
    sys.exit(0)

if __name__ == '__main__':
    for kw in ('0', '1'):
        for st in ('0', '3', '4', '6'):
            os.environ['SEG3D_TC_KWFUSE'] = kw
            os.environ['SEG3D_TC_STAGES'] = st
            out = []
            for c in CASES:
                ms, tf = run(*c)
                out.append('%dx%d@%d: %.3fms %.0fTF' % (c[0], c[1], c[3], ms, tf))
            print('kwfuse=%s stages=%s | ' % (kw, st) + ' | '.join(out), flush=True)
