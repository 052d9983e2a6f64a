"""CPU emulation of where half-precision rounding enters the VBNet / VNet forward (analysis tool, not a test).

TEST INFRASTRUCTURE: imports oracle/.  Runs oracle/reduced_precision.py (the oracle program with the plan's rounding
points injected) so that precision placements can be compared without GPU time (same method as SURVEY A.6).

    python tests/analysis_precision_emulation.py [vbnet|vnet] [classes] [size]

Each row switches rounding OFF ("exact") for one group of layers and reports the parity bars, so the layers whose
operand rounding costs the rare-class Dice stand out.
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import init as oinit, net as onet, reduced_precision as orp   # noqa: E402
from oracle.metrics import parity_report                 # noqa: E402

def seeded_input(seed, shape, kind='smooth'):
    # same generator as tests/test_gpu_kernels.py::seeded_input
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g)
    if kind == 'smooth':
        lo = torch.randn((shape[0], shape[1]) + tuple(max(2, s // 8) for s in shape[2:]), generator=g)
        x = F.interpolate(lo, size=shape[2:], mode='trilinear', align_corners=False) * 2 + 0.3 * x
    return x


def main():
    arch = sys.argv[1] if len(sys.argv) > 1 else 'vbnet'
    cout = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    size = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    torch.set_num_threads(os.cpu_count())
    sd = oinit.init_state_dict(arch, 1, cout, 0)
    x = seeded_input(7, (1, 1, size, size, size), 'smooth')
    ref = onet.forward(sd, x)
    groups = [(), ('in_block',), ('down_32',), ('down_64',), ('down_128',), ('down_256',), ('up_256',), ('up_128',),
              ('up_64',), ('up_32',), ('out_block',), ('in_block', 'down_32', 'up_32', 'out_block'),
              ('down_128', 'down_256', 'up_256', 'up_128'), ('down_64', 'down_128', 'down_256', 'up_256', 'up_128', 'up_64')]
    for ex in groups:
        y = orp.forward(sd, x, torch.float16, exact=ex)
        rep = parity_report(ref[0].numpy(), y[0].numpy())
        print('exact=%-60s max %.2e agree %.5f dice %s' % (','.join(ex) or '-', rep['max_abs'], rep['agree'],
                                                          ' '.join('%.4f' % d for d in rep['dice'])), flush=True)


if __name__ == '__main__':
    main()
