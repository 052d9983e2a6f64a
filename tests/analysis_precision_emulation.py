"""CPU emulation of where half-precision rounding enters the VBNet / VNet forward (analysis tool, not a test).

TEST INFRASTRUCTURE: imports oracle/.  Restates the plan's rounding points on the CPU so that precision placements can
be compared without GPU time (same method as SURVEY A.6):
  W  conv weights rounded to the operand type
  O  raw conv output rounded on store (GroupNorm statistics are taken BEFORE the rounding, as in the kernels)
  G  GroupNorm(+ReLU/+residual) output rounded on store
out_block.conv1's raw output stays fp32 (SEG3D_OUT_F32) and the tail is fp32, as in the plan.

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
from oracle import init as oinit, net as onet            # noqa: E402
from oracle.metrics import parity_report                 # noqa: E402

EPS = 1e-5


class Emu(object):
    def __init__(self, sd, dtype=torch.float16, exact=(), split=()):
        self.sd, self.dt, self.exact, self.split = sd, dtype, tuple(exact), tuple(split)

    def _is(self, name, group):
        return any(name.startswith(p) for p in group)

    def r(self, t, name):
        if self._is(name, self.exact) or self._is(name, self.split):
            return t
        return t.to(self.dt).float()

    def rw(self, w, name):
        if self._is(name, self.exact) or self._is(name, self.split):
            return w
        return w.to(self.dt).float()

    def gn(self, y, name, store_name, relu=True, res=None, raw_f32=False):
        # statistics from the unrounded accumulators; the normalisation reads the STORED (rounded) raw tensor
        mean = y.mean(dim=(1, 2, 3, 4), keepdim=True)
        var = y.var(dim=(1, 2, 3, 4), unbiased=False, keepdim=True)
        ys = y if raw_f32 else self.r(y, store_name)
        g = self.sd[name + '.weight'].view(1, -1, 1, 1, 1)
        b = self.sd[name + '.bias'].view(1, -1, 1, 1, 1)
        z = (ys - mean) * torch.rsqrt(var + EPS) * g + b
        if res is not None:
            z = z + res
        if relu:
            z = F.relu(z)
        return z

    def conv_gn(self, x, name, act, res=None):
        y = F.conv3d(x, self.rw(self.sd[name + '.conv.weight'], name), self.sd[name + '.conv.bias'], padding=1)
        return self.r(self.gn(y, name + '.gn', name, relu=act or res is not None, res=res), name)

    def rblock(self, x, name):
        sd = self.sd
        n = 0
        while (name + '.ops.%d.conv.weight' % n) in sd or (name + '.ops.%d.conv1.conv.weight' % n) in sd:
            n += 1
        y = x
        for i in range(n):
            last = i == n - 1
            op = name + '.ops.%d' % i
            if (op + '.conv.weight') in sd:
                y = self.conv_gn(y, op, not last, res=x if last else None)
            else:
                y = self.conv_gn(y, op + '.conv1', True)
                y = self.conv_gn(y, op + '.conv2', True)
                y = self.conv_gn(y, op + '.conv3', not last, res=x if last else None)
        return y

    def down(self, x, name):
        y = F.conv3d(x, self.rw(self.sd[name + '.down_conv.weight'], name + '.down'), self.sd[name + '.down_conv.bias'], stride=2)
        y = self.r(self.gn(y, name + '.down_gn', name + '.down'), name + '.down')
        return self.rblock(y, name + '.rblock')

    def up(self, x, skip, name):
        y = F.conv_transpose3d(x, self.rw(self.sd[name + '.up_conv.weight'], name + '.up'), self.sd[name + '.up_conv.bias'], stride=2)
        y = self.r(self.gn(y, name + '.up_gn', name + '.up'), name + '.up')
        return self.rblock(torch.cat((y, skip), 1), name + '.rblock')

    def forward(self, x):
        sd = self.sd
        with torch.no_grad():
            # the input block keeps fp32-accurate weights (hi/lo split); its input patch is stored in the operand type
            xin = self.r(x.float(), 'in_block')
            y = F.conv3d(xin, sd['in_block.conv.weight'], sd['in_block.conv.bias'], padding=1)
            o16 = self.r(self.gn(y, 'in_block.gn', 'in_block'), 'in_block')
            o32 = self.down(o16, 'down_32')
            o64 = self.down(o32, 'down_64')
            o128 = self.down(o64, 'down_128')
            o256 = self.down(o128, 'down_256')
            o = self.up(o256, o128, 'up_256')
            o = self.up(o, o64, 'up_128')
            o = self.up(o, o32, 'up_64')
            o = self.up(o, o16, 'up_32')
            y = F.conv3d(o, self.rw(sd['out_block.conv1.weight'], 'out_block'), sd['out_block.conv1.bias'], padding=1)
            y = self.gn(y, 'out_block.gn1', 'out_block', raw_f32=True)
            y = F.conv3d(y, sd['out_block.conv2.weight'], sd['out_block.conv2.bias'])
            y = F.group_norm(y, 1, sd['out_block.gn2.weight'], sd['out_block.gn2.bias'], EPS)
            return F.softmax(y, 1)


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
        y = Emu(sd, torch.float16, exact=ex).forward(x)
        rep = parity_report(ref[0].numpy(), y[0].numpy())
        print('exact=%-60s max %.2e agree %.5f dice %s' % (','.join(ex) or '-', rep['max_abs'], rep['agree'],
                                                          ' '.join('%.4f' % d for d in rep['dice'])), flush=True)


if __name__ == '__main__':
    main()
