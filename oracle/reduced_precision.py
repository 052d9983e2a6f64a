"""Oracle: the VNet / VBNet program of oracle/net.py with the PRODUCT's half-precision rounding points injected.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  torch CPU fp32 arithmetic; a value is "stored" in the reduced type by
a round trip `x.to(dtype).float()`.  This is not a restatement of the reference (which is fp32 throughout): it restates
where the B200 plan (medical-segmentation3d-toolkit_b200/segmentation3d/_b200/plan.py) rounds, so that

  * precision placements can be compared on the CPU (tests/analysis_precision_emulation.py, method of SURVEY.md A.6), and
  * the bf16 TRAINING path can be checked against an autograd that sees the same forward: with bf16 storage ~1 % of the
    ReLU masks of the deep layers differ from the fp32 forward, so the fp32 oracle's gradients differ from any bf16
    implementation's by ~20 % in L2 there; the gradients of THIS program are the ones the kernels must reproduce.

Rounding points (plan.py `_build`):
  W  tensor-core operand weights (the input block keeps fp32-accurate weights through its hi/lo split);
  O  each convolution's raw output on store; the GroupNorm statistics come from the fp32 accumulators BEFORE it;
     out_block.conv1's raw output stays fp32 (SEG3D_OUT_F32) and the tail (GN1, 1x1x1 conv, GN2, softmax) is fp32;
  G  each GroupNorm(+residual)(+ReLU) output on store; the network input patch is stored in the reduced type too.
Under autograd the round trip also rounds the gradient flowing back through O and G (the kernels store dy and the data
gradients in the reduced type); weights are rounded straight-through, as the parameter gradients are fp32.
"""
import torch
import torch.nn.functional as F

from .net import GN_EPS, strip_module_prefix


class ReducedPrecisionNet(object):
    def __init__(self, state_dict, dtype=torch.float16, exact=()):
        """exact: name prefixes ('up_32', 'down_64.rblock', ...) whose rounding points are switched off."""
        self.sd, self.dt, self.exact = strip_module_prefix(state_dict), dtype, tuple(exact)

    def _exact(self, name):
        return any(name.startswith(p) for p in self.exact)

    def r(self, t, name):
        return t if self._exact(name) else t.to(self.dt).float()

    def rw(self, w, name):
        if self._exact(name):
            return w
        return w + (w.to(self.dt).float() - w).detach()

    def gn(self, y, name, store_name, relu=True, res=None, raw_f32=False):
        mean = y.mean(dim=(1, 2, 3, 4), keepdim=True)
        var = y.var(dim=(1, 2, 3, 4), unbiased=False, keepdim=True)
        ys = y if raw_f32 else self.r(y, store_name)
        g = self.sd[name + '.weight'].view(1, -1, 1, 1, 1)
        b = self.sd[name + '.bias'].view(1, -1, 1, 1, 1)
        z = (ys - mean) * torch.rsqrt(var + GN_EPS) * g + b
        if res is not None:
            z = z + res
        return F.relu(z) if relu else z

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
        y = F.group_norm(y, 1, sd['out_block.gn2.weight'], sd['out_block.gn2.bias'], GN_EPS)
        return F.softmax(y, 1)


def forward(state_dict, x, dtype=torch.float16, exact=()):
    """probabilities [B,C,D,H,W] of the rounded program, no autograd."""
    with torch.no_grad():
        return ReducedPrecisionNet(state_dict, dtype, exact).forward(x)


def forward_with_grad(params, x, dtype=torch.bfloat16):
    """same program with autograd enabled (params: dict name -> leaf tensor)."""
    return ReducedPrecisionNet(params, dtype).forward(x)
