"""Oracle: state-dict schema and seeded initialisation of VNet / VBNet.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, without any nn.Module:
  * the parameter schema of segmentation3d/network/vnet.py:23-34 and vbnet.py:24-35
    (names, shapes, registration order);
  * the random stream of `torch.manual_seed(s); net = SegmentationNet(i, o);
    parameters_kaiming_init(net)` (network/vnet.py:10-12, module/weight_init.py:4-17):
    every nn.Conv3d / nn.ConvTranspose3d constructor draws kaiming_uniform_(a=sqrt(5)) weights and
    a uniform bias in registration order, GroupNorm draws nothing (weight=1, bias=0), and
    `net.apply(kaiming_weight_init)` then overwrites conv weights with kaiming_normal_ and
    zeroes conv biases in the same (depth-first registration) order.
`tests/golden/weights_sha256.json` pins the result against the real reference modules.
"""
import math

import torch
import torch.nn as nn


def _rblock_specs(prefix, ch, n, bott, ratio=4):
    specs = []
    for i in range(n):
        op = '%s.ops.%d' % (prefix, i)
        if not bott:
            specs.append((op + '.conv', 'conv', (ch, ch, 3, 3, 3)))
            specs.append((op + '.gn', 'gn', (ch,)))
        else:
            mid = ch // ratio
            for j, (ci, co) in enumerate([(ch, mid), (mid, mid), (mid, ch)]):
                specs.append(('%s.conv%d.conv' % (op, j + 1), 'conv', (co, ci, 3, 3, 3)))
                specs.append(('%s.conv%d.gn' % (op, j + 1), 'gn', (co,)))
    return specs


def layer_specs(arch, in_channels, out_channels):
    """Ordered [(name, kind, weight_shape)], kind in {'conv','convT','gn'}.

    vnet.py:23-34 / vbnet.py:24-35: VBNet sets compression=True on down_64, down_128,
    down_256, up_256, up_128."""
    assert arch in ('vnet', 'vbnet')
    vb = arch == 'vbnet'
    s = [('in_block.conv', 'conv', (16, in_channels, 3, 3, 3)), ('in_block.gn', 'gn', (16,))]
    for cin, n, bott in [(16, 1, False), (32, 2, vb), (64, 3, vb), (128, 3, vb)]:
        cout = 2 * cin
        p = 'down_%d' % cout
        s.append((p + '.down_conv', 'conv', (cout, cin, 2, 2, 2)))
        s.append((p + '.down_gn', 'gn', (cout,)))
        s += _rblock_specs(p + '.rblock', cout, n, bott)
    for cin, cout, n, bott in [(256, 256, 3, vb), (256, 128, 3, vb), (128, 64, 2, False), (64, 32, 1, False)]:
        p = 'up_%d' % cout
        s.append((p + '.up_conv', 'convT', (cin, cout // 2, 2, 2, 2)))  # ConvTranspose3d: [Cin, Cout, k,k,k]
        s.append((p + '.up_gn', 'gn', (cout // 2,)))
        s += _rblock_specs(p + '.rblock', cout, n, bott)
    s.append(('out_block.conv1', 'conv', (out_channels, 32, 3, 3, 3)))
    s.append(('out_block.gn1', 'gn', (out_channels,)))
    s.append(('out_block.conv2', 'conv', (out_channels, out_channels, 1, 1, 1)))
    s.append(('out_block.gn2', 'gn', (out_channels,)))
    return s


def state_dict_schema(arch, in_channels, out_channels):
    """Ordered [(key, shape)] exactly as `net.state_dict()` lists them."""
    out = []
    for name, kind, shape in layer_specs(arch, in_channels, out_channels):
        out.append((name + '.weight', tuple(shape)))
        nb = shape[1] if kind == 'convT' else shape[0]
        out.append((name + '.bias', (nb,)))
    return out


def init_state_dict(arch, in_channels, out_channels, seed, mode='kaiming'):
    """Bit-identical to: torch.manual_seed(seed); net = SegmentationNet(in, out);
    parameters_kaiming_init(net) (or parameters_gaussian_init); net.state_dict()."""
    specs = layer_specs(arch, in_channels, out_channels)
    torch.manual_seed(seed)
    sd = {}
    # 1) constructors (torch.nn.modules.conv._ConvNd.reset_parameters)
    for name, kind, shape in specs:
        if kind == 'gn':
            sd[name + '.weight'] = torch.ones(shape)
            sd[name + '.bias'] = torch.zeros(shape)
            continue
        w = torch.empty(shape)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        nb = shape[1] if kind == 'convT' else shape[0]
        b = torch.empty(nb)
        fan_in, _ = nn.init._calculate_fan_in_and_fan_out(w)
        if fan_in != 0:
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(b, -bound, bound)
        sd[name + '.weight'], sd[name + '.bias'] = w, b
    # 2) net.apply(<init>) in depth-first registration order (weight_init.py:4-17 / 20-29)
    for name, kind, shape in specs:
        if kind == 'gn':
            continue  # class name 'GroupNorm' matches none of Conv3d/ConvTranspose3d/BatchNorm/Linear
        if mode == 'kaiming':
            nn.init.kaiming_normal_(sd[name + '.weight'])
        elif mode == 'gaussian':
            sd[name + '.weight'].normal_(0, 0.01)
        else:
            raise ValueError(mode)
        sd[name + '.bias'].zero_()
    return sd


def randomize_affine(sd, seed):
    """Test helper (not reference behaviour): give GN gamma/beta and conv biases non-trivial
    values so parity tests exercise the affine and bias paths, which init leaves at 1/0/0."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        if v.dim() == 1:
            if '.gn' in k or '_gn' in k:
                if k.endswith('.weight'):
                    out[k] = 1.0 + 0.2 * torch.randn(v.shape, generator=g)
                else:
                    out[k] = 0.1 * torch.randn(v.shape, generator=g)
            else:
                out[k] = 0.05 * torch.randn(v.shape, generator=g)
        else:
            out[k] = v.clone()
    return out
