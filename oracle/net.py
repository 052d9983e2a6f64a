"""Oracle: VNet / VBNet forward as a flat functional program over a state dict.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  torch CPU fp32, NCDHW, the same
ATen ops the reference's nn.Modules dispatch to, in the same order.

Reference call graph restated here (all paths relative to /root/reference):
  segmentation3d/network/vnet.py:36-48          SegmentationNet.forward (VNet)
  segmentation3d/network/vbnet.py:38-50         SegmentationNet.forward (VBNet, same topology)
  segmentation3d/network/module/vnet_inblock.py:13-15    conv k3 p1 -> GN(1,C) -> ReLU
  segmentation3d/network/module/vnet_downblock.py:19-22  conv k2 s2 -> GN -> ReLU -> rblock
  segmentation3d/network/module/vnet_upblock.py:19-23    convT k2 s2 -> GN -> ReLU -> cat((up, skip),1) -> rblock
  segmentation3d/network/module/vnet_outblock.py:20-24   conv k3 -> GN -> ReLU -> conv k1 -> GN -> softmax(dim=1)
  segmentation3d/network/module/residual_block3.py:21-26,44-46   relu(x + ops(x))
  segmentation3d/network/module/conv_gn_relu3.py:16-20,32-34     conv -> GN [-> ReLU]; bottleneck = 3 of them
"""
import torch
import torch.nn.functional as F

GN_EPS = 1e-5  # nn.GroupNorm default, conv_gn_relu3.py:11


def strip_module_prefix(state_dict):
    """core/seg_infer.py:130-142: drop the DataParallel 'module.' prefix if present."""
    if any(k.startswith('module.') for k in state_dict):
        return {k[7:]: v for k, v in state_dict.items()}
    return dict(state_dict)


def _gn(x, sd, name):
    # nn.GroupNorm(1, C): per-sample stats over C*D*H*W, biased variance, per-channel affine
    return F.group_norm(x, 1, sd[name + '.weight'], sd[name + '.bias'], GN_EPS)


def _conv_gn(x, sd, name, act):
    # conv_gn_relu3.py:16-20 (k3, stride 1, pad 1 everywhere it is instantiated)
    x = F.conv3d(x, sd[name + '.conv.weight'], sd[name + '.conv.bias'], stride=1, padding=1)
    x = _gn(x, sd, name + '.gn')
    return F.relu(x) if act else x


def _rblock(x, sd, name):
    # residual_block3.py: ops.<i> are ConvGnRelu3 (plain) or BottConvGnRelu3 (conv1/conv2/conv3)
    n = 0
    while (name + '.ops.%d.conv.weight' % n) in sd or (name + '.ops.%d.conv1.conv.weight' % n) in sd:
        n += 1
    y = x
    for i in range(n):
        last = (i == n - 1)
        op = name + '.ops.%d' % i
        if (op + '.conv.weight') in sd:
            y = _conv_gn(y, sd, op, act=not last)
        else:  # conv_gn_relu3.py:26-34
            y = _conv_gn(y, sd, op + '.conv1', act=True)
            y = _conv_gn(y, sd, op + '.conv2', act=True)
            y = _conv_gn(y, sd, op + '.conv3', act=not last)
    return F.relu(x + y)


def _down(x, sd, name):
    x = F.conv3d(x, sd[name + '.down_conv.weight'], sd[name + '.down_conv.bias'], stride=2)
    x = F.relu(_gn(x, sd, name + '.down_gn'))
    return _rblock(x, sd, name + '.rblock')


def _up(x, skip, sd, name):
    x = F.conv_transpose3d(x, sd[name + '.up_conv.weight'], sd[name + '.up_conv.bias'], stride=2)
    x = F.relu(_gn(x, sd, name + '.up_gn'))
    x = torch.cat((x, skip), 1)  # vnet_upblock.py:21: up-conv output first, skip second
    return _rblock(x, sd, name + '.rblock')


def forward(state_dict, x, return_logits=False):
    """probabilities [B,C,D,H,W] float32 for input [B,Cin,D,H,W] float32 (CPU)."""
    sd = strip_module_prefix(state_dict)
    with torch.no_grad():
        x = x.float()
        out16 = F.relu(_gn(F.conv3d(x, sd['in_block.conv.weight'], sd['in_block.conv.bias'], padding=1),
                           sd, 'in_block.gn'))
        out32 = _down(out16, sd, 'down_32')
        out64 = _down(out32, sd, 'down_64')
        out128 = _down(out64, sd, 'down_128')
        out256 = _down(out128, sd, 'down_256')
        out = _up(out256, out128, sd, 'up_256')
        out = _up(out, out64, sd, 'up_128')
        out = _up(out, out32, sd, 'up_64')
        out = _up(out, out16, sd, 'up_32')
        # vnet_outblock.py:20-24
        out = F.conv3d(out, sd['out_block.conv1.weight'], sd['out_block.conv1.bias'], padding=1)
        out = F.relu(_gn(out, sd, 'out_block.gn1'))
        out = F.conv3d(out, sd['out_block.conv2.weight'], sd['out_block.conv2.bias'])
        out = _gn(out, sd, 'out_block.gn2')
        if return_logits:
            return out
        return F.softmax(out, dim=1)


def forward_with_grad(params, x):
    """Same program with autograd enabled (params: dict name -> leaf tensor). Used by the
    training-step oracle (core/seg_train.py:119-127)."""
    sd = params
    out16 = F.relu(_gn(F.conv3d(x, sd['in_block.conv.weight'], sd['in_block.conv.bias'], padding=1),
                       sd, 'in_block.gn'))
    out32 = _down(out16, sd, 'down_32')
    out64 = _down(out32, sd, 'down_64')
    out128 = _down(out64, sd, 'down_128')
    out256 = _down(out128, sd, 'down_256')
    out = _up(out256, out128, sd, 'up_256')
    out = _up(out, out64, sd, 'up_128')
    out = _up(out, out32, sd, 'up_64')
    out = _up(out, out16, sd, 'up_32')
    out = F.conv3d(out, sd['out_block.conv1.weight'], sd['out_block.conv1.bias'], padding=1)
    out = F.relu(_gn(out, sd, 'out_block.gn1'))
    out = F.conv3d(out, sd['out_block.conv2.weight'], sd['out_block.conv2.bias'])
    out = _gn(out, sd, 'out_block.gn2')
    return F.softmax(out, dim=1)


MAX_STRIDE = 16  # network/vnet.py:50-51
