"""CPU check of the HOST WIRING of the standalone network modules (segmentation3d/network/module/*.py).

The modules' forward is a sequence of C-ABI calls (segmentation3d/_b200/blocks.py).  Here every entry point they use is
replaced by a torch emulation of its documented semantics (include/seg3d_b200.h: argument order, NDHWC pitches, packed
weight layouts, GroupNorm sums, in-place concat, residual and ReLU flags), the modules run on CPU tensors, and the result
is compared with the oracle's functional restatement of the reference blocks.  This pins argument order, layout
conversions and block topology without a GPU; the kernels themselves are pinned by the `-m gpu` tests.
"""
import contextlib

import pytest
import torch
import torch.nn.functional as F

from oracle import net as onet

EPS = 1e-5


class _P(object):
    """stand-in for a device pointer: a tensor and an element offset"""

    def __init__(self, t, off):
        self.t, self.off = t, off


def _rows(p, n_rows, ld, C):
    """[n_rows, C] window (a view) of the flat buffer behind pointer p with pitch ld"""
    flat = p.t.reshape(-1)
    return flat[p.off:p.off + n_rows * ld].view(n_rows, ld)[:, :C] if p.off + n_rows * ld <= flat.numel() \
        else torch.as_strided(flat, (n_rows, C), (ld, 1), p.off)


def _install(monkeypatch):
    from segmentation3d._b200 import blocks, lib
    calls = []

    def ptr(t, off=0):
        return None if t is None else _P(t, off)

    def conv3d_fwd(mode, dtype, impl, x, x_ld, Cin, w, bias, y, y_ld, Cout, N, D, H, W, stats, stream):
        out_f32 = bool(dtype & lib.OUT_F32)
        dtype &= 0xff
        xs = _rows(x, N * D * H * W, x_ld, Cin).float().view(N, D, H, W, Cin).permute(0, 4, 1, 2, 3)
        simt = impl == lib.IMPL_SIMT or (impl == lib.IMPL_AUTO and Cin == 1)
        wp = w.t.float()
        k = {lib.CONV_K3: 3, lib.CONV_K2S2: 2, lib.CONV_T2S2: 2, lib.CONV_K1: 1}[mode]
        if mode == lib.CONV_T2S2:
            wt = wp.view(Cin, 2, 2, 2, Cout) if simt else wp.view(2, 2, 2, Cout, Cin).permute(4, 0, 1, 2, 3)
            wt = wt.permute(0, 4, 1, 2, 3)                                             # IODHW
            r = F.conv_transpose3d(xs, wt, None if bias is None else bias.t.float(), stride=2)
        else:
            wt = wp.view(k, k, k, Cin, Cout).permute(4, 3, 0, 1, 2) if simt else wp.view(k, k, k, Cout, Cin).permute(3, 4, 0, 1, 2)
            r = F.conv3d(xs, wt, None if bias is None else bias.t.float(), stride=2 if mode == lib.CONV_K2S2 else 1,
                         padding=1 if mode == lib.CONV_K3 else 0)
        if stats is not None:
            stats.t[:, 0] += r.double().flatten(1).sum(1)
            stats.t[:, 1] += (r.double() ** 2).flatten(1).sum(1)
        rn = r.permute(0, 2, 3, 4, 1).reshape(-1, Cout)
        dst = _rows(y, rn.shape[0], y_ld, min(Cout, y_ld) if out_f32 else Cout)
        dst.copy_(rn[:, :dst.shape[1]].to(dst.dtype))
        calls.append('conv%d' % mode)
        return 0

    def gn_apply(dtype, y, y_ld, C, stats, gamma, beta, eps, res, res_ld, out, out_ld, relu, N, nvox, stream):
        ys = _rows(y, N * nvox, y_ld, C).float().view(N, nvox, C)
        cnt = float(nvox * C)
        mean = (stats.t[:, 0] / cnt).view(N, 1, 1)
        var = (stats.t[:, 1] / cnt).view(N, 1, 1) - mean * mean
        z = ((ys.double() - mean) / torch.sqrt(var + eps)).float() * gamma.t.view(1, 1, C) + beta.t.view(1, 1, C)
        if res is not None:
            z = z + _rows(res, N * nvox, res_ld, C).float().view(N, nvox, C)
        if relu:
            z = F.relu(z)
        dst = _rows(out, N * nvox, out_ld, C)
        dst.copy_(z.reshape(-1, C).to(dst.dtype))
        calls.append('gn')
        return 0

    def _tail(y1, ld, C, stats1, g1, b1, w2, bias2, eps, N, nvox):
        ys = _rows(y1, N * nvox, ld, C).float().view(N, nvox, C)
        cnt = float(nvox * C)
        mean = (stats1.t[:, 0] / cnt).view(N, 1, 1)
        var = (stats1.t[:, 1] / cnt).view(N, 1, 1) - mean * mean
        a = F.relu(((ys.double() - mean) / torch.sqrt(var + eps)).float() * g1.t.view(1, 1, C) + b1.t.view(1, 1, C))
        return a @ w2.t.view(C, C).t() + bias2.t.view(1, 1, C)                          # w2 is [out][in]

    def tail_stats(dtype, y1, ld, C, stats1, g1, b1, w2, bias2, eps, stats2, N, nvox, stream):
        z = _tail(y1, ld, C, stats1, g1, b1, w2, bias2, eps, N, nvox)
        stats2.t[:, 0] += z.double().flatten(1).sum(1)
        stats2.t[:, 1] += (z.double() ** 2).flatten(1).sum(1)
        calls.append('tail_stats')
        return 0

    def tail_probs(dtype, y1, ld, C, stats1, g1, b1, w2, bias2, stats2, g2, b2, eps, probs, N, nvox, stream):
        z = _tail(y1, ld, C, stats1, g1, b1, w2, bias2, eps, N, nvox)
        cnt = float(nvox * C)
        mean = (stats2.t[:, 0] / cnt).view(N, 1, 1)
        var = (stats2.t[:, 1] / cnt).view(N, 1, 1) - mean * mean
        z = ((z.double() - mean) / torch.sqrt(var + eps)).float() * g2.t.view(1, 1, C) + b2.t.view(1, 1, C)
        probs.t.copy_(F.softmax(z, 2).permute(0, 2, 1).reshape(probs.t.shape))
        calls.append('tail_probs')
        return 0

    table = {'seg3d_conv3d_fwd': conv3d_fwd, 'seg3d_gn_apply': gn_apply, 'seg3d_outblock_tail_stats': tail_stats,
             'seg3d_outblock_tail_probs': tail_probs}

    def call(name, *args):
        assert table[name](*args) == 0

    monkeypatch.setattr(lib, 'ptr', ptr)
    monkeypatch.setattr(lib, 'call', call)
    monkeypatch.setattr(lib, 'stream_ptr', lambda: 0)
    monkeypatch.setattr(torch.cuda, 'device', lambda d: contextlib.nullcontext())
    monkeypatch.setattr(blocks, 'check_input', lambda x, c: None)
    return calls


def _sd(module, prefix=''):
    return {prefix + k: v.detach() for k, v in module.state_dict().items()}


def _randomize(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith('gn.weight') or name.endswith('gn1.weight') or name.endswith('gn2.weight'):
                p.copy_(1.0 + 0.3 * torch.randn(p.shape, generator=g))
            elif name.endswith('.bias'):
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
            else:
                fan = p[0].numel() if p.dim() == 5 else 1
                p.copy_(torch.randn(p.shape, generator=g) * (2.0 / fan) ** 0.5)


@pytest.mark.parametrize('mode,tol', [('fp32', 2e-5), ('fp16', 3e-2), ('bf16', 2e-1)])
def test_standalone_modules_wire_the_kernels_like_the_reference_blocks(monkeypatch, mode, tol):
    calls = _install(monkeypatch)
    monkeypatch.setenv('SEG3D_MODE', mode)
    from segmentation3d.network.module.conv_gn_relu3 import BottConvGnRelu3, ConvGnRelu3
    from segmentation3d.network.module.residual_block3 import BottResidualBlock3, ResidualBlock3
    from segmentation3d.network.module.vnet_downblock import DownBlock
    from segmentation3d.network.module.vnet_inblock import InputBlock
    from segmentation3d.network.module.vnet_outblock import OutputBlock
    from segmentation3d.network.module.vnet_upblock import UpBlock
    g = torch.Generator().manual_seed(0)

    def close(got, ref, what):
        assert got.shape == ref.shape and got.dtype == torch.float32, what
        err = float((got - ref).abs().max()) / max(1.0, float(ref.abs().max()))
        assert err <= tol, (what, mode, err)

    with torch.no_grad():
        # ConvGnRelu3 with and without activation / bias
        for do_act, bias in ((True, True), (False, True), (True, False)):
            m = ConvGnRelu3(16, 32, 3, 1, 1, do_act=do_act, bias=bias)
            _randomize(m, 1)
            x = torch.randn((2, 16, 4, 6, 8), generator=g)
            y = F.group_norm(F.conv3d(x, m.conv.weight, m.conv.bias, padding=1), 1, m.gn.weight, m.gn.bias, EPS)
            close(m(x), F.relu(y) if do_act else y, 'ConvGnRelu3')
        m = BottConvGnRelu3(32, 32, 3, 1, 1, 4, do_act=False)
        _randomize(m, 2)
        x = torch.randn((1, 32, 4, 4, 8), generator=g)
        sd = _sd(m, 'b.ops.0.')
        y = onet._conv_gn(onet._conv_gn(onet._conv_gn(x, sd, 'b.ops.0.conv1', True), sd, 'b.ops.0.conv2', True), sd, 'b.ops.0.conv3', False)
        close(m(x), y, 'BottConvGnRelu3')
        # residual blocks == oracle _rblock
        for blk, C in ((ResidualBlock3(32, 3, 1, 1, 2), 32), (ResidualBlock3(16, 3, 1, 1, 1), 16), (BottResidualBlock3(64, 3, 1, 1, 4, 2), 64)):
            _randomize(blk, 3)
            x = torch.randn((2, C, 4, 4, 8), generator=g)
            close(blk(x), onet._rblock(x, _sd(blk, 'r.'), 'r'), type(blk).__name__)
        # input / down / up / output blocks
        m = InputBlock(1, 16)
        _randomize(m, 4)
        x = torch.randn((2, 1, 8, 8, 8), generator=g)
        close(m(x), F.relu(F.group_norm(F.conv3d(x, m.conv.weight, m.conv.bias, padding=1), 1, m.gn.weight, m.gn.bias, EPS)), 'InputBlock')
        for comp in (False, True):
            m = DownBlock(32, 2, compression=comp)
            _randomize(m, 5)
            x = torch.randn((1, 32, 8, 8, 8), generator=g)
            close(m(x), onet._down(x, _sd(m, 'd.'), 'd'), 'DownBlock')
            m = UpBlock(128, 64, 2, compression=comp)
            _randomize(m, 6)
            x, skip = torch.randn((1, 128, 2, 4, 4), generator=g), torch.randn((1, 32, 4, 8, 8), generator=g)
            close(m(x, skip), onet._up(x, skip, _sd(m, 'u.'), 'u'), 'UpBlock')
        for nc in (2, 5):
            m = OutputBlock(32, nc)
            _randomize(m, 7)
            x = torch.randn((2, 32, 4, 4, 8), generator=g)
            y = F.relu(F.group_norm(F.conv3d(x, m.conv1.weight, m.conv1.bias, padding=1), 1, m.gn1.weight, m.gn1.bias, EPS))
            y = F.softmax(F.group_norm(F.conv3d(y, m.conv2.weight, m.conv2.bias), 1, m.gn2.weight, m.gn2.bias, EPS), 1)
            p = m(x)
            close(p, y, 'OutputBlock')
            assert float((p.sum(1) - 1).abs().max()) <= 1e-5
    assert {'conv0', 'conv1', 'conv2', 'gn', 'tail_stats', 'tail_probs'} <= set(calls)


@pytest.mark.parametrize('mode,tol', [('fp32', 2e-5), ('fp16', 3e-2)])
def test_conv_gn_relu3_output_size_like_the_reference_test(monkeypatch, mode, tol):
    """network/module/conv_gn_relu3_test.py:8-43 (the reference's own test of this module): a k3 s1 p1 and a k2 s2 p0
    ConvGnRelu3 from 1 to 16 channels on a [4, 1, 48, 32, 16] batch keep / halve the spatial size - plus the values, against
    torch's functional ops."""
    _install(monkeypatch)
    monkeypatch.setenv('SEG3D_MODE', mode)
    from segmentation3d.network.module.conv_gn_relu3 import ConvGnRelu3
    in_channels, out_channels = 1, 16
    model1 = ConvGnRelu3(in_channels, out_channels, ksize=3, stride=1, padding=1, do_act=True)
    model2 = ConvGnRelu3(in_channels, out_channels, ksize=2, stride=2, padding=0, do_act=True)
    assert sum(p.numel() for p in model1.parameters()) == 16 * 27 + 16 + 32
    assert sum(p.numel() for p in model2.parameters()) == 16 * 8 + 16 + 32
    batch_size, (dim_x, dim_y, dim_z) = 4, (16, 32, 48)
    inputs = torch.rand([batch_size, in_channels, dim_z, dim_y, dim_x], generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        outputs1, outputs2 = model1(inputs), model2(inputs)
        ref1 = F.relu(F.group_norm(F.conv3d(inputs, model1.conv.weight, model1.conv.bias, padding=1), 1, model1.gn.weight, model1.gn.bias, EPS))
        ref2 = F.relu(F.group_norm(F.conv3d(inputs, model2.conv.weight, model2.conv.bias, stride=2), 1, model2.gn.weight, model2.gn.bias, EPS))
    assert tuple(outputs1.size()) == (batch_size, out_channels, dim_z, dim_y, dim_x)
    assert tuple(outputs2.size()) == (batch_size, out_channels, dim_z // 2, dim_y // 2, dim_x // 2)
    assert float((outputs1 - ref1).abs().max()) <= tol * max(1.0, float(ref1.abs().max()))
    assert float((outputs2 - ref2).abs().max()) <= tol * max(1.0, float(ref2.abs().max()))
