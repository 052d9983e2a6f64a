"""CPU check of the HOST LOGIC of the sliding-window engine (core/seg_infer.py::segmentation_volume_device,
_b200/sliding.py::SlidingWindow) against the reference's own outputs.

The engine's kernels (patch statistics / gather, blend, count-normalise + arg-max) are replaced by numpy emulations of their
documented semantics (include/seg3d_b200.h), the network plan by the oracle's forward, "device" tensors are CPU tensors.
What runs for real is the host side: the patch grid, balanced batches and their pointer offsets, the normaliser
parameters, the separable overlap counts, the bounding-box rule, progressive z-slab finalisation.  Results are compared
with the committed reference goldens (tests/golden/sliding_window.npz, cascade.npz), so a host-side regression shows up
in the CPU gate; the kernels themselves are pinned by the `-m gpu` tests."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import init as oinit
from oracle import net as onet

G = os.path.join(os.path.dirname(__file__), 'golden')


class _P(object):
    def __init__(self, t, off):
        self.t, self.off = t, off

    def flat(self):
        return self.t.reshape(-1)[self.off:]


def _install(monkeypatch, calls):
    from segmentation3d._b200 import lib

    def ptr(t, off=0):
        return None if t is None else _P(t, off)

    def _starts(p, n):
        return p.flat()[:3 * n].numpy().reshape(n, 3)

    def patch_stats(vol, Z, Y, X, starts, N, pz, py, px, stats, stream):
        v = vol.t.numpy().reshape(Z, Y, X)
        for n, s in enumerate(_starts(starts, N)):
            p = v[s[2]:s[2] + pz, s[1]:s[1] + py, s[0]:s[0] + px].astype(np.float64)
            assert p.shape == (pz, py, px)
            stats.t[n, 0] += float(p.sum())
            stats.t[n, 1] += float((p * p).sum())
        calls.append(('stats', N))
        return 0

    def patch_gather(vol, Z, Y, X, starts, N, pz, py, px, norm, mean, std, clip, lo, hi, stats, dtype, out, stream):
        v = vol.t.numpy().reshape(Z, Y, X)
        res = []
        for n, s in enumerate(_starts(starts, N)):
            p = v[s[2]:s[2] + pz, s[1]:s[1] + py, s[0]:s[0] + px].astype(np.float32)
            assert p.shape == (pz, py, px)
            if norm == lib.NORM_ADAPTIVE:                      # utils/normalizer.py:55-62 in float32, as numpy computes it
                m, sd = np.mean(p), max(np.std(p), 1e-6)
                p = (p - m) / sd
            elif norm == lib.NORM_FIXED:
                p = (p - mean) / std
            if clip:
                p = np.clip(p, lo, hi)
            res.append(p.astype(np.float32))
        out.t.reshape(-1)[out.off:out.off + N * pz * py * px] = torch.from_numpy(np.stack(res).reshape(-1))
        calls.append(('gather', N))
        return 0

    def patch_gather_rows(vol, Z, Y, X, starts, N, pz, py, px, norm, mean, std, clip, lo, hi, stats, dtype, out, row_pitch, x_off, stream):
        # the same crop into rows of pitch row_pitch at column x_off; every other column is left as it is
        tmp = torch.empty((N * pz * py * px,), dtype=torch.float32)
        patch_gather(vol, Z, Y, X, starts, N, pz, py, px, norm, mean, std, clip, lo, hi, stats, dtype, type(out)(tmp, 0), stream)
        rows = torch.as_strided(out.t.reshape(-1), (N * pz * py, px), (row_pitch, 1), out.off + x_off)
        rows.copy_(tmp.view(N * pz * py, px).to(rows.dtype))
        return 0

    def blend_accumulate(probs, N, C, pz, py, px, starts, acc, Z, Y, X, x_mult4, stream):
        assert (x_mult4 == 0) or all(s[0] % 4 == 0 for s in _starts(starts, N))
        a = acc.t.numpy().reshape(C, Z, Y, X)
        p = probs.t.numpy().reshape(-1, C, pz, py, px)
        for n, s in enumerate(_starts(starts, N)):
            a[:, s[2]:s[2] + pz, s[1]:s[1] + py, s[0]:s[0] + px] += p[n]
        calls.append(('blend', N))
        return 0

    def finalize_z(acc, C, Z, Y, X, z0, z1, cx, cy, cz, mask, stream):
        a = acc.t.numpy().reshape(C, Z, Y, X)
        cnt = (cz.t.numpy()[z0:z1, None, None].astype(np.float32) * cy.t.numpy()[None, :, None] * cx.t.numpy()[None, None, :])
        with np.errstate(divide='ignore', invalid='ignore'):
            a[:, z0:z1] *= (np.float32(1.0) / cnt).astype(np.float32)
        if mask is not None:
            m = mask.t.numpy().reshape(Z, Y, X)
            m[z0:z1] = torch.from_numpy(a[:, z0:z1].copy()).max(0)[1].numpy().astype(np.int8)     # first arg-max
        calls.append(('finalize', z0, z1))
        return 0

    table = {'seg3d_patch_stats': patch_stats, 'seg3d_patch_gather': patch_gather, 'seg3d_patch_gather_rows': patch_gather_rows,
             'seg3d_blend_accumulate': blend_accumulate,
             'seg3d_blend_finalize_argmax_z': finalize_z}
    monkeypatch.setattr(lib, 'ptr', ptr)
    monkeypatch.setattr(lib, 'call', lambda name, *a: table[name](*a))
    monkeypatch.setattr(lib, 'stream_ptr', lambda: 0)
    return lib


class _OraclePlan(object):
    """stands in for NetPlan: same attributes the engine reads, forward = the oracle's network program"""

    def __init__(self, sd, cout, lib):
        self.sd, self.in_channels, self.out_channels, self.in_dt = sd, 1, cout, lib.F32
        self.forwards = []

    def plan(self, nb, pz, py, px):
        return {'x_in': torch.empty((nb, pz * py * px, 1), dtype=torch.float32)}, (nb, pz, py, px)

    def launches(self, ws):
        return 1

    def run(self, ws, ops):
        nb, pz, py, px = ops
        self.forwards.append(nb)
        return onet.forward(self.sd, ws['x_in'].view(nb, 1, pz, py, px)).contiguous()


def _model(sd, cout, nd, lib, batch):
    from segmentation3d._b200.sliding import SlidingWindow
    from segmentation3d.utils.normalizer import normalizer_from_dict
    plan = _OraclePlan(sd, cout, lib)

    class Net(object):
        def _current_plan(self):
            return plan
    return {'net': Net(), 'engine': SlidingWindow(plan, batch), 'crop_normalizers': [normalizer_from_dict(nd)],
            'out_channels': cout, 'spacing': [1.0, 1.0, 1.0], 'max_stride': 16}, plan


def _seeded(seed, shape):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g)
    lo = torch.randn((shape[0], shape[1]) + tuple(max(2, s // 8) for s in shape[2:]), generator=g)
    return torch.nn.functional.interpolate(lo, size=shape[2:], mode='trilinear', align_corners=False) * 2 + 0.3 * x


def test_engine_host_logic_reproduces_reference_sliding_window(monkeypatch):
    from segmentation3d.core.seg_infer import segmentation_volume_device
    calls = []
    lib = _install(monkeypatch, calls)
    z = np.load(os.path.join(G, 'sliding_window.npz'))
    for name, arch, cout, wseed, aseed, size, psize, pstride, norm, vseed, scale in json.loads(str(z['meta'])):
        sd = oinit.init_state_dict(arch, 1, cout, wseed)
        if aseed is not None:
            sd = oinit.randomize_affine(sd, aseed)
        vol = (_seeded(vseed, (1, 1, size[2], size[1], size[0]))[0, 0].numpy() * scale).astype(np.float32)
        nd = {'type': 0, 'mean': norm[1], 'stddev': norm[2], 'clip': norm[3]} if norm[0] == 'fixed' else {'type': 1, 'clip_sigma': norm[1]}
        model, plan = _model(sd, cout, nd, lib, batch=5)
        cfg = {'partition_type': 'SIZE', 'partition_size': psize, 'partition_stride': pstride}
        del calls[:]
        acc, mask = segmentation_volume_device(model, cfg, torch.from_numpy(vol))
        assert np.abs(acc.numpy() - z[name + '_probs']).max() <= 1e-5, name
        assert (mask.numpy() == z[name + '_mask']).mean() >= 0.99999 and mask.dtype == torch.int8, name
        n_patches = sum(c[1] for c in calls if c[0] == 'gather')
        assert sum(plan.forwards) == n_patches and max(plan.forwards) <= 5
        assert max(plan.forwards) - min(plan.forwards) <= 1                     # balanced batches
        assert ('stats' in [c[0] for c in calls]) == (norm[0] == 'adaptive')


def test_engine_host_logic_reproduces_reference_bounding_box_run(monkeypatch):
    from segmentation3d.core.seg_infer import segmentation_volume_device
    calls = []
    lib = _install(monkeypatch, calls)
    z = np.load(os.path.join(G, 'cascade.npz'))
    m = json.loads(str(z['meta']))
    sd = oinit.randomize_affine(oinit.init_state_dict(m['arch'], 1, m['cout'], m['wseed']), m['aseed'])
    size = m['size']
    vol = (_seeded(m['vseed'], (1, 1, size[2], size[1], size[0]))[0, 0].numpy() * m['scale']).astype(np.float32)
    nd = {'type': 0, 'mean': m['norm'][1], 'stddev': m['norm'][2], 'clip': m['norm'][3]}
    model, plan = _model(sd, m['cout'], nd, lib, batch=7)
    cfg = {'partition_type': 'SIZE', 'partition_size': m['psize'], 'partition_stride': m['pstride']}
    acc, mask = segmentation_volume_device(model, cfg, torch.from_numpy(vol), bbox_start_voxel=list(m['bbox_start']),
                                           bbox_end_voxel=list(m['bbox_end']))
    visited = np.isfinite(z['probs'][0])
    assert np.abs(acc.numpy()[:, visited] - z['probs'][:, visited]).max() <= 1e-5
    assert np.array_equal(mask.numpy(), z['mask'])
    assert not acc.numpy()[:, ~visited].any()                                   # probability 0 where the reference holds NaN


def test_engine_whole_volume_partition_and_patch_sharding(monkeypatch):
    """partition_type = 'DISABLE' (one patch = the volume) and the rank::world patch deal: the per-rank partial accumulators
    add up to the single-process result."""
    from oracle import sliding_window as osw
    from segmentation3d.core.seg_infer import _grid, segmentation_volume_device
    calls = []
    lib = _install(monkeypatch, calls)
    sd = oinit.randomize_affine(oinit.init_state_dict('vnet', 1, 2, 3), 4)
    vol = (_seeded(5, (1, 1, 32, 48, 32))[0, 0].numpy() * 100).astype(np.float32)
    nd = {'type': 0, 'mean': 0.0, 'stddev': 100.0, 'clip': False}
    model, plan = _model(sd, 2, nd, lib, batch=0)
    acc, mask = segmentation_volume_device(model, {'partition_type': 'DISABLE'}, torch.from_numpy(vol))
    ref, ref_mask, _, _ = osw.segmentation_volume(sd, vol, [1.0, 1.0, 1.0], nd, 'DISABLE', double_forward=False, faithful_copies=False)
    assert np.abs(acc.numpy() - ref).max() <= 1e-5 and np.array_equal(mask.numpy(), ref_mask) and plan.forwards == [1]
    cfg = {'partition_type': 'SIZE', 'partition_size': [32, 32, 32], 'partition_stride': [16, 16, 16]}
    full, _ = segmentation_volume_device(model, cfg, torch.from_numpy(vol))
    starts, ends = _grid(model, cfg, [32, 48, 32], [1.0, 1.0, 1.0], None, None)
    # what each rank accumulates before the collective: patches rank::world, no normalisation yet
    parts = []
    for r in range(3):
        part = torch.zeros((2, 32, 48, 32))
        model['engine'].accumulate(torch.from_numpy(vol), starts[r::3], [32, 32, 32], nd, part)
        parts.append(part)
    total = sum(parts)
    from segmentation3d._b200.sliding import axis_counts
    model['engine'].finalize(total, axis_counts([32, 48, 32], starts, ends))
    assert np.abs(total.numpy() - full.numpy()).max() <= 1e-5
