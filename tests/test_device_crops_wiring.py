"""CPU check of the HOST WIRING of the device-side training crops (dataloader/device_loader.py, image_tools.crop_image_device).

The three entry points involved (seg3d_crop_resample, seg3d_patch_stats, seg3d_patch_gather) are replaced by numpy
emulations of their documented semantics (include/seg3d_b200.h), "device" tensors are CPU tensors, and the batches the
DeviceCropLoader yields are compared with the DataLoader(SegmentationDataset) path under the same numpy seed: same crop
geometry, same RNG order, same normaliser, same item layout.  The kernels themselves still have to see a GPU
(tests/test_gpu_blocks.py::test_device_crops_match_host_crops, gated)."""
import contextlib
import os

import numpy as np
import pytest
import torch

from segmentation3d.utils.image3d import Image3d, write_image


class _P(object):
    def __init__(self, t, off):
        self.t, self.off = t, off


def _install(monkeypatch):
    from segmentation3d._b200 import lib
    from segmentation3d.utils import image3d
    calls = []

    def ptr(t, off=0):
        return None if t is None else _P(t, off)

    def crop_resample(src, sz, sy, sx, dst, dz, dy, dx, oz, oy, ox, rz, ry, rx, linear, dflt, stream):
        s = src.t.numpy().reshape(sz, sy, sx).astype(np.float64)
        c = [o + np.arange(n, dtype=np.float64) * r for o, n, r in ((ox, dx, rx), (oy, dy, ry), (oz, dz, rz))]
        n_in = [sx, sy, sz]
        ok = [(c[a] >= -0.5) & (c[a] < n_in[a] - 0.5) for a in range(3)]
        inside = ok[2][:, None, None] & ok[1][None, :, None] & ok[0][None, None, :]
        if linear:
            f = [np.floor(v) for v in c]
            lo = [np.clip(f[a].astype(np.int64), 0, n_in[a] - 1) for a in range(3)]
            hi = [np.clip(f[a].astype(np.int64) + 1, 0, n_in[a] - 1) for a in range(3)]
            w = [c[a] - f[a] for a in range(3)]
            wx, wy, wz = w[0][None, None, :], w[1][None, :, None], w[2][:, None, None]

            def g(zi, yi, xi):
                return s[np.ix_(zi, yi, xi)]
            a00 = g(lo[2], lo[1], lo[0]) + (g(lo[2], lo[1], hi[0]) - g(lo[2], lo[1], lo[0])) * wx
            a01 = g(lo[2], hi[1], lo[0]) + (g(lo[2], hi[1], hi[0]) - g(lo[2], hi[1], lo[0])) * wx
            a10 = g(hi[2], lo[1], lo[0]) + (g(hi[2], lo[1], hi[0]) - g(hi[2], lo[1], lo[0])) * wx
            a11 = g(hi[2], hi[1], lo[0]) + (g(hi[2], hi[1], hi[0]) - g(hi[2], hi[1], lo[0])) * wx
            b0, b1 = a00 + (a01 - a00) * wy, a10 + (a11 - a10) * wy
            val = b0 + (b1 - b0) * wz
        else:
            idx = [np.clip(np.floor(c[a] + 0.5).astype(np.int64), 0, n_in[a] - 1) for a in range(3)]
            val = s[np.ix_(idx[2], idx[1], idx[0])]
        out = np.where(inside, val, dflt).astype(np.float32)
        flat = dst.t.reshape(-1)
        assert dst.t.is_contiguous()
        flat[dst.off:dst.off + out.size] = torch.from_numpy(out.reshape(-1))
        calls.append('crop%d' % linear)
        return 0

    def _patches(vol, Z, Y, X, starts, N, pz, py, px):
        v = vol.t.numpy().reshape(Z, Y, X)
        st = starts.t.numpy().reshape(-1, 3)
        return [v[s[2]:s[2] + pz, s[1]:s[1] + py, s[0]:s[0] + px] for s in st[:N]]

    def patch_stats(vol, Z, Y, X, starts, N, pz, py, px, stats, stream):
        for n, p in enumerate(_patches(vol, Z, Y, X, starts, N, pz, py, px)):
            stats.t[n, 0] += float(p.astype(np.float64).sum())
            stats.t[n, 1] += float((p.astype(np.float64) ** 2).sum())
        calls.append('stats')
        return 0

    def patch_gather(vol, Z, Y, X, starts, N, pz, py, px, norm, mean, std, clip, lo, hi, stats, dtype, out, stream):
        assert dtype == lib.F32
        res = []
        for n, p in enumerate(_patches(vol, Z, Y, X, starts, N, pz, py, px)):
            p = p.astype(np.float32)
            if norm == lib.NORM_ADAPTIVE:
                cnt = float(p.size)
                m = stats.t[n, 0].item() / cnt
                sd = max(np.sqrt(max(stats.t[n, 1].item() / cnt - m * m, 0.0)), 1e-6)
                p = ((p - np.float32(m)) / np.float32(sd)).astype(np.float32)
            elif norm == lib.NORM_FIXED:
                p = ((p - np.float32(mean)) / np.float32(std)).astype(np.float32)
            if clip:
                p = np.clip(p, np.float32(lo), np.float32(hi))
            res.append(p)
        out.t.copy_(torch.from_numpy(np.stack(res)).view(out.t.shape))
        calls.append('gather%d' % norm)
        return 0

    table = {'seg3d_crop_resample': crop_resample, 'seg3d_patch_stats': patch_stats, 'seg3d_patch_gather': patch_gather}
    monkeypatch.setattr(lib, 'ptr', ptr)
    monkeypatch.setattr(lib, 'call', lambda name, *a: table[name](*a))
    monkeypatch.setattr(lib, 'stream_ptr', lambda: 0)
    monkeypatch.setattr(torch.cuda, 'device', lambda d: contextlib.nullcontext())
    monkeypatch.setattr(image3d.Image3d, 'is_cuda', lambda self: torch.is_tensor(self.data))   # CPU tensors stand in for device tensors
    return calls


def _cases(tmp_path):
    rng = np.random.RandomState(5)
    lines = ['3']
    for k, (size, spacing, origin) in enumerate((((40, 36, 30), (1.0, 1.0, 1.0), (0.0, 0.0, 0.0)),
                                                 ((48, 40, 20), (0.8, 0.8, 1.5), (-10.0, 5.0, 30.0)),
                                                 ((30, 44, 38), (1.2, 0.9, 1.0), (3.5, -7.25, 12.0)))):
        im = (rng.standard_normal((size[2], size[1], size[0])) * 80 + 20).astype(np.float32)
        lab = rng.randint(0, 3, size=(size[2], size[1], size[0])).astype(np.float32)
        d = tmp_path / ('case%d' % k)
        os.makedirs(d)
        write_image(Image3d(im, spacing, origin), str(d / 'im.mha'), True)
        write_image(Image3d(lab, spacing, origin), str(d / 'seg.mha'), True)
        lines += [str(d / 'im.mha'), str(d / 'seg.mha')]
    with open(str(tmp_path / 'train.txt'), 'w') as f:
        f.write('\n'.join(lines) + '\n')
    return str(tmp_path / 'train.txt')


@pytest.mark.parametrize('norm_name', ['fixed', 'adaptive', 'none'])
def test_device_crop_loader_yields_the_dataloader_batches(tmp_path, monkeypatch, norm_name):
    import random
    from torch.utils.data import DataLoader
    from segmentation3d.dataloader.dataset import SegmentationDataset
    from segmentation3d.dataloader.device_loader import DeviceCropLoader
    from segmentation3d.dataloader.sampler import EpochConcateSampler
    from segmentation3d.utils.normalizer import AdaptiveNormalizer, FixedNormalizer
    calls = _install(monkeypatch)
    txt = _cases(tmp_path)
    norm = {'fixed': FixedNormalizer(20.0, 80.0, True), 'adaptive': AdaptiveNormalizer(2.5), 'none': None}[norm_name]

    def make():
        return SegmentationDataset(txt, 3, [1.0, 1.0, 1.2], [16, 24, 16], 'HYBRID', [5, 4, 3], [0.9, 1.1], 'LINEAR', [norm])
    ds = make()
    random.seed(3)
    np.random.seed(11)
    ref = list(DataLoader(ds, sampler=EpochConcateSampler(ds, 2), batch_size=4, num_workers=0))
    ds = make()
    random.seed(3)
    np.random.seed(11)
    loader = DeviceCropLoader(ds, EpochConcateSampler(ds, 2), 4, device='cpu', cache_gb=1)
    got = list(loader)
    assert len(got) == len(ref) == len(loader) == 2 and got[1][0].shape[0] == 2          # 6 items: 4 + 2
    for (c0, m0, f0, n0), (c1, m1, f1, n1) in zip(ref, got):
        assert tuple(c1.shape) == tuple(c0.shape) and c1.dtype == torch.float32 and tuple(m1.shape) == tuple(m0.shape)
        assert list(n0) == list(n1)
        assert torch.allclose(f0.float(), f1.float(), atol=1e-5)
        assert torch.equal(m0, m1)                                                        # nearest-neighbour labels
        assert float((c0 - c1).abs().max()) <= 2e-5, norm_name
    assert 'crop1' in calls and 'crop0' in calls
    if norm_name == 'adaptive':
        assert 'stats' in calls and 'gather2' in calls
    elif norm_name == 'fixed':
        assert 'gather1' in calls
    # volumes are uploaded once and stay resident
    assert len(loader.resident) == 6 and loader.bytes == sum(v.data.numel() * 4 for v in loader.resident.values())
