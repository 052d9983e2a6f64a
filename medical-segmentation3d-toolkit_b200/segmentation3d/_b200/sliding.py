"""Device-resident sliding-window engine (reference core/seg_infer.py:313-339).

The resampled volume, the C fp32 accumulators and the mask stay in HBM; patches are cropped and
normalised by a kernel, pushed through the network plan several at a time (GroupNorm statistics
are per sample, so batching does not change results), blended with red.global.add and finalised
with the analytic overlap count.  One forward per patch: the reference's second forward
(seg_infer.py:230-234) is bit-identical to the first and is not repeated.
"""
import numpy as np
import torch

from . import lib


def axis_counts(size_xyz, starts, ends):
    """Separable overlap count: the reference grid (utils/image_tools.py:202-216) is the Cartesian
    product of per-axis box lists, so count(x,y,z) = cx[x]*cy[y]*cz[z].  Returns three int32 arrays."""
    starts = np.asarray(starts, dtype=np.int64).reshape(-1, 3)
    ends = np.asarray(ends, dtype=np.int64).reshape(-1, 3)
    boxes = [sorted(set(zip(starts[:, a].tolist(), ends[:, a].tolist()))) for a in range(3)]
    if len(boxes[0]) * len(boxes[1]) * len(boxes[2]) != len(starts):
        raise ValueError('patch list is not a Cartesian product of per-axis boxes')
    out = []
    for a in range(3):
        c = np.zeros(int(size_xyz[a]), dtype=np.int32)
        for s, e in boxes[a]:
            c[s:e] += 1
        out.append(c)
    return out


def batch_ranges(n, batch):
    """(start, count) of the patch batches for n patches with at most `batch` per forward, balanced: 23 patches with
    batch 20 run as 12 + 11, not 20 + 3 (a 3-patch forward is latency-bound and costs nearly as much as a 12-patch one)."""
    if n <= 0:
        return []
    nbat = (n + batch - 1) // batch
    per = (n + nbat - 1) // nbat
    return [(b0, min(per, n - b0)) for b0 in range(0, n, per)]


class SlidingWindow(object):
    def __init__(self, plan, batch=0):
        self.plan = plan
        self.batch = int(batch)          # 0: let the caller pick from the patch size
        self.kernel_launches = 0

    def accumulate(self, vol, starts, patch_xyz, normalizer, acc):
        """Run patches `starts` (list of [x,y,z]) of size patch_xyz through the net and add their
        probabilities into acc [C,Z,Y,X] (fp32, CUDA)."""
        plan = self.plan
        if plan.in_channels != 1:
            raise RuntimeError('sliding window supports single-modality volumes (as the reference engine does)')
        Z, Y, X = vol.shape
        px, py, pz = [int(v) for v in patch_xyz]
        n = len(starts)
        if n == 0:
            return
        x_mult4 = 1 if all(int(s[0]) % 4 == 0 for s in starts) else 0
        starts_dev = torch.tensor(np.asarray(starts, dtype=np.int32).reshape(-1, 3), dtype=torch.int32, device=vol.device)
        norm, mean, std, clip, lo, hi = lib.NORM_NONE, 0.0, 1.0, 0, -1.0, 1.0
        if normalizer is not None:
            if normalizer['type'] == 0:
                norm, mean, std = lib.NORM_FIXED, float(normalizer['mean']), float(normalizer['stddev'])
                clip = 1 if normalizer['clip'] else 0
            elif normalizer['type'] == 1:
                norm, clip = lib.NORM_ADAPTIVE, 1
                lo, hi = -float(normalizer['clip_sigma']), float(normalizer['clip_sigma'])
            else:
                raise ValueError('Unsupported normalization type.')
        st = lib.stream_ptr
        C = plan.out_channels
        if self.batch <= 0:
            self.batch = 6
        for b0, nb in batch_ranges(n, self.batch):
            ws, ops = plan.plan(nb, pz, py, px)
            sp = lib.ptr(starts_dev, 3 * b0)
            pstats = None
            if norm == lib.NORM_ADAPTIVE:
                pstats = ws.setdefault('patch_stats', torch.zeros((nb, 2), dtype=torch.float64, device=vol.device))
                pstats.zero_()
                lib.call('seg3d_patch_stats', lib.ptr(vol), Z, Y, X, sp, nb, pz, py, px, lib.ptr(pstats), st())
                self.kernel_launches += 1
            if ws.get('x_pad'):          # row-padded input layout of the Toeplitz input block (plan.py)
                lib.call('seg3d_patch_gather_rows', lib.ptr(vol), Z, Y, X, sp, nb, pz, py, px, norm, mean, std, clip, lo, hi,
                         lib.ptr(pstats), plan.in_dt, lib.ptr(ws['x_in']), px + lib.CIN1_PAD, lib.CIN1_LEFT, st())
            else:
                lib.call('seg3d_patch_gather', lib.ptr(vol), Z, Y, X, sp, nb, pz, py, px, norm, mean, std, clip, lo, hi,
                         lib.ptr(pstats), plan.in_dt, lib.ptr(ws['x_in']), st())
            probs = plan.run(ws, ops)
            lib.call('seg3d_blend_accumulate', lib.ptr(probs), nb, C, pz, py, px, sp, lib.ptr(acc), Z, Y, X, x_mult4, st())
            self.kernel_launches += 2 + plan.launches(ws)

    def finalize(self, acc, counts, want_mask=True, z_range=None, mask=None):
        """acc *= 1/count in place; returns the int8 first-argmax mask [Z,Y,X].  z_range=(z0, z1) finishes only those
        planes (into the caller's `mask`), so a slab no remaining patch touches can be finished and copied out early."""
        C, Z, Y, X = acc.shape
        key = (tuple(np.asarray(c).tobytes() for c in counts), str(acc.device))
        if getattr(self, '_counts_key', None) != key:
            self._counts_dev = [torch.as_tensor(c, dtype=torch.int32, device=acc.device) for c in counts]
            self._counts_key = key
        cx, cy, cz = self._counts_dev
        if mask is None:
            mask = torch.empty((Z, Y, X), dtype=torch.int8, device=acc.device) if want_mask else None
        z0, z1 = z_range if z_range is not None else (0, Z)
        lib.call('seg3d_blend_finalize_argmax_z', lib.ptr(acc), C, Z, Y, X, int(z0), int(z1), lib.ptr(cx), lib.ptr(cy), lib.ptr(cz),
                 lib.ptr(mask), lib.stream_ptr())
        self.kernel_launches += 1
        return mask

    def segment(self, vol, starts, ends, normalizer):
        """Full single-GPU pass: returns (mean_probs [C,Z,Y,X] fp32, mask [Z,Y,X] int8), both on device."""
        Z, Y, X = vol.shape
        patch = [ends[0][i] - starts[0][i] for i in range(3)]
        for s, e in zip(starts, ends):
            assert [e[i] - s[i] for i in range(3)] == patch, 'all boxes of a partition have the same size'
        acc = torch.zeros((self.plan.out_channels, Z, Y, X), dtype=torch.float32, device=vol.device)
        self.accumulate(vol, starts, patch, normalizer, acc)
        mask = self.finalize(acc, axis_counts([X, Y, Z], starts, ends))
        return acc, mask
