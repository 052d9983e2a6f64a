"""CPU restatement of the resampling either side of the hot path.  TEST INFRASTRUCTURE ONLY.

reference: segmentation3d/utils/image_tools.py:329-343 (`resample`: sitk.Resample(image, reference, identity, interp,
padding)) and :346-377 (`resample_spacing`: sitk.Resample(image, out_size, identity, interp, in_origin, out_spacing,
in_direction)), called from core/seg_infer.py:267 and :330-333.

The arithmetic lives in the third-party dependency SimpleITK (README pin 1.2.3, ITK 4.13), which is absent here, so this
is a restatement of ITK's published algorithm - ResampleImageFilter with an identity transform,
LinearInterpolateImageFunction / NearestNeighborInterpolateImageFunction, double-precision continuous indices - and the
parity of this row is UNPINNED (no golden vectors from the reference itself exist or can be generated in this image).

Both call sites resample between grids with the same origin and direction, so output index i maps to the continuous
input index c = i * spacing_out / spacing_in per axis.
"""
import numpy as np


def out_size(in_size, in_spacing, out_spacing, max_stride):
    """image_tools.py:363-366."""
    out = [int(in_size[i] * in_spacing[i] / out_spacing[i] + 0.5) for i in range(3)]
    for i in range(3):
        if out[i] % max_stride:
            out[i] = max_stride * (out[i] // max_stride + 1)
    return out


def resample_grid(src_zyx, in_spacing_xyz, out_size_xyz, out_spacing_xyz, interp='LINEAR', default_value=0.0):
    """src [z,y,x] float32 on spacing in_spacing -> [z,y,x] float32 of out_size on out_spacing (same origin/direction)."""
    src = np.asarray(src_zyx, dtype=np.float32)
    sz, sy, sx = src.shape
    dx, dy, dz = [int(v) for v in out_size_xyz]
    rx, ry, rz = [float(out_spacing_xyz[a]) / float(in_spacing_xyz[a]) for a in range(3)]
    cz = np.arange(dz, dtype=np.float64) * rz
    cy = np.arange(dy, dtype=np.float64) * ry
    cx = np.arange(dx, dtype=np.float64) * rx
    inside = (cz < sz - 0.5)[:, None, None] & (cy < sy - 0.5)[None, :, None] & (cx < sx - 0.5)[None, None, :]
    if interp == 'NN':
        zi = np.minimum(np.floor(cz + 0.5).astype(np.int64), sz - 1)
        yi = np.minimum(np.floor(cy + 0.5).astype(np.int64), sy - 1)
        xi = np.minimum(np.floor(cx + 0.5).astype(np.int64), sx - 1)
        val = src[zi[:, None, None], yi[None, :, None], xi[None, None, :]].astype(np.float32)
    elif interp == 'LINEAR':
        def axis(c, n):
            f = np.floor(c)
            i0 = np.minimum(f.astype(np.int64), n - 1)
            i1 = np.minimum(i0 + 1, n - 1)
            return i0, i1, c - f
        z0, z1, wz = axis(cz, sz)
        y0, y1, wy = axis(cy, sy)
        x0, x1, wx = axis(cx, sx)
        s = src.astype(np.float64)
        wx_ = wx[None, None, :]
        wy_ = wy[None, :, None]
        wz_ = wz[:, None, None]

        def g(zi, yi, xi):
            return s[zi[:, None, None], yi[None, :, None], xi[None, None, :]]
        a00 = g(z0, y0, x0) + (g(z0, y0, x1) - g(z0, y0, x0)) * wx_
        a01 = g(z0, y1, x0) + (g(z0, y1, x1) - g(z0, y1, x0)) * wx_
        a10 = g(z1, y0, x0) + (g(z1, y0, x1) - g(z1, y0, x0)) * wx_
        a11 = g(z1, y1, x0) + (g(z1, y1, x1) - g(z1, y1, x0)) * wx_
        b0 = a00 + (a01 - a00) * wy_
        b1 = a10 + (a11 - a10) * wy_
        val = (b0 + (b1 - b0) * wz_).astype(np.float32)
    else:
        raise ValueError('Unsupported interpolation type.')      # image_tools.py:340,372
    return np.where(inside, val, np.float32(default_value)).astype(np.float32)


def resample_spacing(src_zyx, in_spacing_xyz, out_spacing_xyz, max_stride, interp='LINEAR'):
    """image_tools.py:346-377 -> (resampled [z,y,x], out_size_xyz)."""
    in_size = [src_zyx.shape[2], src_zyx.shape[1], src_zyx.shape[0]]
    osz = out_size(in_size, in_spacing_xyz, out_spacing_xyz, max_stride)
    return resample_grid(src_zyx, in_spacing_xyz, osz, out_spacing_xyz, interp, 0.0), osz
