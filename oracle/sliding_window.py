"""Oracle: patch-partitioned sliding-window inference on numpy volumes.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Volumes are numpy arrays in [z, y, x] order (what sitk.GetArrayFromImage returns,
utils/image_tools.py:279); voxel coordinates are [x, y, z] lists like the reference's.

Reference (relative to /root/reference):
  segmentation3d/utils/image_tools.py:163-218   image_partition_by_fixed_size   (integer grid, bit-exact)
  segmentation3d/utils/image_tools.py:346-377   resample_spacing (only the output-size rounding is restated)
  segmentation3d/utils/image_tools.py:221-238   normalize_image
  segmentation3d/utils/normalizer.py:22-25,55-62  Fixed / Adaptive normalisers
  segmentation3d/utils/image_tools.py:435-469   add_image_region / add_image_value
  segmentation3d/core/seg_infer.py:208-246      segmentation_voi (two forwards, mean)
  segmentation3d/core/seg_infer.py:249-350      segmentation_volume
"""
import copy

import numpy as np
import torch

from . import net as onet


def partition_grid(image_size, image_spacing, bbox_start_voxel, bbox_end_voxel,
                   partition_size, partition_stride, max_stride):
    """image_tools.py:163-218.  Returns (start_voxels, end_voxels), lists of [x,y,z] ints, x outermost.
    Mutates its bbox list arguments in place exactly like the reference (:184-187)."""
    image_size = [int(v) for v in image_size]
    for idx in range(3):
        assert image_size[idx] >= max_stride and image_size[idx] % max_stride == 0          # :177

    bbox_size = [min(image_size[idx], bbox_end_voxel[idx] - bbox_start_voxel[idx]) for idx in range(3)]
    for idx in range(3):
        if bbox_size[idx] % max_stride != 0:
            bbox_size[idx] = max_stride * (bbox_size[idx] // max_stride + 1)
        bbox_size[idx] = min(bbox_size[idx], image_size[idx])
        bbox_end_voxel[idx] = bbox_start_voxel[idx] + bbox_size[idx]
        if bbox_end_voxel[idx] > image_size[idx]:
            bbox_end_voxel[idx] = image_size[idx]
            bbox_start_voxel[idx] = bbox_end_voxel[idx] - bbox_size[idx]
        assert bbox_start_voxel[idx] >= 0

    box_size = [int(partition_size[idx] / image_spacing[idx] + 0.5) for idx in range(3)]   # :190
    for idx in range(3):
        if box_size[idx] % max_stride:
            box_size[idx] = max_stride * (box_size[idx] // max_stride + 1)
        box_size[idx] = min(bbox_size[idx], box_size[idx])

    stride_size = [int(partition_stride[idx] / image_spacing[idx] + 0.5) for idx in range(3)]  # :196
    for idx in range(3):
        stride_size[idx] = min(bbox_size[idx], stride_size[idx])

    num_partitions = [int(np.ceil((bbox_size[idx] - box_size[idx]) / stride_size[idx]) + 1) for idx in range(3)]
    start_voxels, end_voxels = [], []
    for ix in range(num_partitions[0]):
        for iy in range(num_partitions[1]):
            for iz in range(num_partitions[2]):
                start = [bbox_start_voxel[0] + ix * stride_size[0],
                         bbox_start_voxel[1] + iy * stride_size[1],
                         bbox_start_voxel[2] + iz * stride_size[2]]
                end = [start[d] + box_size[d] for d in range(3)]
                for d in range(3):
                    if end[d] > bbox_end_voxel[d]:
                        end[d] = bbox_end_voxel[d]
                        start[d] = end[d] - box_size[d]
                        assert start[d] >= 0
                start_voxels.append(start)
                end_voxels.append(end)
    return start_voxels, end_voxels


def resample_out_size(in_size, in_spacing, out_spacing, max_stride):
    """image_tools.py:363-366: output size of resample_spacing."""
    out = [int(in_size[i] * in_spacing[i] / out_spacing[i] + 0.5) for i in range(3)]
    for i in range(3):
        if out[i] % max_stride:
            out[i] = max_stride * (out[i] // max_stride + 1)
    return out


def normalize_fixed(vol, mean, stddev, clip, clip_min=-1.0, clip_max=1.0):
    """normalizer.py:22-25 -> image_tools.py:221-238, float32 numpy arithmetic."""
    a = np.array(vol, dtype=np.float32, copy=True)
    a = (a - mean) / stddev
    if clip:
        a[a < clip_min] = clip_min
        a[a > clip_max] = clip_max
    return a.astype(np.float32)


def normalize_adaptive(vol, clip_sigma):
    """normalizer.py:55-62: z-score with the crop's own np.mean/np.std, clip to +-clip_sigma."""
    a = np.asarray(vol, dtype=np.float32)
    m, s = np.mean(a), np.std(a)
    s = max(s, 1e-6)
    return normalize_fixed(a, m, s, True, -clip_sigma, clip_sigma)


def apply_normalizer(vol, normalizer):
    """normalizer = dict in the checkpoint's `crop_normalizers` format (normalizer.py:36-39,78-81)."""
    if normalizer is None:
        return np.asarray(vol, dtype=np.float32)
    if normalizer['type'] == 0:
        return normalize_fixed(vol, normalizer['mean'], normalizer['stddev'], normalizer['clip'])
    if normalizer['type'] == 1:
        return normalize_adaptive(vol, normalizer['clip_sigma'])
    raise ValueError('Unsupported normalization type.')   # core/seg_infer.py:162


def segmentation_volume(state_dict, vol_zyx, spacing, normalizer, partition_type='SIZE',
                        partition_size=(96, 96, 96), partition_stride=(96, 96, 96), max_stride=16,
                        bbox_start_voxel=None, bbox_end_voxel=None,
                        double_forward=True, faithful_copies=True, max_patches=None, forward_fn=None):
    """core/seg_infer.py:249-339 for an input already at the model spacing (resample == identity:
    size % max_stride == 0 and spacing == model spacing, the BASELINE configs).

    Returns (mean_probs [C,z,y,x] float32, mask [z,y,x] int8, start_voxels, end_voxels).
    double_forward: run the net twice per patch and average, as :230-234 does.
    faithful_copies: add_image_region / add_image_value copy the whole accumulator in and out of
      a fresh array per call (image_tools.py:446-449,464-466) - kept so CPU-baseline timings carry
      the reference's host cost.  Values are identical either way.
    max_patches: process only the first k patches (bounded CPU-baseline sample); the result is
      then only valid inside those patches.
    """
    vol = np.asarray(vol_zyx, dtype=np.float32)
    size_xyz = [int(vol.shape[2]), int(vol.shape[1]), int(vol.shape[0])]
    if partition_type == 'DISABLE':
        starts, ends = [[0, 0, 0]], [list(size_xyz)]                                       # :277-279
    elif partition_type == 'SIZE':
        if bbox_start_voxel is None or bbox_end_voxel is None:
            bs, be = [0, 0, 0], list(size_xyz)                                             # :304
        else:
            bs = [max(0, int(v)) for v in bbox_start_voxel]
            be = [min(int(bbox_end_voxel[i]), size_xyz[i]) for i in range(3)]
        starts, ends = partition_grid(size_xyz, spacing, bs, be, copy.deepcopy(list(partition_size)),
                                      copy.deepcopy(list(partition_stride)), max_stride)
    else:
        raise ValueError('Unsupported partition type!')                                    # :311

    fwd = forward_fn if forward_fn is not None else (lambda t: onet.forward(state_dict, t))
    acc = None
    count = np.zeros(vol.shape, dtype=np.float32)
    n_done = 0
    for s, e in zip(starts, ends):
        if max_patches is not None and n_done >= max_patches:
            break
        roi = np.ascontiguousarray(vol[s[2]:e[2], s[1]:e[1], s[0]:e[0]])                   # :221 (xyz slicing -> own buffer)
        roi = apply_normalizer(roi, normalizer)                                            # :223-224
        t = torch.from_numpy(np.ascontiguousarray(roi)).unsqueeze(0).unsqueeze(0).float()  # :226
        p = fwd(t)
        if double_forward:
            p2 = fwd(t)
            p = torch.mean(torch.cat((p.unsqueeze(0), p2.unsqueeze(0)), 0), 0)             # :231-234
        p = p[0].numpy()
        if acc is None:
            acc = [np.zeros(vol.shape, dtype=np.float32) for _ in range(p.shape[0])]
        for c in range(p.shape[0]):
            if faithful_copies:
                a = acc[c].copy()                                                          # GetArrayFromImage
                a[s[2]:e[2], s[1]:e[1], s[0]:e[0]] += p[c]                                  # image_tools.py:448
                acc[c] = a.copy()                                                          # GetImageFromArray
            else:
                acc[c][s[2]:e[2], s[1]:e[1], s[0]:e[0]] += p[c]
        if faithful_copies:
            a = count.copy()
            a[s[2]:e[2], s[1]:e[1], s[0]:e[0]] += 1.0                                       # image_tools.py:465
            count = a.copy()
        else:
            count[s[2]:e[2], s[1]:e[1], s[0]:e[0]] += 1.0
        n_done += 1

    with np.errstate(divide='ignore'):
        recip = (np.float32(1.0) / count).astype(np.float32)                               # :325
    probs = np.stack([a * recip for a in acc], 0).astype(np.float32)                       # :326-327
    if max_patches is not None:
        probs = np.nan_to_num(probs, nan=0.0, posinf=0.0, neginf=0.0)
    mask = argmax_first(probs)                                                             # :336-338
    return probs, mask, starts, ends


def argmax_first(probs):
    """core/seg_infer.py:337: `tensor.max(0)` index, lowest class index on ties; int8."""
    return torch.from_numpy(np.ascontiguousarray(probs)).max(0)[1].numpy().astype(np.int8)


def overlap_count_axes(size_xyz, starts, ends):
    """Rasterised overlap count (what add_image_value accumulates, image_tools.py:455-469); used
    to check the product's separable per-axis count (the grid is a Cartesian product of per-axis
    box lists, clamping included, so count(x,y,z) = cx(x)*cy(y)*cz(z))."""
    cnt = np.zeros((size_xyz[2], size_xyz[1], size_xyz[0]), dtype=np.float32)
    for s, e in zip(starts, ends):
        cnt[s[2]:e[2], s[1]:e[1], s[0]:e[0]] += 1.0
    return cnt


def segmentation_volume_resampled(state_dict, vol_zyx, image_spacing, model_spacing, normalizer, interpolation='LINEAR',
                                  max_stride=16, **kw):
    """core/seg_infer.py:262-339 for a scan whose spacing differs from the model's: resample to the model spacing (:267),
    segment there, resample every class map back to the scan's grid with padding 1.0 for class 0 and 0.0 otherwise
    (:329-333), first-argmax on the scan's grid (:336-338).  Returns (mean_probs [C,z,y,x], mask [z,y,x] int8)."""
    from oracle import resample as orz
    vol = np.asarray(vol_zyx, dtype=np.float32)
    in_size = [vol.shape[2], vol.shape[1], vol.shape[0]]
    iso, _ = orz.resample_spacing(vol, image_spacing, model_spacing, max_stride, interpolation)
    probs_iso, _, _, _ = segmentation_volume(state_dict, iso, model_spacing, normalizer, max_stride=max_stride, **kw)
    back = np.stack([orz.resample_grid(probs_iso[c], model_spacing, in_size, image_spacing, 'LINEAR', 1.0 if c == 0 else 0.0)
                     for c in range(probs_iso.shape[0])], 0)
    return back, argmax_first(back)
