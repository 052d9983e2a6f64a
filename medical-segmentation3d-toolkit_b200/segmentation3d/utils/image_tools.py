"""Host-side image helpers on the inference path (drop-in subset of reference utils/image_tools.py).

Bit-exact integer pieces: `image_partition_by_fixed_size` (image_tools.py:163-218) and the output
size rounding of `resample_spacing` (:363-366).  The array work the reference does here on host
numpy (normalise :221-238, add_image_region/add_image_value :435-469, argmax) runs in CUDA kernels
in this build (segmentation3d/_b200/sliding.py); the host versions below exist for API parity on
small images and are not on the engine's path.
"""
import os

import numpy as np
import torch

from segmentation3d.utils.image3d import Image3d, as_image3d


def image_partition_by_fixed_size(image, bbox_start_voxel, bbox_end_voxel, partition_size, partition_stride, max_stride):
    """Sliding-window lattice.  Returns (start_voxels, end_voxels): lists of [x,y,z], x outermost,
    z innermost; the last box per axis is clamped back inside the (max_stride-rounded) bounding box.
    Like the reference, the two bbox list arguments are updated in place."""
    size, spacing = image.GetSize(), image.GetSpacing()
    for a in range(3):
        assert size[a] >= max_stride and size[a] % max_stride == 0
    extent = [0, 0, 0]
    for a in range(3):
        e = min(size[a], bbox_end_voxel[a] - bbox_start_voxel[a])
        if e % max_stride:
            e = (e // max_stride + 1) * max_stride
        extent[a] = min(e, size[a])
        bbox_end_voxel[a] = bbox_start_voxel[a] + extent[a]
        if bbox_end_voxel[a] > size[a]:
            bbox_end_voxel[a] = size[a]
            bbox_start_voxel[a] = size[a] - extent[a]
        assert bbox_start_voxel[a] >= 0
    box, step, count = [0, 0, 0], [0, 0, 0], [0, 0, 0]
    for a in range(3):
        b = int(partition_size[a] / spacing[a] + 0.5)
        if b % max_stride:
            b = (b // max_stride + 1) * max_stride
        box[a] = min(extent[a], b)
        step[a] = min(extent[a], int(partition_stride[a] / spacing[a] + 0.5))
        count[a] = int(np.ceil((extent[a] - box[a]) / step[a]) + 1)
    axis_starts = []
    for a in range(3):
        lst = []
        for i in range(count[a]):
            s = bbox_start_voxel[a] + i * step[a]
            if s + box[a] > bbox_end_voxel[a]:
                s = bbox_end_voxel[a] - box[a]
                assert s >= 0
            lst.append(s)
        axis_starts.append(lst)
    starts = [[sx, sy, sz] for sx in axis_starts[0] for sy in axis_starts[1] for sz in axis_starts[2]]
    ends = [[s[0] + box[0], s[1] + box[1], s[2] + box[2]] for s in starts]
    return starts, ends


def resample_size(in_size, in_spacing, out_spacing, max_stride):
    """Output size of resample_spacing: round(size*spacing_in/spacing_out), then up to a multiple of max_stride."""
    out = []
    for a in range(3):
        n = int(in_size[a] * in_spacing[a] / out_spacing[a] + 0.5)
        if n % max_stride:
            n = (n // max_stride + 1) * max_stride
        out.append(n)
    return out


def is_identity_resample(image, spacing, max_stride):
    size = image.GetSize()
    same = all(abs(float(image.GetSpacing()[a]) - float(spacing[a])) <= 1e-9 * max(1.0, abs(float(spacing[a]))) for a in range(3))
    return same and resample_size(size, image.GetSpacing(), spacing, max_stride) == list(size)


def _resample_device(src, in_spacing, out_size_xyz, out_spacing, interp_method, padding_value, device=None):
    """fp32 [z,y,x] tensor on `in_spacing` -> CUDA fp32 [z,y,x] of out_size on out_spacing (same origin / direction)."""
    from segmentation3d._b200 import lib
    if interp_method not in ('LINEAR', 'NN'):
        raise ValueError('Unsupported interpolation type.')
    if device is None:
        device = src.device if (torch.is_tensor(src) and src.is_cuda) else torch.device('cuda')
    t = src if torch.is_tensor(src) else torch.from_numpy(np.ascontiguousarray(src))
    t = t.to(device=device, dtype=torch.float32).contiguous()
    if t.device.type != 'cuda':
        raise RuntimeError('seg3d_b200: resampling runs on CUDA devices only (no CPU fallback)')
    dx, dy, dz = [int(v) for v in out_size_xyz]
    out = torch.empty((dz, dy, dx), dtype=torch.float32, device=t.device)
    r = [float(out_spacing[a]) / float(in_spacing[a]) for a in range(3)]          # x, y, z
    with torch.cuda.device(t.device):
        lib.call('seg3d_resample', lib.ptr(t), t.shape[0], t.shape[1], t.shape[2], lib.ptr(out), dz, dy, dx,
                 r[2], r[1], r[0], 1 if interp_method == 'LINEAR' else 0, float(padding_value), lib.stream_ptr())
    return out


def resample(image, reference, interp_method, padding_value=0.0):
    """Resample `image` onto the grid of `reference` (utils/image_tools.py:329-343).  Both images must share origin and
    direction, which is how the engine uses it (core/seg_infer.py:330-333)."""
    img, ref = as_image3d(image), as_image3d(reference)
    if (max(abs(a - b) for a, b in zip(img.GetOrigin(), ref.GetOrigin())) > 1e-6 or
            max(abs(a - b) for a, b in zip(img.GetDirection(), ref.GetDirection())) > 1e-9):
        raise ValueError('resample: image and reference must share origin and direction')
    out = _resample_device(img.data, img.GetSpacing(), ref.GetSize(), ref.GetSpacing(), interp_method, padding_value)
    return Image3d(out, ref.GetSpacing(), ref.GetOrigin(), ref.GetDirection())


def resample_spacing(image, resampled_spacing, max_stride, interp_method):
    """Resample to `resampled_spacing`; the output size is rounded up to a multiple of max_stride
    (utils/image_tools.py:346-377)."""
    img = as_image3d(image)
    osz = resample_size(img.GetSize(), img.GetSpacing(), resampled_spacing, max_stride)
    out = _resample_device(img.data, img.GetSpacing(), osz, resampled_spacing, interp_method, 0.0)
    return Image3d(out, [float(v) for v in resampled_spacing], img.GetOrigin(), img.GetDirection())


def get_image_frame(image):
    """[spacing(3), origin(3), direction(9)] as float32 (utils/image_tools.py:24-39)."""
    img = as_image3d(image)
    return np.array(list(img.GetSpacing()) + list(img.GetOrigin()) + list(img.GetDirection()), dtype=np.float32)


def set_image_frame(image, frame):
    """inverse of get_image_frame (utils/image_tools.py:42-59)."""
    frame = np.asarray(frame)
    image.SetSpacing(frame[:3].astype(np.double))
    image.SetOrigin(frame[3:6].astype(np.double))
    image.SetDirection(frame[6:15].astype(np.double))


def select_random_voxels_in_multi_class_mask(mask, num_selected, selected_label):
    """`num_selected` random [x,y,z] voxels carrying `selected_label`, with replacement; [] when the label is absent
    (utils/image_tools.py:252-271; one np.random.randint per selected voxel)."""
    valid = np.argwhere(as_image3d(mask).to_numpy() == selected_label)
    selected = []
    while len(valid) > 0 and len(selected) < num_selected:
        selected.append(valid[np.random.randint(0, len(valid))][::-1])
    return selected


def crop_geometry(cropping_center, cropping_size, cropping_spacing):
    """origin (world position of the first voxel's centre) of a crop of `cropping_size` voxels at `cropping_spacing`
    centred on `cropping_center` - along the world axes, as the reference computes it (utils/image_tools.py:127-133)."""
    center = [float(v) for v in cropping_center]
    size = [int(v) for v in cropping_size]
    spacing = [float(v) for v in cropping_spacing]
    return [center[a] - size[a] * spacing[a] / 2.0 + spacing[a] / 2.0 for a in range(3)], size, spacing


def crop_image(image, cropping_center, cropping_size, cropping_spacing, interp_method):
    """Crop a patch around a world-coordinate centre at a given voxel spacing (utils/image_tools.py:111-146:
    sitk.Resample with an identity transform onto a grid that keeps the volume's direction; pixels outside the volume
    are 0).  ITK semantics restated: output voxel i sits at origin_crop + D (i * spacing_crop), its continuous input index
    is c = D^-1 (p - origin) / spacing in double precision; inside means -0.5 <= c < size - 0.5 per axis; 'LINEAR' blends
    the 8 neighbours with both neighbours clamped to the volume, 'NN' takes floor(c + 0.5).
    Runs on the host (numpy): like the reference's SimpleITK call it executes inside DataLoader worker processes."""
    if interp_method not in ('LINEAR', 'NN'):
        raise ValueError('Unsupported interpolation type.')
    img = as_image3d(image)
    if img.is_cuda():
        return crop_image_device(img, cropping_center, cropping_size, cropping_spacing, interp_method)
    src = img.to_numpy()
    origin_c, size, spacing_c = crop_geometry(cropping_center, cropping_size, cropping_spacing)
    d = np.asarray(img.GetDirection(), dtype=np.float64).reshape(3, 3)
    sp = np.asarray(img.GetSpacing(), dtype=np.float64)
    off = np.linalg.solve(d, np.asarray(origin_c, dtype=np.float64) - np.asarray(img.GetOrigin(), dtype=np.float64)) / sp
    n_in = [src.shape[2], src.shape[1], src.shape[0]]
    coords = [off[a] + np.arange(size[a], dtype=np.float64) * (spacing_c[a] / sp[a]) for a in range(3)]     # x, y, z
    ok = [(coords[a] >= -0.5) & (coords[a] < n_in[a] - 0.5) for a in range(3)]
    inside = ok[2][:, None, None] & ok[1][None, :, None] & ok[0][None, None, :]
    if interp_method == 'NN':
        idx = [np.clip(np.floor(coords[a] + 0.5).astype(np.int64), 0, n_in[a] - 1) for a in range(3)]
        val = src[np.ix_(idx[2], idx[1], idx[0])]
    else:
        lo, hi, w = [], [], []
        for a in range(3):
            f = np.floor(coords[a])
            lo.append(np.clip(f.astype(np.int64), 0, n_in[a] - 1))
            hi.append(np.clip(f.astype(np.int64) + 1, 0, n_in[a] - 1))
            w.append(coords[a] - f)
        # only the block of the volume the crop touches is converted to float64
        base = [int(min(lo[a].min(), hi[a].min())) for a in range(3)]
        top = [int(max(lo[a].max(), hi[a].max())) + 1 for a in range(3)]
        s64 = src[base[2]:top[2], base[1]:top[1], base[0]:top[0]].astype(np.float64)
        lo = [lo[a] - base[a] for a in range(3)]
        hi = [hi[a] - base[a] for a in range(3)]
        # trilinear = three 1-D interpolations (x, then y, then z): same weights, a third of the gathers
        ax = s64[:, :, lo[0]] * (1.0 - w[0])[None, None, :] + s64[:, :, hi[0]] * w[0][None, None, :]
        ay = ax[:, lo[1], :] * (1.0 - w[1])[None, :, None] + ax[:, hi[1], :] * w[1][None, :, None]
        val = ay[lo[2]] * (1.0 - w[2])[:, None, None] + ay[hi[2]] * w[2][:, None, None]
    out = np.where(inside, val, 0).astype(src.dtype)
    return Image3d(out, spacing_c, origin_c, img.GetDirection())


def percentiles(image, percentiles):
    """np.percentile of the voxel values (utils/image_tools.py:241-249)."""
    return np.percentile(as_image3d(image).to_numpy(), percentiles)


def copy_image(source_image, target_start_voxel, target_end_voxel, target_image):
    """Paste the region of `source_image` that lies at target voxels [start, end) into a copy of `target_image`; both
    images share their orientation (utils/image_tools.py:149-166: sitk.Paste).  The bbox lists are cast to int in place."""
    src, tgt = as_image3d(source_image), as_image3d(target_image)
    for idx in range(3):
        target_start_voxel[idx] = int(target_start_voxel[idx])
        target_end_voxel[idx] = int(target_end_voxel[idx])
    start_world = tgt.TransformContinuousIndexToPhysicalPoint([float(v) for v in target_start_voxel])
    s = src.TransformPhysicalPointToIndex(start_world)
    size = [int(target_end_voxel[idx] - target_start_voxel[idx]) for idx in range(3)]
    t = target_start_voxel
    out = np.array(tgt.to_numpy(), copy=True)
    out[t[2]:t[2] + size[2], t[1]:t[1] + size[1], t[0]:t[0] + size[0]] = \
        src.to_numpy()[s[2]:s[2] + size[2], s[1]:s[1] + size[1], s[0]:s[0] + size[0]]
    res = Image3d(out)
    res.CopyInformation(tgt)
    return res


def get_bounding_box(mask, selected_labels):
    """[x,y,z] start (inclusive) / end (exclusive) of the voxels carrying one of `selected_labels` (all non-zero labels
    when None); (None, None) for an empty selection (utils/image_tools.py:481-510: LabelShapeStatisticsImageFilter).
    Works on host arrays and on CUDA masks alike."""
    m = as_image3d(mask).data
    m = m if torch.is_tensor(m) else torch.from_numpy(np.ascontiguousarray(m))
    sel = (m > 0) if selected_labels is None else torch.isin(m, torch.tensor(list(selected_labels), device=m.device, dtype=m.dtype))
    if not bool(sel.any()):
        print('Fail to get the bounding box.')
        return None, None
    # per-axis occupancy (three small reductions) instead of materialising the coordinates of every selected voxel
    lo, hi = [], []
    for axis in (2, 1, 0):                                      # x, y, z
        other = tuple(a for a in range(3) if a != axis)
        idx = sel.any(dim=other[1]).any(dim=other[0]).nonzero().flatten()
        lo.append(int(idx[0]))
        hi.append(int(idx[-1]) + 1)
    return lo, hi


def save_intermediate_results(idxs, crops, masks, outputs, frames, file_names, out_folder):
    """Write the crops, masks and network outputs of the batch items `idxs` for inspection (utils/image_tools.py:62-108):
    <out_folder>/<file_name>/batch_<i>_crop_<m>.nii.gz, batch_<i>_mask.nii.gz, batch_<i>_output_<c>.nii.gz, each with the
    crop's frame."""
    from segmentation3d.utils.image3d import write_image
    if not os.path.isdir(out_folder):
        os.makedirs(out_folder)
    for i in idxs:
        case_out_folder = os.path.join(out_folder, file_names[i])
        if not os.path.isdir(case_out_folder):
            os.makedirs(case_out_folder)
        frame = frames[i].numpy() if torch.is_tensor(frames[i]) else np.asarray(frames[i])
        if crops is not None:
            for modality_idx, image in enumerate(convert_tensor_to_image(crops[i], dtype=np.float32)):
                set_image_frame(image, frame)
                write_image(image, os.path.join(case_out_folder, 'batch_{}_crop_{}.nii.gz'.format(i, modality_idx)))
        if masks is not None:
            mask = convert_tensor_to_image(masks[i, 0], dtype=np.int32)
            set_image_frame(mask, frame)
            write_image(mask, os.path.join(case_out_folder, 'batch_{}_mask.nii.gz'.format(i)))
        if outputs is not None:
            for cls in range(outputs.size()[1]):
                output = convert_tensor_to_image(outputs[i, cls].data, dtype=np.float32)
                set_image_frame(output, frame)
                write_image(output, os.path.join(case_out_folder, 'batch_{}_output_{}.nii.gz'.format(i, cls)))


def crop_image_device(image, cropping_center, cropping_size, cropping_spacing, interp_method, out=None):
    """crop_image for a volume resident on the GPU (seg3d_crop_resample): same geometry and ITK semantics as the host
    version above; `out` (float32 CUDA tensor [z,y,x], e.g. one item of a batch buffer) receives the crop when given.
    Not yet run on a GPU (added after round 1's GPU budget was spent); wiring pinned in tests/test_device_crops_wiring.py."""
    from segmentation3d._b200 import lib
    if interp_method not in ('LINEAR', 'NN'):
        raise ValueError('Unsupported interpolation type.')
    img = as_image3d(image)
    if not img.is_cuda():
        raise RuntimeError('crop_image_device needs a CUDA-resident image (crop_image handles host images)')
    src = img.data
    if src.dtype != torch.float32 or not src.is_contiguous():
        src = src.float().contiguous()
    origin_c, size, spacing_c = crop_geometry(cropping_center, cropping_size, cropping_spacing)
    d = np.asarray(img.GetDirection(), dtype=np.float64).reshape(3, 3)
    sp = np.asarray(img.GetSpacing(), dtype=np.float64)
    off = np.linalg.solve(d, np.asarray(origin_c, dtype=np.float64) - np.asarray(img.GetOrigin(), dtype=np.float64)) / sp
    if out is None:
        out = torch.empty((size[2], size[1], size[0]), dtype=torch.float32, device=src.device)
    assert out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (size[2], size[1], size[0])
    Z, Y, X = src.shape
    with torch.cuda.device(src.device):
        lib.call('seg3d_crop_resample', lib.ptr(src), Z, Y, X, lib.ptr(out), size[2], size[1], size[0],
                 float(off[2]), float(off[1]), float(off[0]), float(spacing_c[2] / sp[2]), float(spacing_c[1] / sp[1]),
                 float(spacing_c[0] / sp[0]), 1 if interp_method == 'LINEAR' else 0, 0.0, lib.stream_ptr())
    return Image3d(out, spacing_c, origin_c, img.GetDirection())


def pick_largest_connected_component(mask, labels):
    """Keep, per label, the largest 26-connected component (utils/image_tools.py:380-404)."""
    from segmentation3d.core.seg_infer import _cc_filter_device
    img = as_image3d(mask)
    data = img.data if torch.is_tensor(img.data) else torch.from_numpy(np.ascontiguousarray(img.data))
    out = _cc_filter_device(data.to(device='cuda', dtype=torch.int8), list(labels), 0)
    return Image3d(out, img.GetSpacing(), img.GetOrigin(), img.GetDirection())


def remove_small_connected_component(mask, labels, threshold):
    """Drop, per label, the 26-connected components smaller than `threshold` voxels (utils/image_tools.py:407-432)."""
    from segmentation3d.core.seg_infer import _cc_filter_device
    img = as_image3d(mask)
    data = img.data if torch.is_tensor(img.data) else torch.from_numpy(np.ascontiguousarray(img.data))
    out = _cc_filter_device(data.to(device='cuda', dtype=torch.int8), list(labels), int(threshold))
    return Image3d(out, img.GetSpacing(), img.GetOrigin(), img.GetDirection())


def convert_image_to_tensor(image):
    """Image (or list of images) -> float tensor [1,z,y,x] ([n,z,y,x] for a list)."""
    if isinstance(image, (list, tuple)):
        return torch.cat([convert_image_to_tensor(im) for im in image], 0)
    img = as_image3d(image)
    data = img.data if torch.is_tensor(img.data) else torch.from_numpy(np.ascontiguousarray(img.data))
    return data.unsqueeze(0).float()


def convert_tensor_to_image(tensor, dtype=None):
    """3-D tensor -> Image3d; 4-D tensor -> list of Image3d."""
    assert isinstance(tensor, torch.Tensor), 'input must be a tensor'
    if tensor.dim() == 4:
        return [convert_tensor_to_image(t, dtype) for t in tensor]
    if tensor.dim() != 3:
        raise ValueError('Only supports 3-dimsional or 4-dimensional image volume')
    arr = tensor.detach().cpu().numpy()
    if dtype is not None:
        arr = arr.astype({float: np.float32, int: np.int32}.get(dtype, dtype))
    return Image3d(arr)


def add_image_region(image, start_voxel, end_voxel, patch):
    """image[start:end] += patch on the host (API parity; the engine blends on the device)."""
    img, pat = as_image3d(image), as_image3d(patch)
    s, e = [int(v) for v in start_voxel], [int(v) for v in end_voxel]
    for a in range(3):
        assert pat.GetSize()[a] == e[a] - s[a]
    out = np.array(img.to_numpy(), copy=True)
    out[s[2]:e[2], s[1]:e[1], s[0]:e[0]] += pat.to_numpy()
    res = Image3d(out)
    res.CopyInformation(img)
    return res


def add_image_value(image, start_voxel, end_voxel, value):
    img = as_image3d(image)
    s, e = [int(v) for v in start_voxel], [int(v) for v in end_voxel]
    out = np.array(img.to_numpy(), copy=True)
    out[s[2]:e[2], s[1]:e[1], s[0]:e[0]] += value
    res = Image3d(out)
    res.CopyInformation(img)
    return res


def normalize_image(image, mean, std, clip, clip_min=-1.0, clip_max=1.0):
    """(v-mean)/std with optional clipping, float32 numpy (host; API parity)."""
    img = as_image3d(image)
    a = (np.asarray(img.to_numpy(), dtype=np.float32) - mean) / std
    if clip:
        np.clip(a, clip_min, clip_max, out=a)
    res = Image3d(a.astype(np.float32))
    res.CopyInformation(img)
    return res


def get_mean_std_from_image(image):
    a = as_image3d(image).to_numpy()
    return np.mean(a), np.std(a)
