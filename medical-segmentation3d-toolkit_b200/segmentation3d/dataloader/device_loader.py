"""Training crops drawn on the GPU (opt-in replacement for DataLoader(SegmentationDataset) in core/seg_train.py:68-70).

The reference crops on CPU workers with SimpleITK; here the decoded volumes stay resident in HBM (180 GB holds whole
training sets) and every batch is produced by a handful of launches: seg3d_crop_resample per item (linear for the image,
nearest for the mask) into one batch buffer, then the crop normaliser over the whole batch (seg3d_patch_stats /
seg3d_patch_gather, the kernels of the inference path).  The sampling decisions - crop centre, random translation, random
rescale - are the dataset's own `sample_crop`, drawn from numpy's RNG in the reference's order, so a seeded run sees the
same crops as the DataLoader path.  Items come back as (crops [B,1,D,H,W] f32, masks [B,1,D,H,W] f32, frames [B,15],
names) with crops and masks already on the device.

STATUS: opt-in (SEG3D_DEVICE_CROPS=1).  Written after round 1's GPU budget was spent: not yet run on a GPU; the host wiring
is pinned on the CPU against the DataLoader path with the entry points emulated (tests/test_device_crops_wiring.py).
"""
import collections
import os

import numpy as np
import torch

from segmentation3d._b200 import lib
from segmentation3d.utils.image3d import Image3d
from segmentation3d.utils.image_tools import crop_image_device, get_image_frame


class DeviceCropLoader(object):
    def __init__(self, dataset, sampler, batch_size, device=None, cache_gb=None):
        self.dataset, self.sampler, self.batch_size = dataset, sampler, int(batch_size)
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        gb = float(os.environ.get('SEG3D_DEVICE_CACHE_GB', '64')) if cache_gb is None else float(cache_gb)
        self.max_bytes, self.bytes, self.resident = int(gb * (1 << 30)), 0, collections.OrderedDict()

    def __len__(self):
        return (len(self.sampler) + self.batch_size - 1) // self.batch_size

    def _device_image(self, path, host):
        """the decoded volume as a float32 CUDA tensor, kept resident (least recently used first out)"""
        if path in self.resident:
            self.resident.move_to_end(path)
            return self.resident[path]
        t = torch.from_numpy(np.ascontiguousarray(host.to_numpy(), dtype=np.float32)).to(self.device)
        img = Image3d(t, host.GetSpacing(), host.GetOrigin(), host.GetDirection())
        n = t.numel() * 4
        if n <= self.max_bytes:
            self.resident[path] = img
            self.bytes += n
            while self.bytes > self.max_bytes:
                _, old = self.resident.popitem(last=False)
                self.bytes -= old.data.numel() * 4
        return img

    def _normalise(self, raw, crops):
        """crop normaliser over the whole batch: raw [B,D,H,W] f32 -> crops [B,1,D,H,W] f32"""
        norm = self.dataset.crop_normalizers[0]
        B, D, H, W = raw.shape
        if norm is None:
            crops.view(B, D, H, W).copy_(raw)
            return
        nd = norm.to_dict()
        kind, mean, std, clip, lo, hi = lib.NORM_FIXED, 0.0, 1.0, 0, -1.0, 1.0
        if nd['type'] == 0:
            mean, std, clip = float(nd['mean']), float(nd['stddev']), 1 if nd['clip'] else 0
        elif nd['type'] == 1:
            kind, clip, lo, hi = lib.NORM_ADAPTIVE, 1, -float(nd['clip_sigma']), float(nd['clip_sigma'])
        else:
            raise ValueError('Unsupported normalization type.')
        starts = torch.tensor([[0, 0, b * D] for b in range(B)], dtype=torch.int32, device=raw.device)
        stats = None
        with torch.cuda.device(raw.device):
            if kind == lib.NORM_ADAPTIVE:
                stats = torch.zeros((B, 2), dtype=torch.float64, device=raw.device)
                lib.call('seg3d_patch_stats', lib.ptr(raw), B * D, H, W, lib.ptr(starts), B, D, H, W, lib.ptr(stats), lib.stream_ptr())
            lib.call('seg3d_patch_gather', lib.ptr(raw), B * D, H, W, lib.ptr(starts), B, D, H, W, kind, mean, std, clip, lo, hi,
                     lib.ptr(stats), lib.F32, lib.ptr(crops), lib.stream_ptr())

    def __iter__(self):
        ds = self.dataset
        cx, cy, cz = [int(v) for v in ds.crop_size]
        batch = []
        indices = list(self.sampler)
        for pos, index in enumerate(indices):
            batch.append(index)
            if len(batch) < self.batch_size and pos + 1 < len(indices):
                continue
            B = len(batch)
            raw = torch.empty((B, cz, cy, cx), dtype=torch.float32, device=self.device)
            crops = torch.empty((B, 1, cz, cy, cx), dtype=torch.float32, device=self.device)
            masks = torch.empty((B, 1, cz, cy, cx), dtype=torch.float32, device=self.device)
            frames, names = [], []
            for b, idx in enumerate(batch):
                image_path, seg_path = ds.im_list[idx], ds.seg_list[idx]
                names.append(os.path.basename(os.path.dirname(image_path)) + '_' + os.path.basename(image_path))
                seg_host = ds._cache.get(seg_path, np.float32)
                center, crop_spacing = ds.sample_crop(idx, seg_host)          # same RNG draws as SegmentationDataset.__getitem__
                image_dev = self._device_image(image_path, ds._cache.get(image_path, np.float32))
                seg_dev = self._device_image(seg_path, seg_host)
                crop_image_device(image_dev, center, ds.crop_size, crop_spacing, ds.interpolation, out=raw[b])
                seg_crop = crop_image_device(seg_dev, center, ds.crop_size, crop_spacing, 'NN', out=masks[b, 0])
                frames.append(get_image_frame(seg_crop))
            self._normalise(raw, crops)
            yield crops, masks, torch.from_numpy(np.stack(frames)), names
            batch = []
