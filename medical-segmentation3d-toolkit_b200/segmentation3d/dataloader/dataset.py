"""Training crops (drop-in for reference dataloader/dataset.py:12-209).

The data loader is the caller on the near side of the training path.  Same constructor, same item format -
(image [1,D,H,W] f32, mask [1,D,H,W] f32, frame [15] = spacing, origin, direction, case name) - and the same sampling,
drawing from numpy's global RNG in the same order as the reference, so a seeded run picks the same crops:
CENTER / GLOBAL / MASK / HYBRID centre (:116-173), random translation in mm (:176), random isotropic rescale of the crop
spacing (:179), linear crop of the image and nearest-neighbour crop of the mask through `crop_image` (:182-188).
Like the reference's SimpleITK calls, the crop runs on the host inside DataLoader workers; images are read with the
MetaImage reader of utils/image3d.py (SimpleITK formats when that package is installed).
"""
import collections
import os

import numpy as np
from torch.utils.data import Dataset

from segmentation3d.utils.file_io import readlines
from segmentation3d.utils.image3d import read_image
from segmentation3d.utils.image_tools import (convert_image_to_tensor, crop_image, get_image_frame,
                                              select_random_voxels_in_multi_class_mask)


def read_train_txt(imlist_file):
    """first line = N, then N (image path, mask path) line pairs (reference dataset.py:12-33)."""
    lines = readlines(imlist_file)
    num_cases = int(lines[0])
    if len(lines) - 1 < num_cases * 2:
        raise ValueError('too few lines in imlist file')
    im_list, seg_list = [], []
    for i in range(num_cases):
        im_path, seg_path = lines[1 + i * 2], lines[2 + i * 2]
        assert os.path.isfile(im_path), 'image not exist: {}'.format(im_path)
        assert os.path.isfile(seg_path), 'mask not exist: {}'.format(seg_path)
        im_list.append(im_path)
        seg_list.append(seg_path)
    return im_list, seg_list


def read_train_csv(imlist_file, mode='train'):
    """csv with image_name, image_path[, mask_path] columns (reference dataset.py:36-54)."""
    import pandas as pd
    df = pd.read_csv(imlist_file)
    if mode == 'test':
        return df['image_name'].tolist(), df['image_path'].tolist()
    if mode in ('train', 'validation'):
        return df['image_path'].tolist(), df['mask_path'].tolist()
    raise ValueError('Unsupported mode type.')


class _ImageCache(object):
    """decoded volumes kept per (worker) process, least recently used first out, bounded in bytes.  The reference re-reads
    and decompresses both files for every crop; the crops drawn are the same either way."""

    def __init__(self, max_bytes):
        self.max_bytes, self.bytes, self.items = max_bytes, 0, collections.OrderedDict()
        self.hits = self.misses = 0

    def get(self, path, dtype):
        key = (path, np.dtype(dtype).str)
        if key in self.items:
            self.items.move_to_end(key)
            self.hits += 1
            return self.items[key]
        self.misses += 1
        img = read_image(path, dtype)
        n = int(img.to_numpy().nbytes)
        if n <= self.max_bytes:
            self.items[key] = img
            self.bytes += n
            while self.bytes > self.max_bytes:
                _, old = self.items.popitem(last=False)
                self.bytes -= int(old.to_numpy().nbytes)
        return img


SAMPLING_METHODS = ('CENTER', 'GLOBAL', 'MASK', 'HYBRID')


def _vector(value, length, dtype, what):
    v = np.array(value, dtype=dtype)
    assert v.size == length, 'only %d-element of %s is supported' % (length, what)
    return v


class SegmentationDataset(Dataset):
    """Constructor arguments as the reference's (dataloader/dataset.py:69-101): list file (txt / csv), class count, crop
    spacing [mm], crop size [voxels, x y z], sampling method, +-translation [mm], (lo, hi) isotropic rescale, image
    interpolation, one normaliser (or None) per modality."""

    def __init__(self, imlist_file, num_classes, spacing, crop_size, sampling_method, random_translation,
                 random_scale, interpolation, crop_normalizers):
        readers = {'txt': read_train_txt, 'csv': read_train_csv}
        suffix = imlist_file[-3:]
        if suffix not in readers:
            raise ValueError('imseg_list must be a txt file')
        self.im_list, self.seg_list = readers[suffix](imlist_file)
        self.num_classes = num_classes
        self.spacing = _vector(spacing, 3, np.double, 'spacing')
        self.crop_size = _vector(crop_size, 3, np.int32, 'crop size')
        self.random_translation = _vector(random_translation, 3, np.double, 'random translation')
        self.random_scale = _vector(random_scale, 2, np.double, 'random scale')
        assert sampling_method in SAMPLING_METHODS, 'sampling_method must be one of %s' % (SAMPLING_METHODS,)
        assert interpolation in ('LINEAR', 'NN'), 'interpolation must be LINEAR or NN'
        assert isinstance(crop_normalizers, list), 'crop normalizers must be a list'
        self.sampling_method, self.interpolation, self.crop_normalizers = sampling_method, interpolation, crop_normalizers
        # SEG3D_DATASET_CACHE_MB: decoded volumes kept per process (0 = re-read every item like the reference)
        self._cache = _ImageCache(int(float(os.environ.get('SEG3D_DATASET_CACHE_MB', '2048')) * (1 << 20)))

    def __len__(self):
        return len(self.im_list)

    def num_modality(self):
        return 1

    def global_sample(self, image):
        """Uniformly random crop centre [world mm] with the crop inside the image along every axis where the image is the
        larger of the two; along the other axes the crop starts at the image origin (:108-124).  One np.random.uniform
        draw per free axis, x first - the reference's RNG order."""
        start = np.array(image.GetOrigin(), dtype=np.double)
        extent_mm = np.array(image.GetSize(), dtype=np.double) * np.array(image.GetSpacing(), dtype=np.double)
        crop_mm = self.crop_size * self.spacing
        for axis in np.flatnonzero(extent_mm > crop_mm):
            start[axis] += np.random.uniform(0, extent_mm[axis] - crop_mm[axis])
        return start + crop_mm / 2

    def center_sample(self, image):
        """world coordinate of the image centre (:126-138)."""
        origin = image.GetOrigin()
        end_point_world = image.TransformContinuousIndexToPhysicalPoint([float(image.GetSize()[idx] - 1) for idx in range(3)])
        return np.array([(origin[idx] + end_point_world[idx]) / 2.0 for idx in range(3)], dtype=np.double)

    def _mask_sample(self, seg):
        centers = select_random_voxels_in_multi_class_mask(seg, 1, np.random.randint(1, self.num_classes))
        if len(centers) > 0:
            return np.array(seg.TransformContinuousIndexToPhysicalPoint([float(int(centers[0][idx])) for idx in range(3)]))
        return self.global_sample(seg)      # no voxel of the drawn label

    def sample_crop(self, index, seg):
        """(centre [world mm], crop spacing) for item `index`: the reference's RNG call order (:156-179)."""
        if self.sampling_method == 'CENTER':
            center = self.center_sample(seg)
        elif self.sampling_method == 'GLOBAL':
            center = self.global_sample(seg)
        elif self.sampling_method == 'MASK':
            center = self._mask_sample(seg)
        elif self.sampling_method == 'HYBRID':
            center = self.global_sample(seg) if index % 2 else self._mask_sample(seg)
        else:
            raise ValueError('Only CENTER, GLOBAL, MASK and HYBRID are supported as sampling methods')
        center = center + np.random.uniform(-self.random_translation, self.random_translation, size=[3])
        crop_spacing = self.spacing * np.random.uniform(self.random_scale[0], self.random_scale[1])
        return center, crop_spacing

    def __getitem__(self, index):
        image_path, seg_path = self.im_list[index], self.seg_list[index]
        case_name = os.path.basename(os.path.dirname(image_path)) + '_' + os.path.basename(image_path)
        images = [self._cache.get(image_path, np.float32)]
        seg = self._cache.get(seg_path, np.float32)
        center, crop_spacing = self.sample_crop(index, seg)
        for idx in range(len(images)):
            images[idx] = crop_image(images[idx], center, self.crop_size, crop_spacing, self.interpolation)
            if self.crop_normalizers[idx] is not None:
                images[idx] = self.crop_normalizers[idx](images[idx])
        seg = crop_image(seg, center, self.crop_size, crop_spacing, 'NN')
        frame = get_image_frame(seg)
        return convert_image_to_tensor(images), convert_image_to_tensor(seg), frame, case_name
