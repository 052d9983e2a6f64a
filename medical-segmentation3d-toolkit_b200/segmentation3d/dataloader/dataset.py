"""Training crops (reduced drop-in for reference dataloader/dataset.py:54-209).

Host-side data feeding is outside the accelerated path: this class keeps the
reference's constructor and item format - (image [1,D,H,W] f32, mask [1,D,H,W] f32, frame [15], name) - and its
CENTER / GLOBAL / MASK / HYBRID centre sampling with random translation, on images that are already at the
training spacing (no SimpleITK resampling / random scaling in this build).
"""
import numpy as np
import torch
from torch.utils.data import Dataset

from segmentation3d.utils.file_io import readlines
from segmentation3d.utils.image3d import read_image


def read_train_txt(imlist_file):
    """first line = N, then N (image path, mask path) line pairs (reference dataset.py:12-33)."""
    lines = readlines(imlist_file)
    n = int(lines[0])
    if len(lines) - 1 < 2 * n:
        raise ValueError('too few lines in the training list')
    return [lines[1 + 2 * i] for i in range(n)], [lines[2 + 2 * i] for i in range(n)]


def read_train_csv(imlist_file):
    import pandas as pd
    df = pd.read_csv(imlist_file)
    return df['image_path'].tolist(), df['mask_path'].tolist()


class SegmentationDataset(Dataset):
    def __init__(self, imlist_file, num_classes, spacing, crop_size, sampling_method, random_translation,
                 random_scale, interpolation, crop_normalizers):
        if imlist_file.endswith('.txt'):
            self.im_list, self.seg_list = read_train_txt(imlist_file)
        elif imlist_file.endswith('.csv'):
            self.im_list, self.seg_list = read_train_csv(imlist_file)
        else:
            raise ValueError('imseg_list must either be a txt file or a csv file')
        self.num_classes = num_classes
        self.spacing = np.array(spacing, dtype=np.double)
        self.crop_size = np.array(crop_size, dtype=np.int32)      # x, y, z
        assert sampling_method in ('CENTER', 'GLOBAL', 'MASK', 'HYBRID'), 'sampling_method must be CENTER, GLOBAL, MASK or HYBRID'
        self.sampling_method = sampling_method
        self.random_translation = np.array(random_translation, dtype=np.double)
        self.random_scale = random_scale
        assert interpolation in ('LINEAR', 'NN'), 'interpolation must either be a LINEAR or an NN'
        self.interpolation = interpolation
        self.crop_normalizers = crop_normalizers

    def __len__(self):
        return len(self.im_list)

    def num_modality(self):
        return 1

    def _centre(self, seg, size_xyz):
        method = self.sampling_method
        if method == 'HYBRID':
            method = 'GLOBAL' if np.random.randint(0, 2) == 0 else 'MASK'
        if method == 'CENTER':
            c = np.array(size_xyz, dtype=np.double) / 2
        elif method == 'MASK' and (seg > 0).any():
            zyx = np.argwhere(seg > 0)
            c = zyx[np.random.randint(0, len(zyx))][::-1].astype(np.double)
        else:
            c = np.array([np.random.uniform(0, s) for s in size_xyz])
        c += np.random.uniform(-self.random_translation, self.random_translation) / self.spacing
        return c

    def __getitem__(self, index):
        image = read_image(self.im_list[index], np.float32)
        mask = read_image(self.seg_list[index])
        if not np.allclose(image.GetSpacing(), self.spacing):
            raise NotImplementedError('this build trains on images already resampled to dataset.spacing')
        im, seg = image.to_numpy(), mask.to_numpy()
        size = image.GetSize()
        c = self._centre(seg, size)
        start = [int(np.clip(round(c[a] - self.crop_size[a] / 2), 0, max(0, size[a] - self.crop_size[a]))) for a in range(3)]
        cx, cy, cz = [int(v) for v in self.crop_size]
        crop = np.zeros((cz, cy, cx), np.float32)
        lab = np.zeros((cz, cy, cx), np.float32)
        sub = im[start[2]:start[2] + cz, start[1]:start[1] + cy, start[0]:start[0] + cx]
        crop[:sub.shape[0], :sub.shape[1], :sub.shape[2]] = sub
        sub = seg[start[2]:start[2] + cz, start[1]:start[1] + cy, start[0]:start[0] + cx]
        lab[:sub.shape[0], :sub.shape[1], :sub.shape[2]] = sub
        if self.crop_normalizers is not None:
            crop = self.crop_normalizers[0](crop_to_image(crop)).to_numpy().astype(np.float32)
        origin = [image.GetOrigin()[a] + start[a] * self.spacing[a] for a in range(3)]
        frame = np.array(list(origin) + list(self.spacing) + list(image.GetDirection()), dtype=np.float32)
        name = self.im_list[index]
        return torch.from_numpy(crop).unsqueeze(0), torch.from_numpy(lab).unsqueeze(0), torch.from_numpy(frame), name


def crop_to_image(arr):
    from segmentation3d.utils.image3d import Image3d
    return Image3d(arr)
