"""Crop normalisers (drop-in for reference utils/normalizer.py:6-81).  `to_dict()` is the format
stored in params.pth['crop_normalizers'] and consumed by the device-side patch gather kernel."""
from segmentation3d.utils.image_tools import normalize_image, get_mean_std_from_image


class _Normalizer(object):
    def __call__(self, image):
        if isinstance(image, (list, tuple)):
            for i, im in enumerate(image):
                image[i] = self.normalize(im)
            return image
        return self.normalize(image)


class FixedNormalizer(_Normalizer):
    """(v - mean) / stddev, clipped to [-1, 1] when `clip`."""

    def __init__(self, mean, stddev, clip=True):
        assert stddev > 0, 'stddev must be positive'
        assert isinstance(clip, bool), 'clip must be a boolean'
        self.mean, self.stddev, self.clip = mean, stddev, clip

    def normalize(self, image):
        return normalize_image(image, self.mean, self.stddev, self.clip)

    def to_dict(self):
        return {'type': 0, 'mean': self.mean, 'stddev': self.stddev, 'clip': self.clip}


class AdaptiveNormalizer(_Normalizer):
    """z-score with the crop's own mean/std, clipped to +-clip_sigma."""

    def __init__(self, clip_sigma=3):
        assert clip_sigma > 0
        self.clip_sigma = clip_sigma

    def normalize(self, image):
        mean, std = get_mean_std_from_image(image)
        return normalize_image(image, mean, max(std, 1e-6), True, -self.clip_sigma, self.clip_sigma)

    def to_dict(self):
        return {'type': 1, 'clip_sigma': self.clip_sigma}


def normalizer_from_dict(d):
    if d['type'] == 0:
        return FixedNormalizer(d['mean'], d['stddev'], d['clip'])
    if d['type'] == 1:
        return AdaptiveNormalizer(d['clip_sigma'])
    raise ValueError('Unsupported normalization type.')
