"""Stand-ins that let the UNMODIFIED reference (/root/reference) import and run in this image.

Test/golden infrastructure only - never imported by the product.  The image lacks SimpleITK,
easydict and tensorboardX, numpy 2.x dropped np.int/np.float/np.uint (reference
utils/image_tools.py:11,17,20) and torch 2.x dropped DataLoaderIter.next() (core/seg_train.py:113).
`install()` must run before any `import segmentation3d...` of the reference.

The SimpleITK subset is numpy-backed and covers exactly what core/seg_infer.segmentation_volume and
utils/image_tools touch on the identity-resample configs (image spacing == model spacing and
size % max_stride == 0).  Like SimpleITK, GetArrayFromImage / GetImageFromArray COPY.
Assumption recorded for the goldens: `1.0 / image` on a float32 image is a float32 division
(core/seg_infer.py:325 casts the result to float32 anyway).
"""
import sys
import types

import numpy as np

sitkUInt8, sitkInt8, sitkUInt16, sitkInt16, sitkUInt32, sitkInt32, sitkUInt64, sitkInt64, sitkFloat32, sitkFloat64 = range(10)
_NP = {sitkUInt8: np.uint8, sitkInt8: np.int8, sitkUInt16: np.uint16, sitkInt16: np.int16, sitkUInt32: np.uint32,
       sitkInt32: np.int32, sitkUInt64: np.uint64, sitkInt64: np.int64, sitkFloat32: np.float32, sitkFloat64: np.float64}
_ID = {np.dtype(v): k for k, v in _NP.items()}
sitkLinear, sitkNearestNeighbor, sitkIdentity = 100, 101, 200


class Image(object):
    def __init__(self, size=None, pixel_id=sitkFloat32, _arr=None):
        if _arr is not None:
            self._a = _arr
        else:
            sx, sy, sz = [int(v) for v in size]
            self._a = np.zeros((sz, sy, sx), dtype=_NP[pixel_id])
        self._spacing, self._origin = (1.0, 1.0, 1.0), (0.0, 0.0, 0.0)
        self._direction = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)

    # frame
    def GetSize(self): return (int(self._a.shape[2]), int(self._a.shape[1]), int(self._a.shape[0]))
    def GetSpacing(self): return tuple(self._spacing)
    def GetOrigin(self): return tuple(self._origin)
    def GetDirection(self): return tuple(self._direction)
    def SetSpacing(self, s): self._spacing = tuple(float(v) for v in s)
    def SetOrigin(self, o): self._origin = tuple(float(v) for v in o)
    def SetDirection(self, d): self._direction = tuple(float(v) for v in d)
    def GetPixelID(self): return _ID[self._a.dtype]

    def CopyInformation(self, other):
        assert self.GetSize() == other.GetSize()
        self._spacing, self._origin, self._direction = other._spacing, other._origin, other._direction

    def TransformContinuousIndexToPhysicalPoint(self, idx):
        return tuple(self._origin[i] + self._spacing[i] * float(idx[i]) for i in range(3))

    def TransformPhysicalPointToIndex(self, pt):
        return tuple(int(np.floor((pt[i] - self._origin[i]) / self._spacing[i] + 0.5)) for i in range(3))

    def __getitem__(self, key):
        kx, ky, kz = key
        sub = Image(_arr=self._a[kz, ky, kx].copy())
        sub._spacing, sub._direction = self._spacing, self._direction
        sub._origin = tuple(self._origin[i] + self._spacing[i] * (k.start or 0) for i, k in enumerate((kx, ky, kz)))
        return sub

    def _wrap(self, arr):
        out = Image(_arr=arr)
        out._spacing, out._origin, out._direction = self._spacing, self._origin, self._direction
        return out

    def __mul__(self, other):
        o = other._a if isinstance(other, Image) else other
        return self._wrap((self._a * o).astype(self._a.dtype))
    __rmul__ = __mul__

    def __rtruediv__(self, other):
        with np.errstate(divide='ignore'):
            return self._wrap((self._a.dtype.type(other) / self._a).astype(self._a.dtype))


def GetArrayFromImage(image): return image._a.copy()
def GetImageFromArray(arr): return Image(_arr=np.array(arr, copy=True))
def Cast(image, pixel_id): return image._wrap(image._a.astype(_NP[pixel_id]))
def Add(a, b): return a._wrap(a._a + (b._a if isinstance(b, Image) else b))


class Transform(object):
    def __init__(self, dim, kind):
        assert dim == 3 and kind == sitkIdentity


def Resample(image, *args):
    """Two call shapes (utils/image_tools.py:343 and :376); identity cases only."""
    if isinstance(args[0], Image):          # (image, reference, transform, interp, padding_value)
        ref = args[0]
        assert ref.GetSize() == image.GetSize() and np.allclose(ref.GetSpacing(), image.GetSpacing()), \
            'stub Resample only supports the identity case'
        out = Image(_arr=image._a.copy())
        out.CopyInformation(ref)
        return out
    out_size, _, _, origin, spacing, direction = args   # (image, size, transform, interp, origin, spacing, direction)
    assert tuple(int(v) for v in out_size) == image.GetSize() and np.allclose(spacing, image.GetSpacing()), \
        'stub Resample only supports the identity case (size %% max_stride == 0, spacing == model spacing)'
    out = Image(_arr=image._a.copy())
    out.SetOrigin(origin), out.SetSpacing(spacing), out.SetDirection(direction)
    return out


def _unsupported(*a, **k):
    raise NotImplementedError('not covered by the SimpleITK stand-in')


class _AttrDict(dict):
    """easydict.EasyDict stand-in: attribute access, nested dicts converted on set."""
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _AttrDict):
            v = _AttrDict(v)
        super().__setitem__(k, v)

    def __setattr__(self, k, v): self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def install():
    """Put the stand-ins into sys.modules and restore the numpy aliases the reference uses."""
    if 'SimpleITK' not in sys.modules:
        m = types.ModuleType('SimpleITK')
        for k, v in globals().items():
            if k.startswith('sitk') or k in ('Image', 'GetArrayFromImage', 'GetImageFromArray', 'Cast', 'Add',
                                             'Transform', 'Resample'):
                setattr(m, k, v)
        for k in ('ReadImage', 'WriteImage', 'Paste', 'ConnectedComponentImageFilter', 'RelabelComponent',
                  'LabelShapeStatisticsImageFilter', 'ImageSeriesReader', 'ImageFileWriter'):
            setattr(m, k, _unsupported)
        sys.modules['SimpleITK'] = m
    if 'easydict' not in sys.modules:
        e = types.ModuleType('easydict')
        e.EasyDict = _AttrDict
        sys.modules['easydict'] = e
    for name, val in (('int', int), ('float', float), ('uint', np.uint64)):
        if name not in np.__dict__:
            setattr(np, name, val)
    return sys.modules['SimpleITK']
