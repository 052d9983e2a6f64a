"""Minimal 3-D image container + MetaImage (.mha/.mhd) I/O.

The reference uses SimpleITK images everywhere above the hot path (core/seg_infer.py:414,467-481);
SimpleITK is not part of this image, so the engine carries its own container with the same accessor
names (GetSize/GetSpacing/GetOrigin/GetDirection/CopyInformation, sizes in x,y,z order) backed by
a numpy array in [z,y,x] order - or by a CUDA tensor when the data should stay on the device.
`as_image3d` also accepts real SimpleITK images when that package is installed.
"""
import os
import zlib

import numpy as np
import torch

_MET = {'MET_CHAR': np.int8, 'MET_UCHAR': np.uint8, 'MET_SHORT': np.int16, 'MET_USHORT': np.uint16,
        'MET_INT': np.int32, 'MET_UINT': np.uint32, 'MET_LONG': np.int64, 'MET_ULONG': np.uint64,
        'MET_FLOAT': np.float32, 'MET_DOUBLE': np.float64}
_MET_INV = {np.dtype(v): k for k, v in _MET.items()}


class Image3d(object):
    def __init__(self, data, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0),
                 direction=(1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)):
        assert data.ndim == 3, 'Image3d holds a [z,y,x] array'
        self.data = data
        self.spacing = tuple(float(v) for v in spacing)
        self.origin = tuple(float(v) for v in origin)
        self.direction = tuple(float(v) for v in direction)

    def GetSize(self):
        return (int(self.data.shape[2]), int(self.data.shape[1]), int(self.data.shape[0]))

    def GetSpacing(self):
        return self.spacing

    def GetOrigin(self):
        return self.origin

    def GetDirection(self):
        return self.direction

    def SetSpacing(self, s):
        self.spacing = tuple(float(v) for v in s)

    def SetOrigin(self, o):
        self.origin = tuple(float(v) for v in o)

    def SetDirection(self, d):
        self.direction = tuple(float(v) for v in d)

    def CopyInformation(self, other):
        assert self.GetSize() == other.GetSize(), 'image sizes differ'
        self.spacing, self.origin, self.direction = other.GetSpacing(), other.GetOrigin(), other.GetDirection()

    def is_cuda(self):
        return torch.is_tensor(self.data) and self.data.is_cuda

    def to_numpy(self):
        if torch.is_tensor(self.data):
            return self.data.detach().cpu().numpy()
        return self.data

    def TransformContinuousIndexToPhysicalPoint(self, idx):
        d = np.asarray(self.direction).reshape(3, 3)
        return tuple((np.asarray(self.origin) + d.dot(np.asarray(idx, dtype=np.float64) * np.asarray(self.spacing))).tolist())

    def TransformPhysicalPointToIndex(self, pt):
        d = np.asarray(self.direction).reshape(3, 3)
        c = np.linalg.solve(d, np.asarray(pt, dtype=np.float64) - np.asarray(self.origin)) / np.asarray(self.spacing)
        return tuple(int(np.floor(v + 0.5)) for v in c)


def as_image3d(obj):
    """Image3d from an Image3d, a SimpleITK image, a numpy array or a tensor ([z,y,x])."""
    if isinstance(obj, Image3d):
        return obj
    if isinstance(obj, np.ndarray) or torch.is_tensor(obj):
        return Image3d(obj)
    try:
        import SimpleITK as sitk
        if isinstance(obj, sitk.Image):
            return Image3d(sitk.GetArrayFromImage(obj), obj.GetSpacing(), obj.GetOrigin(), obj.GetDirection())
    except ImportError:
        pass
    raise TypeError('unsupported image type %r' % type(obj))


def read_image(path, dtype=None):
    """Read .mha / .mhd (MetaImage, raw or zlib-compressed).  dtype: optional numpy dtype to cast to
    (the reference reads test images as float32, core/seg_infer.py:414)."""
    low = path.lower()
    if not (low.endswith('.mha') or low.endswith('.mhd')):
        try:
            import SimpleITK as sitk
            img = sitk.ReadImage(path)
            out = as_image3d(img)
            if dtype is not None:
                out.data = out.data.astype(dtype)
            return out
        except ImportError:
            raise ValueError('only .mha/.mhd can be read without SimpleITK: %s' % path)
    with open(path, 'rb') as f:
        blob = f.read()
    hdr, pos = {}, 0
    while True:
        end = blob.index(b'\n', pos)
        line = blob[pos:end].decode('ascii', 'replace').strip()
        pos = end + 1
        if '=' in line:
            k, v = [t.strip() for t in line.split('=', 1)]
            hdr[k] = v
            if k == 'ElementDataFile':
                break
    assert int(hdr.get('NDims', 3)) == 3, 'only 3-D MetaImages are supported'
    size = [int(v) for v in hdr['DimSize'].split()]
    np_t = np.dtype(_MET[hdr['ElementType']])
    if hdr.get('BinaryDataByteOrderMSB', hdr.get('ElementByteOrderMSB', 'False')).lower() == 'true':
        np_t = np_t.newbyteorder('>')
    if hdr['ElementDataFile'] == 'LOCAL':
        payload = blob[pos:]
    else:
        with open(os.path.join(os.path.dirname(path), hdr['ElementDataFile']), 'rb') as f:
            payload = f.read()
    if hdr.get('CompressedData', 'False').lower() == 'true':
        payload = zlib.decompress(payload)
    arr = np.frombuffer(payload, dtype=np_t, count=size[0] * size[1] * size[2]).reshape(size[2], size[1], size[0])
    arr = arr.astype(dtype if dtype is not None else np_t.newbyteorder('='), copy=True)
    spacing = [float(v) for v in hdr.get('ElementSpacing', hdr.get('ElementSize', '1 1 1')).split()]
    origin = [float(v) for v in hdr.get('Offset', hdr.get('Position', hdr.get('Origin', '0 0 0'))).split()]
    direction = [float(v) for v in hdr.get('TransformMatrix', hdr.get('Rotation', '1 0 0 0 1 0 0 0 1')).split()]
    # MetaImage stores direction cosines column-wise per axis; ITK's GetDirection is the transpose
    direction = np.asarray(direction).reshape(3, 3).T.reshape(-1).tolist()
    return Image3d(arr, spacing, origin, direction)


def write_image(image, path, compress=False):
    """Write a MetaImage .mha (header + raw or zlib payload), as sitk.WriteImage(img, path, compress)."""
    image = as_image3d(image)
    arr = np.ascontiguousarray(image.to_numpy())
    if arr.dtype == np.bool_:
        arr = arr.astype(np.uint8)
    if not path.lower().endswith('.mha'):
        raise ValueError('only .mha can be written without SimpleITK: %s' % path)
    payload = arr.tobytes()
    if compress:
        payload = zlib.compress(payload, 1)
    x, y, z = image.GetSize()
    tm = np.asarray(image.direction).reshape(3, 3).T.reshape(-1)
    lines = ['ObjectType = Image', 'NDims = 3', 'BinaryData = True', 'BinaryDataByteOrderMSB = False',
             'CompressedData = %s' % ('True' if compress else 'False')]
    if compress:
        lines.append('CompressedDataSize = %d' % len(payload))
    lines += ['TransformMatrix = ' + ' '.join('%.17g' % v for v in tm),
              'Offset = ' + ' '.join('%.17g' % v for v in image.origin),
              'CenterOfRotation = 0 0 0', 'AnatomicalOrientation = RAI',
              'ElementSpacing = ' + ' '.join('%.17g' % v for v in image.spacing),
              'DimSize = %d %d %d' % (x, y, z), 'ElementType = %s' % _MET_INV[arr.dtype],
              'ElementDataFile = LOCAL']
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, 'wb') as f:
        f.write(('\n'.join(lines) + '\n').encode('ascii'))
        f.write(payload)
