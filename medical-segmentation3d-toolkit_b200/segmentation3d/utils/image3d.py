"""Minimal 3-D image container + MetaImage (.mha/.mhd) and NIfTI-1 (.nii/.nii.gz) I/O.

The reference uses SimpleITK images everywhere above the hot path (core/seg_infer.py:414,467-481);
SimpleITK is not part of this image, so the engine carries its own container with the same accessor
names (GetSize/GetSpacing/GetOrigin/GetDirection/CopyInformation, sizes in x,y,z order) backed by
a numpy array in [z,y,x] order - or by a CUDA tensor when the data should stay on the device.
`as_image3d` also accepts real SimpleITK images when that package is installed.
"""
import gzip
import os
import struct
import zlib

import numpy as np
import torch

_MET = {'MET_CHAR': np.int8, 'MET_UCHAR': np.uint8, 'MET_SHORT': np.int16, 'MET_USHORT': np.uint16,
        'MET_INT': np.int32, 'MET_UINT': np.uint32, 'MET_LONG': np.int64, 'MET_ULONG': np.uint64,
        'MET_FLOAT': np.float32, 'MET_DOUBLE': np.float64}
_MET_INV = {np.dtype(v): k for k, v in _MET.items()}


def zlib_compress(payload, level=1, chunk=4 << 20, threads=None):
    """One valid zlib stream for `payload`.  Large buffers are deflated in independent chunks on worker threads (zlib
    releases the GIL): every chunk is a raw deflate stream ended by a sync flush, so the pieces concatenate on byte
    boundaries; the 2-byte zlib header goes in front and the Adler-32 of the whole buffer behind (the pigz construction).
    Any inflater - ITK's MetaImage reader included - reads it as an ordinary stream."""
    threads = io_threads() if threads is None else threads
    if threads < 2 or len(payload) < 2 * chunk:
        return zlib.compress(payload, level)
    from concurrent.futures import ThreadPoolExecutor
    view = memoryview(payload)
    pieces = [view[i:i + chunk] for i in range(0, len(view), chunk)]

    def deflate(args):
        i, piece = args
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        return c.compress(piece) + c.flush(zlib.Z_FINISH if i == len(pieces) - 1 else zlib.Z_SYNC_FLUSH)
    with ThreadPoolExecutor(max_workers=threads) as pool:
        parts = list(pool.map(deflate, enumerate(pieces)))
    header = b'\x78\x01'                                    # CMF = deflate / 32 KB window, FLG = fastest, check bits
    return header + b''.join(parts) + struct.pack('>I', zlib.adler32(payload) & 0xffffffff)


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
    if low.endswith('.nii') or low.endswith('.nii.gz'):
        return read_nifti(path, dtype)
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
    if path.lower().endswith('.nii') or path.lower().endswith('.nii.gz'):
        return write_nifti(image, path)
    if not path.lower().endswith('.mha'):
        raise ValueError('only .mha / .nii / .nii.gz can be written without SimpleITK: %s' % path)
    payload = arr.tobytes()
    if compress:
        payload = zlib_compress(payload, 1)
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


# ---- host-side I/O overlap for batch inference: the next case is read (and decompressed) while the current one is on the GPU,
# and results are compressed and written in the background (zlib releases the GIL) --------------------------------------------
def io_threads():
    """worker threads for background image I/O (default: a quarter of the host cores, between 4 and 8);
    SEG3D_IO_THREADS=0 restores strictly serial reads and writes"""
    default = min(8, max(4, (os.cpu_count() or 4) // 4))
    try:
        return max(0, int(os.environ.get('SEG3D_IO_THREADS', default)))
    except ValueError:
        return default


def prefetch_images(paths, dtype=None, enabled=True, depth=None):
    """yield (image, seconds spent waiting for it) for every path in order, with up to `depth` (default: io_threads())
    later paths being read on worker threads.  A read error surfaces when its own case is reached, as in the serial loop."""
    import collections
    import time
    from concurrent.futures import ThreadPoolExecutor
    paths = list(paths)
    depth = io_threads() if depth is None else depth
    if not enabled or depth < 1 or len(paths) < 2:
        for p in paths:
            t0 = time.time()
            img = read_image(p, dtype)
            yield img, time.time() - t0
        return
    pool = ThreadPoolExecutor(max_workers=depth)
    try:
        queue, nxt = collections.deque(), 0
        while nxt < len(paths) and len(queue) < depth:
            queue.append(pool.submit(read_image, paths[nxt], dtype))
            nxt += 1
        while queue:
            t0 = time.time()
            img = queue.popleft().result()
            waited = time.time() - t0
            if nxt < len(paths):
                queue.append(pool.submit(read_image, paths[nxt], dtype))
                nxt += 1
            yield img, waited
    finally:
        pool.shutdown(wait=False)


class AsyncImageWriter(object):
    """write_image in the background.  The device-to-host copy happens in `write` on the caller's thread (so it is ordered
    with the caller's CUDA stream); only compression and the file write run on the pool.  `close` waits for every file
    and re-raises the first error."""

    def __init__(self, threads):
        from concurrent.futures import ThreadPoolExecutor
        self.pool = ThreadPoolExecutor(max_workers=threads) if threads > 0 else None
        self.pending = []

    def write(self, image, path, compress=False):
        image = as_image3d(image)
        host = Image3d(np.ascontiguousarray(image.to_numpy()), image.GetSpacing(), image.GetOrigin(), image.GetDirection())
        if self.pool is None:
            write_image(host, path, compress)
        else:
            self.pending.append(self.pool.submit(write_image, host, path, compress))

    def close(self):
        err = None
        for f in self.pending:
            try:
                f.result()
            except Exception as e:          # keep draining so no file is left half-written behind a raised error
                err = err or e
        self.pending = []
        if self.pool is not None:
            self.pool.shutdown(wait=True)
            self.pool = None
        if err is not None:
            raise err


# ---- NIfTI-1 (single file, "n+1"): the reference reads and writes it through ITK's NiftiImageIO --------------------------------
# NIfTI stores its affine in RAS+ coordinates, ITK images live in LPS+: on the way in the x and y rows of the affine change
# sign (origin and direction), on the way out they change back.  The qform (quaternion) is preferred when its code is set,
# the sform (three affine rows) otherwise; with neither the axes are the identity scaled by pixdim.
_NII_DT = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32, 1024: np.int64, 1280: np.uint64}
_NII_DT_INV = {np.dtype(v): k for k, v in _NII_DT.items()}


def _quatern_to_mat(b, c, d, qfac):
    """rotation matrix of NIfTI's unit quaternion (a, b, c, d), a = sqrt(1 - b^2 - c^2 - d^2), last column times qfac"""
    a = 1.0 - (b * b + c * c + d * d)
    if a < 1e-7:                          # special case of nifti_quatern_to_mat44: 180 degree rotation
        a = 1.0 / np.sqrt(b * b + c * c + d * d)
        b, c, d, a = b * a, c * a, d * a, 0.0
    else:
        a = np.sqrt(a)
    r = np.array([[a * a + b * b - c * c - d * d, 2 * b * c - 2 * a * d, 2 * b * d + 2 * a * c],
                  [2 * b * c + 2 * a * d, a * a + c * c - b * b - d * d, 2 * c * d - 2 * a * b],
                  [2 * b * d - 2 * a * c, 2 * c * d + 2 * a * b, a * a + d * d - c * c - b * b]], dtype=np.float64)
    r[:, 2] *= qfac
    return r


def _mat_to_quatern(r):
    """inverse of _quatern_to_mat for an orthonormal matrix: (b, c, d, qfac) (nifti_mat44_to_quatern)"""
    r = np.array(r, dtype=np.float64)
    qfac = 1.0
    if np.linalg.det(r) < 0:
        r[:, 2] = -r[:, 2]
        qfac = -1.0
    a = r[0, 0] + r[1, 1] + r[2, 2] + 1.0
    if a > 0.5:
        a = 0.5 * np.sqrt(a)
        b, c, d = 0.25 * (r[2, 1] - r[1, 2]) / a, 0.25 * (r[0, 2] - r[2, 0]) / a, 0.25 * (r[1, 0] - r[0, 1]) / a
    else:
        xd, yd, zd = 1.0 + r[0, 0] - (r[1, 1] + r[2, 2]), 1.0 + r[1, 1] - (r[0, 0] + r[2, 2]), 1.0 + r[2, 2] - (r[0, 0] + r[1, 1])
        if xd > 1.0:
            b = 0.5 * np.sqrt(xd)
            c, d, a = 0.25 * (r[0, 1] + r[1, 0]) / b, 0.25 * (r[0, 2] + r[2, 0]) / b, 0.25 * (r[2, 1] - r[1, 2]) / b
        elif yd > 1.0:
            c = 0.5 * np.sqrt(yd)
            b, d, a = 0.25 * (r[0, 1] + r[1, 0]) / c, 0.25 * (r[1, 2] + r[2, 1]) / c, 0.25 * (r[0, 2] - r[2, 0]) / c
        else:
            d = 0.5 * np.sqrt(zd)
            b, c, a = 0.25 * (r[0, 2] + r[2, 0]) / d, 0.25 * (r[1, 2] + r[2, 1]) / d, 0.25 * (r[1, 0] - r[0, 1]) / d
        if a < 0.0:
            b, c, d = -b, -c, -d
    return float(b), float(c), float(d), qfac


def read_nifti(path, dtype=None):
    """NIfTI-1 single file (.nii / .nii.gz) -> Image3d in ITK's LPS convention."""
    opener = gzip.open if path.lower().endswith('.gz') else open
    with opener(path, 'rb') as f:
        blob = f.read()
    if len(blob) < 348:
        raise ValueError('not a NIfTI-1 file: %s' % path)
    end = '<' if struct.unpack('<i', blob[0:4])[0] == 348 else '>'
    if struct.unpack(end + 'i', blob[0:4])[0] != 348 or blob[344:347] not in (b'n+1', b'ni1'):
        raise ValueError('not a NIfTI-1 file: %s' % path)
    if blob[344:347] == b'ni1':
        raise ValueError('two-file NIfTI (.hdr/.img) is not supported: %s' % path)
    dim = struct.unpack(end + '8h', blob[40:56])
    if dim[0] < 3 or any(d != 1 for d in dim[4:1 + dim[0]]):
        raise ValueError('only 3-D NIfTI volumes are supported (dim = %s)' % (dim,))
    nx, ny, nz = dim[1:4]
    datatype = struct.unpack(end + 'h', blob[70:72])[0]
    if datatype not in _NII_DT:
        raise ValueError('unsupported NIfTI datatype %d' % datatype)
    pixdim = struct.unpack(end + '8f', blob[76:108])
    vox_offset = int(struct.unpack(end + 'f', blob[108:112])[0])
    slope, inter = struct.unpack(end + '2f', blob[112:120])
    qform_code, sform_code = struct.unpack(end + '2h', blob[252:256])
    np_t = np.dtype(_NII_DT[datatype]).newbyteorder(end)
    arr = np.frombuffer(blob, dtype=np_t, count=nx * ny * nz, offset=max(vox_offset, 352)).reshape(nz, ny, nx)
    arr = arr.astype(np_t.newbyteorder('='), copy=True)
    if slope != 0.0 and np.isfinite(slope) and (slope != 1.0 or inter != 0.0):
        arr = arr.astype(np.float64) * slope + inter
    spacing = [abs(float(v)) if v != 0 else 1.0 for v in pixdim[1:4]]
    if qform_code > 0:
        b, c, d, qx, qy, qz = struct.unpack(end + '6f', blob[256:280])
        rot = _quatern_to_mat(b, c, d, -1.0 if pixdim[0] < 0 else 1.0)
        origin = np.array([qx, qy, qz], dtype=np.float64)
    elif sform_code > 0:
        rows = np.array(struct.unpack(end + '12f', blob[280:328]), dtype=np.float64).reshape(3, 4)
        spacing = [float(np.linalg.norm(rows[:, a])) or 1.0 for a in range(3)]
        rot = rows[:, :3] / np.array(spacing)
        origin = rows[:, 3].copy()
    else:
        rot, origin = np.eye(3), np.zeros(3)
    flip = np.diag([-1.0, -1.0, 1.0])                # RAS+ (NIfTI) -> LPS+ (ITK)
    direction = flip.dot(rot)
    origin = flip.dot(origin)
    if dtype is not None:
        arr = arr.astype(dtype)
    return Image3d(arr, spacing, origin.tolist(), direction.reshape(-1).tolist())


def write_nifti(image, path):
    """Image3d -> NIfTI-1 single file with qform and sform set (codes 1), gzip-compressed when the name ends in .gz."""
    image = as_image3d(image)
    arr = np.ascontiguousarray(image.to_numpy())
    if arr.dtype == np.bool_:
        arr = arr.astype(np.uint8)
    if arr.dtype not in _NII_DT_INV:
        raise ValueError('unsupported dtype for NIfTI: %s' % arr.dtype)
    nz, ny, nx = arr.shape
    flip = np.diag([-1.0, -1.0, 1.0])                # LPS+ -> RAS+
    rot = flip.dot(np.asarray(image.direction, dtype=np.float64).reshape(3, 3))
    origin = flip.dot(np.asarray(image.origin, dtype=np.float64))
    sp = np.asarray(image.spacing, dtype=np.float64)
    b, c, d, qfac = _mat_to_quatern(rot)
    hdr = bytearray(348)
    struct.pack_into('<i', hdr, 0, 348)
    struct.pack_into('<8h', hdr, 40, 3, nx, ny, nz, 1, 1, 1, 1)
    struct.pack_into('<h', hdr, 70, _NII_DT_INV[arr.dtype])
    struct.pack_into('<h', hdr, 72, arr.dtype.itemsize * 8)
    struct.pack_into('<8f', hdr, 76, qfac, sp[0], sp[1], sp[2], 0.0, 0.0, 0.0, 0.0)
    struct.pack_into('<f', hdr, 108, 352.0)
    struct.pack_into('<2f', hdr, 112, 1.0, 0.0)
    hdr[123] = 2                                      # xyzt_units: millimetres
    struct.pack_into('<2h', hdr, 252, 1, 1)
    struct.pack_into('<6f', hdr, 256, b, c, d, origin[0], origin[1], origin[2])
    aff = rot * sp
    for r in range(3):
        struct.pack_into('<4f', hdr, 280 + 16 * r, aff[r, 0], aff[r, 1], aff[r, 2], origin[r])
    hdr[344:348] = b'n+1\0'
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    payload = bytes(hdr) + b'\0\0\0\0' + arr.astype(arr.dtype.newbyteorder('<'), copy=False).tobytes()
    if path.lower().endswith('.gz'):
        with gzip.open(path, 'wb', compresslevel=1) as f:
            f.write(payload)
    else:
        with open(path, 'wb') as f:
            f.write(payload)
