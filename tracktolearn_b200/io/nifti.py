"""Minimal NIfTI-1 reader / writer (.nii, .nii.gz) -- nibabel is not available in the target
image.  Covers what the tracking CLI needs: dims, datatype, scaling, zooms and the vox->RAS mm
affine (sform, else qform, else zooms)."""
import gzip
import struct

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8,
           512: np.uint16, 768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).name: k for k, v in _DTYPES.items()}


class NiftiImage(object):
    def __init__(self, data, affine, zooms=None):
        self.data = data
        self.affine = np.asarray(affine, dtype=np.float64)
        self.zooms = tuple(zooms) if zooms is not None else tuple(
            float(np.linalg.norm(self.affine[:3, i])) for i in range(3))

    @property
    def shape(self):
        return self.data.shape

    def get_fdata(self, dtype=np.float64):
        return np.asarray(self.data, dtype=dtype)


def _quat_to_affine(b, c, d, qx, qy, qz, dx, dy, dz, qfac):
    a = np.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
    R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
    A = np.eye(4)
    A[:3, :3] = R * np.array([dx, dy, dz * qfac])
    A[:3, 3] = [qx, qy, qz]
    return A


def load(path):
    opener = gzip.open if str(path).endswith('.gz') else open
    with opener(path, 'rb') as f:
        raw = f.read()
    if len(raw) < 348:
        raise ValueError('%s: not a NIfTI-1 file' % path)
    endian = '<'
    if struct.unpack('<i', raw[:4])[0] != 348:
        endian = '>'
        if struct.unpack('>i', raw[:4])[0] != 348:
            raise ValueError('%s: bad NIfTI-1 header size' % path)
    dim = struct.unpack(endian + '8h', raw[40:56])
    datatype, bitpix = struct.unpack(endian + '2h', raw[70:74])
    pixdim = struct.unpack(endian + '8f', raw[76:108])
    vox_offset, slope, inter = struct.unpack(endian + '3f', raw[108:120])
    qform_code, sform_code = struct.unpack(endian + '2h', raw[252:256])
    quat = struct.unpack(endian + '6f', raw[256:280])
    srow = np.array(struct.unpack(endian + '12f', raw[280:328]), dtype=np.float64).reshape(3, 4)
    if raw[344:348] not in (b'n+1\x00', b'ni1\x00'):
        raise ValueError('%s: unsupported NIfTI magic %r' % (path, raw[344:348]))
    if datatype not in _DTYPES:
        raise ValueError('%s: unsupported NIfTI datatype %d' % (path, datatype))
    ndim = dim[0]
    shape = tuple(int(d) for d in dim[1:1 + ndim])
    while len(shape) > 3 and shape[-1] == 1:
        shape = shape[:-1]
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(endian)
    n = int(np.prod(shape))
    off = int(vox_offset) if vox_offset >= 348 else 352
    data = np.frombuffer(raw, dtype=dt, count=n, offset=off).reshape(shape, order='F')
    if slope not in (0.0, 1.0) or inter != 0.0:
        if slope != 0.0 and not np.isnan(slope):
            data = data.astype(np.float64) * slope + inter
    if sform_code > 0:
        affine = np.vstack((srow, [0, 0, 0, 1]))
    elif qform_code > 0:
        qfac = -1.0 if pixdim[0] < 0 else 1.0
        affine = _quat_to_affine(quat[0], quat[1], quat[2], quat[3], quat[4], quat[5],
                                 pixdim[1], pixdim[2], pixdim[3], qfac)
    else:
        affine = np.diag([pixdim[1], pixdim[2], pixdim[3], 1.0])
    return NiftiImage(np.ascontiguousarray(data), affine, pixdim[1:4])


def save(path, data, affine):
    data = np.asarray(data)
    if data.dtype.name not in _CODES:
        data = data.astype(np.float32)
    affine = np.asarray(affine, dtype=np.float64)
    hdr = bytearray(348)
    struct.pack_into('<i', hdr, 0, 348)
    dim = [data.ndim] + list(data.shape) + [1] * (7 - data.ndim)
    struct.pack_into('<8h', hdr, 40, *dim)
    struct.pack_into('<2h', hdr, 70, _CODES[data.dtype.name], data.dtype.itemsize * 8)
    zooms = [float(np.linalg.norm(affine[:3, i])) for i in range(3)]
    struct.pack_into('<8f', hdr, 76, 1.0, zooms[0], zooms[1], zooms[2], 1.0, 1.0, 1.0, 1.0)
    struct.pack_into('<3f', hdr, 108, 352.0, 1.0, 0.0)
    struct.pack_into('<2h', hdr, 252, 0, 1)
    struct.pack_into('<12f', hdr, 280, *affine[:3].reshape(-1))
    hdr[344:348] = b'n+1\x00'
    payload = bytes(hdr) + b'\x00' * 4 + np.asfortranarray(data).tobytes(order='F')
    opener = gzip.open if str(path).endswith('.gz') else open
    with opener(path, 'wb') as f:
        f.write(payload)
