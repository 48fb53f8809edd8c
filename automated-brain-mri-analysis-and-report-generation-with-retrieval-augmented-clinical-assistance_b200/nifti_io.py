"""Minimal NIfTI-1 single-file reader / writer (.nii, .nii.gz) — numpy + gzip only.

The reference reads modalities through SimpleITK inside nnU-Net v1 (`trainer.preprocess_patient`,
run_brats2021_inference_singlethread.py:89) and segmentations through nibabel (`:217-243`, `:299-308`); neither package
exists in this image.  Arrays are returned in SimpleITK order (z, y, x): the file stores x fastest, so the raw buffer
reshaped C-contiguously to (dim3, dim2, dim1) IS that array — no transposition, bit-identical voxels.
"""
import gzip
import struct

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


class NiftiImage:
    """data: ndarray (z, y, x) in file dtype (scaling NOT applied); header: the 348 raw bytes; endian '<' or '>'."""

    def __init__(self, data, header, endian="<"):
        self.data, self.header, self.endian = data, bytes(header), endian

    def _f(self, fmt, off):
        return struct.unpack_from(self.endian + fmt, self.header, off)

    @property
    def zooms(self):
        """Voxel sizes (x, y, z) in mm — nibabel's `img.header.get_zooms()`."""
        return tuple(float(v) for v in self._f("3f", 80))

    @property
    def slope_inter(self):
        slope, inter = self._f("2f", 112)
        if slope == 0 or not np.isfinite(slope):
            return 1.0, 0.0
        return float(slope), float(inter if np.isfinite(inter) else 0.0)

    def get_fdata(self):
        """float64 array (z, y, x) with scl_slope / scl_inter applied, as nibabel's get_fdata() up to axis order."""
        slope, inter = self.slope_inter
        out = self.data.astype(np.float64)
        if slope != 1.0 or inter != 0.0:
            out = out * slope + inter
        return out


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def load(path):
    with _open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 352:
        raise ValueError(f"{path}: not a NIfTI-1 file (too short)")
    endian = "<"
    if struct.unpack_from("<i", raw, 0)[0] != 348:
        endian = ">"
        if struct.unpack_from(">i", raw, 0)[0] != 348:
            raise ValueError(f"{path}: sizeof_hdr is not 348")
    if raw[344:347] != b"n+1":
        raise ValueError(f"{path}: only single-file NIfTI-1 (magic 'n+1') is supported")
    dim = struct.unpack_from(endian + "8h", raw, 40)
    if dim[0] < 3 or any(d != 1 for d in dim[4:1 + dim[0]]):
        raise ValueError(f"{path}: expected a 3-D volume, dim = {dim}")
    code, = struct.unpack_from(endian + "h", raw, 70)
    if code not in _DTYPES:
        raise ValueError(f"{path}: unsupported datatype code {code}")
    vox_offset = int(struct.unpack_from(endian + "f", raw, 108)[0])
    nx, ny, nz = dim[1], dim[2], dim[3]
    dt = np.dtype(_DTYPES[code]).newbyteorder(endian)
    data = np.frombuffer(raw, dtype=dt, count=nx * ny * nz, offset=max(vox_offset, 352)).reshape(nz, ny, nx)
    return NiftiImage(np.array(data, dtype=dt.newbyteorder("=")), raw[:348], endian)  # own, writable, native-endian


def save(path, data, like):
    """Write `data` (z, y, x) with the geometry (dims checked, pixdim, qform / sform) of `like`'s header; dtype, bitpix,
    vox_offset and the scaling fields are rewritten for `data` — what SimpleITK / nibabel do when saving a label map
    with the source image's header."""
    data = np.ascontiguousarray(data)
    if np.dtype(data.dtype) not in _CODES:
        raise ValueError(f"unsupported dtype {data.dtype}")
    hdr = bytearray(like.header)
    e = like.endian
    dim = list(struct.unpack_from(e + "8h", hdr, 40))
    if (dim[3], dim[2], dim[1]) != tuple(data.shape):
        raise ValueError(f"shape {data.shape} does not match the header geometry {(dim[3], dim[2], dim[1])}")
    struct.pack_into(e + "8h", hdr, 40, 3, dim[1], dim[2], dim[3], 1, 1, 1, 1)
    struct.pack_into(e + "h", hdr, 70, _CODES[np.dtype(data.dtype)])
    struct.pack_into(e + "h", hdr, 72, data.dtype.itemsize * 8)
    struct.pack_into(e + "f", hdr, 108, 352.0)
    struct.pack_into(e + "2f", hdr, 112, 1.0, 0.0)
    struct.pack_into(e + "2f", hdr, 124, 0.0, 0.0)  # cal_max, cal_min
    hdr[344:348] = b"n+1\0"
    with _open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(b"\0\0\0\0")
        f.write(data.astype(data.dtype.newbyteorder(e), copy=False).tobytes())


def new_header(shape_zyx, zooms_xyz=(1.0, 1.0, 1.0), dtype=np.float32):
    """A plain little-endian NIfTI-1 header (identity orientation) — for tests and synthetic cases."""
    hdr = bytearray(348)
    nz, ny, nx = shape_zyx
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, nx, ny, nz, 1, 1, 1, 1)
    struct.pack_into("<h", hdr, 70, _CODES[np.dtype(dtype)])
    struct.pack_into("<h", hdr, 72, np.dtype(dtype).itemsize * 8)
    struct.pack_into("<8f", hdr, 76, 1.0, zooms_xyz[0], zooms_xyz[1], zooms_xyz[2], 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<2f", hdr, 112, 1.0, 0.0)
    struct.pack_into("<h", hdr, 254, 2)  # sform_code = aligned
    struct.pack_into("<4f", hdr, 280, zooms_xyz[0], 0.0, 0.0, 0.0)
    struct.pack_into("<4f", hdr, 296, 0.0, zooms_xyz[1], 0.0, 0.0)
    struct.pack_into("<4f", hdr, 312, 0.0, 0.0, zooms_xyz[2], 0.0)
    hdr[344:348] = b"n+1\0"
    return NiftiImage(np.zeros(shape_zyx, dtype=dtype), bytes(hdr), "<")
