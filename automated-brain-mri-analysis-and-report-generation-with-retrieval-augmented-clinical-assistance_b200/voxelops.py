"""Device-side voxel operations on label volumes (thin torch wrappers over the C ABI in postproc.cu).

Everything here runs on ``cuda`` tensors through libbrainseg_b200.so; there is no CPU implementation.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

COMP_DTYPE = np.dtype([("count", "<u8"), ("s0", "<u8"), ("s1", "<u8"), ("s2", "<u8"), ("n1", "<u8"), ("n2", "<u8"),
                       ("n3", "<u8"), ("mn0", "<i4"), ("mn1", "<i4"), ("mn2", "<i4"), ("mx0", "<i4"), ("mx1", "<i4"),
                       ("mx2", "<i4"), ("pad0", "<i4"), ("pad1", "<i4")])
MOM_DTYPE = np.dtype([("count", "<u8"), ("s0", "<u8"), ("s1", "<u8"), ("s2", "<u8"), ("s00", "<u8"), ("s11", "<u8"),
                      ("s22", "<u8"), ("s01", "<u8"), ("s02", "<u8"), ("s12", "<u8"), ("surface", "<u8"),
                      ("mn0", "<i4"), ("mn1", "<i4"), ("mn2", "<i4"), ("mx0", "<i4"), ("mx1", "<i4"), ("mx2", "<i4"),
                      ("pad0", "<i4"), ("pad1", "<i4")])
assert COMP_DTYPE.itemsize == 88 and MOM_DTYPE.itemsize == 120

MASK_GT0 = 0xFFFFFFFE  # seg > 0 (labels 1..31)


def bits_of(*labels):
    b = 0
    for l in labels:
        if not 1 <= int(l) <= 31:
            raise ValueError(f"label {l} outside 1..31")
        b |= 1 << int(l)
    return b


def device():
    if not torch.cuda.is_available():
        raise L.BsgError("brainseg_b200 needs a CUDA device (sm_100); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def as_label_volume(x):
    """numpy / torch array of labels (any dtype, 3-D) -> contiguous cuda uint8 tensor.

    Float inputs get ``np.round(x).astype(np.uint8)`` on the device (convert_labels_to_brats.py:37,
    feature_extraction/utils.py:169); integer inputs are cast."""
    dev = device()
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not torch.is_tensor(x):
        raise TypeError(f"expected a numpy array or torch tensor, got {type(x)}")
    x = x.to(dev, non_blocking=True).contiguous()
    if x.dtype == torch.uint8:
        return x
    if x.dtype in (torch.float32, torch.float64):
        out = torch.empty(x.shape, dtype=torch.uint8, device=dev)
        L.check(L.lib().bsg_round_to_u8(_ptr(x), 0 if x.dtype == torch.float32 else 1, _ptr(out), x.numel(),
                                        L.stream_ptr()))
        return out
    if x.dtype == torch.bool or not x.dtype.is_floating_point:
        return x.to(torch.uint8)
    return as_label_volume(x.float())


def label_lut(vol, lut):
    """out = lut[vol]; lut: 256 uint8 values."""
    lut = np.ascontiguousarray(np.asarray(lut, dtype=np.uint8))
    assert lut.size == 256
    out = torch.empty_like(vol)
    L.check(L.lib().bsg_label_lut_u8(_ptr(vol), _ptr(out), vol.numel(), lut.tobytes(), L.stream_ptr()))
    return out


def ensemble_round(a, b, post_lut=None):
    """np.round((a+b)/2.0).astype(uint8) with an optional fused remap LUT."""
    assert a.shape == b.shape
    out = torch.empty_like(a)
    post = None if post_lut is None else np.ascontiguousarray(np.asarray(post_lut, dtype=np.uint8)).tobytes()
    L.check(L.lib().bsg_label_pair_round_u8(_ptr(a), _ptr(b), _ptr(out), a.numel(), post, L.stream_ptr()))
    return out


def ensemble_remap_hist(a, b, gt, post_lut=None):
    """ensemble_round(a, b, post_lut) plus the joint histogram of the result against `gt`, one pass.  Returns
    (labels, hist_buf): hist_buf is the DEVICE buffer (257 int64: 16x16 bins + the count of labels >= 16) —
    joint_hist_from_buffer() reads it back when the metrics are wanted."""
    assert a.shape == b.shape == gt.shape
    out = torch.empty_like(a)
    post = None if post_lut is None else np.ascontiguousarray(np.asarray(post_lut, dtype=np.uint8)).tobytes()
    buf = torch.empty(257, dtype=torch.int64, device=a.device)
    L.check(L.lib().bsg_label_pair_round_hist_u8(_ptr(a), _ptr(b), _ptr(gt), _ptr(out), a.numel(), post, _ptr(buf),
                                                 C.c_void_p(buf.data_ptr() + 2048), L.stream_ptr()))
    return out, buf


def joint_hist_from_buffer(buf):
    h = buf.cpu().numpy()
    if h[256] != 0:
        raise L.BsgError(f"{int(h[256])} voxels carry labels >= 16; the Dice histogram supports labels 0..15")
    return h[:256].reshape(16, 16).copy()


def joint_hist(pred, gt):
    """16x16 int64 joint label histogram hist[p, g]; raises on labels >= 16."""
    assert pred.shape == gt.shape
    buf = torch.empty(257, dtype=torch.int64, device=pred.device)
    L.check(L.lib().bsg_joint_hist_u8(_ptr(pred), _ptr(gt), pred.numel(), _ptr(buf), C.c_void_p(buf.data_ptr() + 2048),
                                      L.stream_ptr()))
    h = buf.cpu().numpy()
    if h[256] != 0:
        raise L.BsgError(f"{int(h[256])} voxels carry labels >= 16; the Dice histogram supports labels 0..15")
    return h[:256].reshape(16, 16).copy()


def ccl26(vol, maskbits=MASK_GT0, stats_cap=1 << 17, want_labels=True):
    """26-connected labelling in SciPy order.  Returns (labels int32 cuda tensor, ncomp, stats structured array)."""
    assert vol.dim() == 3 and vol.dtype == torch.uint8
    d0, d1, d2 = vol.shape
    lib = L.lib()
    ws_bytes = lib.bsg_ccl26_workspace_bytes(d0, d1, d2)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=vol.device)
    labels = torch.empty((d0, d1, d2), dtype=torch.int32, device=vol.device)
    ncomp = torch.zeros(1, dtype=torch.int32, device=vol.device)
    st = torch.empty(max(stats_cap, 1) * COMP_DTYPE.itemsize, dtype=torch.uint8, device=vol.device)
    L.check(lib.bsg_ccl26_stats(_ptr(vol), d0, d1, d2, maskbits & 0xFFFFFFFF, _ptr(labels), _ptr(ncomp), _ptr(st),
                                stats_cap, _ptr(ws), ws_bytes, L.stream_ptr()))
    n = int(ncomp.item())
    if n > stats_cap:  # rare: re-run with room for every component
        return ccl26(vol, maskbits, stats_cap=n, want_labels=want_labels)
    stats = np.frombuffer(st[: n * COMP_DTYPE.itemsize].cpu().numpy().tobytes(), dtype=COMP_DTYPE).copy()
    return (labels if want_labels else None), n, stats


def masked_moments(vol, maskbits_list, surface_flags=0):
    """One pass: count / moments / bbox (/ 6-connected surface count) for up to 8 label sets."""
    assert vol.dim() == 3 and vol.dtype == torch.uint8
    d0, d1, d2 = vol.shape
    n = len(maskbits_list)
    arr = (C.c_uint32 * n)(*[int(b) & 0xFFFFFFFF for b in maskbits_list])
    out = torch.empty(n * MOM_DTYPE.itemsize, dtype=torch.uint8, device=vol.device)
    L.check(L.lib().bsg_masked_moments(_ptr(vol), d0, d1, d2, arr, n, surface_flags, _ptr(out), L.stream_ptr()))
    return np.frombuffer(out.cpu().numpy().tobytes(), dtype=MOM_DTYPE).copy()


# ------------------------------------------------------------------------------------------------------------------
# Voxel operations of the remaining feature-extraction steps (csrc/morph.cu)
# ------------------------------------------------------------------------------------------------------------------

def as_mask(x):
    """bool / numeric array (numpy or torch, 3-D) -> contiguous cuda uint8 tensor, non-zero = set."""
    dev = device()
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not torch.is_tensor(x):
        raise TypeError(f"expected a numpy array or torch tensor, got {type(x)}")
    x = x.to(dev, non_blocking=True)
    if x.dtype != torch.uint8:
        x = (x != 0).to(torch.uint8)
    return x.contiguous()


def as_intensity(x):
    """MRI intensities -> contiguous cuda float32 tensor.  The reference holds them as float64 (`get_fdata()`), but NIfTI
    stores int16 / float32, so float32 is exact for file-backed data; other float64 inputs are rounded."""
    dev = device()
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not torch.is_tensor(x):
        raise TypeError(f"expected a numpy array or torch tensor, got {type(x)}")
    return x.to(dev, non_blocking=True).to(torch.float32).contiguous()


def ccl(vol, maskbits=MASK_GT0, connectivity=26, want_labels=True):
    """scipy.ndimage.label with generate_binary_structure(3, 1 / 2 / 3) <-> connectivity 6 / 18 / 26.
    Returns (labels int32 cuda tensor or None, ncomp)."""
    assert vol.dim() == 3 and vol.dtype == torch.uint8
    d0, d1, d2 = vol.shape
    lib = L.lib()
    ws_bytes = lib.bsg_ccl26_workspace_bytes(d0, d1, d2)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=vol.device)
    labels = torch.empty((d0, d1, d2), dtype=torch.int32, device=vol.device)
    ncomp = torch.zeros(1, dtype=torch.int32, device=vol.device)
    L.check(lib.bsg_ccl_stats(_ptr(vol), d0, d1, d2, maskbits & 0xFFFFFFFF, int(connectivity), _ptr(labels), _ptr(ncomp),
                              None, 0, _ptr(ws), ws_bytes, L.stream_ptr()))
    return (labels if want_labels else None), int(ncomp.item())


def _morph6(mask, dilate, iterations):
    assert mask.dim() == 3 and mask.dtype == torch.uint8
    if int(iterations) < 1:
        raise ValueError("iterations < 1 (repeat until stable) is not supported")
    d0, d1, d2 = mask.shape
    out = torch.empty_like(mask)
    tmp = torch.empty_like(mask) if iterations > 1 else None
    L.check(L.lib().bsg_binary_morph6(_ptr(mask), _ptr(out), _ptr(tmp) if tmp is not None else None, d0, d1, d2,
                                      1 if dilate else 0, int(iterations), L.stream_ptr()))
    return out


def binary_erosion(mask, iterations=1):
    """scipy.ndimage.binary_erosion(mask, iterations=iterations) (default 6-connected structure, border_value 0)."""
    return _morph6(mask, False, iterations)


def binary_dilation(mask, iterations=1):
    """scipy.ndimage.binary_dilation(mask, iterations=iterations)."""
    return _morph6(mask, True, iterations)


def mask_andnot(a, b=None):
    """a & ~b as a 0/1 uint8 volume (b None: a != 0)."""
    assert a.dtype == torch.uint8 and (b is None or (b.dtype == torch.uint8 and b.shape == a.shape))
    out = torch.empty_like(a)
    L.check(L.lib().bsg_mask_andnot(_ptr(a), _ptr(b) if b is not None else None, a.numel(), _ptr(out), L.stream_ptr()))
    return out


def distance_transform_edt(mask, sampling=None):
    """scipy.ndimage.distance_transform_edt(mask, sampling) -> cuda float64 tensor."""
    assert mask.dim() == 3 and mask.dtype == torch.uint8
    d0, d1, d2 = mask.shape
    out = torch.empty(mask.shape, dtype=torch.float64, device=mask.device)
    tmp = torch.empty_like(out)
    samp = None
    if sampling is not None:
        s = [float(v) for v in (sampling if np.ndim(sampling) else (sampling,) * 3)]
        assert len(s) == 3
        samp = (C.c_double * 3)(*s)
    L.check(L.lib().bsg_edt(_ptr(mask), d0, d1, d2, samp, _ptr(out), _ptr(tmp), L.stream_ptr()))
    return out


def surface_gradient_stats(mask, dist_in, dist_out):
    """(count, mean, std) of |np.gradient(dist_in - dist_out)| over mask & ~binary_erosion(mask) (two passes: the second
    one sums squared deviations from the mean, as np.std does)."""
    d0, d1, d2 = mask.shape
    out = torch.empty(3, dtype=torch.float64, device=mask.device)
    lib = L.lib()
    L.check(lib.bsg_surface_gradient_sums(_ptr(mask), _ptr(dist_in), _ptr(dist_out), d0, d1, d2, 0.0, _ptr(out),
                                          L.stream_ptr()))
    cnt, s1, _ = out.cpu().tolist()
    if cnt == 0:
        return 0, float("nan"), float("nan")
    mean = s1 / cnt
    L.check(lib.bsg_surface_gradient_sums(_ptr(mask), _ptr(dist_in), _ptr(dist_out), d0, d1, d2, mean, _ptr(out),
                                          L.stream_ptr()))
    _, r1, r2 = out.cpu().tolist()
    return int(cnt), mean + r1 / cnt, float(np.sqrt(max(r2 / cnt - (r1 / cnt) ** 2, 0.0)))


def intensity_moments(data, mask=None):
    """(count, mean, std, min, max) of data[mask > 0] (mask None: data[data > 0]); fp64 sums, np.std's two passes."""
    assert data.dtype == torch.float32 and (mask is None or (mask.dtype == torch.uint8 and mask.shape == data.shape))
    out = torch.empty(3, dtype=torch.float64, device=data.device)
    mm = torch.empty(2, dtype=torch.float32, device=data.device)
    lib = L.lib()
    mp = _ptr(mask) if mask is not None else None
    L.check(lib.bsg_intensity_moments(_ptr(data), mp, data.numel(), 0.0, _ptr(out), _ptr(mm), L.stream_ptr()))
    cnt, s1, _ = out.cpu().tolist()
    if cnt == 0:
        return 0, None, None, None, None
    mean = s1 / cnt
    L.check(lib.bsg_intensity_moments(_ptr(data), mp, data.numel(), mean, _ptr(out), _ptr(mm), L.stream_ptr()))
    _, r1, r2 = out.cpu().tolist()
    lo, hi = mm.cpu().tolist()
    return int(cnt), mean + r1 / cnt, float(np.sqrt(max(r2 / cnt - (r1 / cnt) ** 2, 0.0))), lo, hi


def _lerp(a, b, t):
    """numpy.lib._function_base_impl._lerp for scalars (float64)."""
    a, b, t = np.float64(a), np.float64(b), np.float64(t)
    d = b - a
    return b - d * (1 - t) if t >= 0.5 else a + d * t


class MaskedValues:
    """The values data[mask > 0] (or data[data > 0]) held on the device as order-preserving keys, for exact
    np.percentile / np.median queries by radix select."""

    def __init__(self, data, mask=None):
        assert data.dtype == torch.float32 and (mask is None or (mask.dtype == torch.uint8 and mask.shape == data.shape))
        self.keys = torch.empty(data.numel(), dtype=torch.int32, device=data.device)
        cnt = torch.zeros(1, dtype=torch.int64, device=data.device)
        L.check(L.lib().bsg_masked_compact_keys(_ptr(data), _ptr(mask) if mask is not None else None, data.numel(),
                                                _ptr(self.keys), _ptr(cnt), L.stream_ptr()))
        self.count = int(cnt.item())

    def order_stats(self, ranks):
        """values (float64 list) at the given 0-based ranks of the sorted selection"""
        lib = L.lib()
        out = []
        ranks = [int(r) for r in ranks]
        ws_bytes = lib.bsg_select_workspace_bytes()
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.keys.device)
        for i in range(0, len(ranks), 8):
            chunk = ranks[i:i + 8]
            res = torch.empty(8, dtype=torch.float32, device=self.keys.device)
            arr = (C.c_ulonglong * len(chunk))(*chunk)
            L.check(lib.bsg_select_ranks(_ptr(self.keys), self.count, arr, len(chunk), _ptr(res), _ptr(ws), ws_bytes,
                                         L.stream_ptr()))
            out += [float(v) for v in res.cpu().numpy()[:len(chunk)].astype(np.float64)]
        return out

    def percentiles(self, qs):
        """np.percentile(values, q) (method 'linear') for each q, bit-exact for float32-representable data"""
        n = self.count
        if n == 0:
            raise ValueError("percentile of an empty selection")
        plan = []
        for q in qs:
            quant = np.true_divide(np.float64(q), 100)
            vi = (n - 1) * quant  # numpy's 'linear' method: get_virtual_index = (n - 1) * quantiles
            prev = int(np.floor(vi))
            gamma = vi - prev
            lo = min(max(prev, 0), n - 1)
            hi = min(max(prev + 1, 0), n - 1)
            plan.append((lo, hi, gamma))
        vals = self.order_stats([r for lo, hi, _ in plan for r in (lo, hi)])
        return [float(_lerp(vals[2 * k], vals[2 * k + 1], g)) for k, (_, _, g) in enumerate(plan)]

    def median(self):
        """np.median(values)"""
        n = self.count
        if n % 2 == 1:
            return self.order_stats([n // 2])[0]
        a, b = self.order_stats([n // 2 - 1, n // 2])
        return float((np.float64(a) + np.float64(b)) / 2.0)


def masked_threshold_count(mask, x1=None, t1=0.0, x2=None, t2=0.0, x3=None, t3=0.0):
    """number of voxels with mask != 0 and x1 < t1 and x2 > t2 and x3 < t3 (None skips a test)"""
    cnt = torch.zeros(1, dtype=torch.int64, device=mask.device)
    p = lambda t: _ptr(t) if t is not None else None  # noqa: E731
    L.check(L.lib().bsg_masked_threshold_count(p(x1), p(x2), p(x3), _ptr(mask), mask.numel(), float(t1), float(t2),
                                               float(t3), _ptr(cnt), L.stream_ptr()))
    return int(cnt.item())
