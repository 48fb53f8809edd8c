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
