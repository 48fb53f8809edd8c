"""ctypes binding of libbrainseg_b200.so (the C ABI in include/brainseg_b200.h).

There is deliberately no fallback: if the shared library is missing or the device is not sm_100 the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# the library is built next to the import shim (brainseg_b200/), i.e. under a short in-tree path (see build.py)
LIB_PATH = os.path.join(os.path.dirname(_HERE), "brainseg_b200", "libbrainseg_b200.so")

BSG_CONV_K3, BSG_CONVT_K2S2, BSG_CONV_K1 = 0, 1, 2
BSG_ACT_NONE, BSG_ACT_LRELU = 0, 1


class BsgError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("stride", C.c_int),
        ("N", C.c_int), ("D", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("cin", C.c_int), ("in_ptr", C.c_void_p), ("in_ctot", C.c_int),
        ("cout", C.c_int), ("out_ptr", C.c_void_p), ("out_ctot", C.c_int), ("out_coff", C.c_int),
        ("weights", C.c_void_p), ("bias", C.c_void_p),
        ("act", C.c_int), ("slope", C.c_float), ("stats", C.c_void_p), ("out_f16", C.c_int),
        ("use_khshift", C.c_int), ("max_ctas", C.c_int), ("in_f16", C.c_int), ("algo", C.c_int), ("pair", C.c_int),
        ("overflow", C.c_void_p), ("in_norm", C.c_void_p), ("in_norm_c", C.c_int), ("in_norm_cc", C.c_int),
        ("out_split_stride", C.c_int), ("kw_taps", C.c_int), ("tma_store", C.c_int), ("mblock", C.c_int),
    ]


class ConvInfo(C.Structure):
    _fields_ = [
        ("bw", C.c_int), ("bh", C.c_int), ("bd", C.c_int), ("bn", C.c_int),
        ("ntile", C.c_int), ("n_ntiles", C.c_int), ("cc", C.c_int), ("nstages", C.c_int),
        ("khshift", C.c_int), ("grid", C.c_int), ("smem_bytes", C.c_size_t), ("flops", C.c_double),
    ]


_lib = None


def _declare(lib):
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    lib.bsg_version.restype = i
    lib.bsg_last_error.restype = C.c_size_t
    lib.bsg_last_error.argtypes = [C.c_char_p, C.c_size_t]
    lib.bsg_check_device.restype = i
    lib.bsg_sm_count.restype = i
    lib.bsg_conv_plan_create.restype = i
    lib.bsg_conv_plan_create.argtypes = [C.POINTER(ConvDesc), C.POINTER(vp)]
    lib.bsg_conv_plan_run.restype = i
    lib.bsg_conv_plan_run.argtypes = [vp, vp]
    lib.bsg_conv_plan_destroy.restype = None
    lib.bsg_conv_plan_destroy.argtypes = [vp]
    lib.bsg_conv_plan_info.restype = i
    lib.bsg_conv_plan_info.argtypes = [vp, C.POINTER(ConvInfo)]
    for name, sig in _EXTRA_SIGS.items():
        fn = getattr(lib, name)
        fn.restype = i
        fn.argtypes = sig
    lib.bsg_ccl26_workspace_bytes.restype = C.c_size_t
    lib.bsg_ccl26_workspace_bytes.argtypes = [i, i, i]
    lib.bsg_select_workspace_bytes.restype = C.c_size_t
    lib.bsg_select_workspace_bytes.argtypes = []
    lib.bsg_conv_desc_size.restype = C.c_size_t
    lib.bsg_conv_desc_size.argtypes = []
    if lib.bsg_conv_desc_size() != C.sizeof(ConvDesc):
        raise BsgError(f"bsg_conv_desc layout mismatch: library {lib.bsg_conv_desc_size()} bytes, binding "
                       f"{C.sizeof(ConvDesc)} (stale libbrainseg_b200.so? rebuild with __graft_entry__.build())")


_vp, _i, _sz, _u32, _f, _d = C.c_void_p, C.c_int, C.c_size_t, C.c_uint32, C.c_float, C.c_double
# name -> argtypes for the flat (non-plan) entry points (postproc.cu, tail.cu); all return int
_EXTRA_SIGS = {
    "bsg_label_lut_u8": [_vp, _vp, _sz, C.c_char_p, _vp],
    "bsg_label_pair_round_u8": [_vp, _vp, _vp, _sz, C.c_char_p, _vp],
    "bsg_label_pair_round_hist_u8": [_vp, _vp, _vp, _vp, _sz, C.c_char_p, _vp, _vp, _vp],
    "bsg_round_to_u8": [_vp, _i, _vp, _sz, _vp],
    "bsg_joint_hist_u8": [_vp, _vp, _sz, _vp, _vp, _vp],
    "bsg_ccl26_stats": [_vp, _i, _i, _i, _u32, _vp, _vp, _vp, _i, _vp, _sz, _vp],
    "bsg_ccl_stats": [_vp, _i, _i, _i, _u32, _i, _vp, _vp, _vp, _i, _vp, _sz, _vp],
    "bsg_nonzero_mask": [_vp, _i, _i, _i, _i, _vp, _vp],
    "bsg_fill_holes_u8": [_vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp, _sz, _vp],
    "bsg_masked_channel_stats": [_vp, _i, _sz, _vp, _vp, _vp],
    "bsg_crop_normalize": [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "bsg_masked_moments": [_vp, _i, _i, _i, C.POINTER(_u32), _i, _u32, _vp, _vp],
    "bsg_binary_morph6": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "bsg_mask_andnot": [_vp, _vp, _sz, _vp, _vp],
    "bsg_edt": [_vp, _i, _i, _i, C.POINTER(_d), _vp, _vp, _vp],
    "bsg_surface_gradient_sums": [_vp, _vp, _vp, _i, _i, _i, _d, _vp, _vp],
    "bsg_intensity_moments": [_vp, _vp, _sz, _d, _vp, _vp, _vp],
    "bsg_masked_compact_keys": [_vp, _vp, _sz, _vp, _vp, _vp],
    "bsg_select_ranks": [_vp, _sz, C.POINTER(C.c_ulonglong), _i, _vp, _vp, _sz, _vp],
    "bsg_masked_threshold_count": [_vp, _vp, _vp, _vp, _sz, _d, _d, _d, _vp, _vp],
    "bsg_gather_patch_tta": [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, C.POINTER(_i), _i, _vp, _i, _i, _i, _vp],
    "bsg_norm_finalize": [_vp, _i, _i, _i, _d, _f, _vp, _vp, _vp, _vp],
    "bsg_norm_finalize_table": [_vp, _i, _i, _i, _d, _f, _vp, _vp, _f, _vp, _i, _i, _vp],
    "bsg_norm_apply_lrelu_split": [_vp, _sz, _i, _i, _i, _i, _vp, _f, _vp],
    "bsg_head_tta_accumulate_split": [_vp, _i, _i, _i, _i, _i, C.POINTER(_i), _i, _f, C.POINTER(_f), C.POINTER(_f), _i, _i,
                                      _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "bsg_norm_apply_lrelu": [_vp, _sz, _i, _i, _i, _i, _vp, _f, _i, _i, _vp],
    "bsg_head_tta_accumulate": [_vp, _i, _i, _i, _i, _i, _i, C.POINTER(_i), _i, _f, C.POINTER(_f), C.POINTER(_f), _i, _i,
                                _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _f, _vp],
    "bsg_finalize": [C.POINTER(_vp), _i, _vp, _i, _sz, _i, C.POINTER(_i), _vp, _vp, _vp],
    "bsg_finalize_peer": [_vp, _i, _i, _vp, _i, _sz, _sz, _sz, _i, C.POINTER(_i), _vp, _i, _vp],
    "bsg_finalize_peer_signal": [_vp, _i, _i, _vp, _i, _sz, _sz, _sz, _i, C.POINTER(_i), _vp, _i, _vp, _i, _u32, _vp],
    "bsg_enable_peer_access": [_i],
    "bsg_ipc_export": [_vp, _vp, C.POINTER(_sz)],
    "bsg_ipc_open": [_vp, C.POINTER(_vp)],
    "bsg_ipc_close": [_vp],
    "bsg_nccl_unique_id": [_vp],
    "bsg_nccl_comm_create": [_vp, _i, _i, C.POINTER(_vp)],
    "bsg_nccl_reduce_accumulator": [_vp, _vp, _sz, _i, _vp],
    "bsg_nccl_comm_destroy": [_vp],
}


def lib():
    """Returns the loaded library; raises if it was never built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BsgError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                "brainseg_b200 has no CPU fallback")
        l = C.CDLL(LIB_PATH)
        _declare(l)
        _lib = l
    return _lib


def last_error():
    buf = C.create_string_buffer(512)
    lib().bsg_last_error(buf, 512)
    return buf.value.decode()


def check(rc):
    if rc != 0:
        raise BsgError(f"brainseg_b200 error {rc}: {last_error()}")


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


class ConvPlan:
    """Owns one bsg_conv_plan (tensor maps + launch geometry bound to fixed device buffers)."""

    def __init__(self, **kw):
        d = ConvDesc()
        d.use_khshift = -1
        d.algo = -1
        d.pair = -1
        for k, v in kw.items():
            setattr(d, k, v)
        self._h = C.c_void_p()
        check(lib().bsg_conv_plan_create(C.byref(d), C.byref(self._h)))
        self.desc = d

    def run(self, stream=None):
        check(lib().bsg_conv_plan_run(self._h, stream_ptr(stream)))

    def info(self):
        inf = ConvInfo()
        check(lib().bsg_conv_plan_info(self._h, C.byref(inf)))
        return inf

    def __del__(self):
        try:
            if self._h:
                lib().bsg_conv_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass
