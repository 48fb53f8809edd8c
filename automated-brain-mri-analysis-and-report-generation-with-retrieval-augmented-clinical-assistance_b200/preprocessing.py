"""Device-side case preprocessing: the step right before predict_3D (`trainer.preprocess_patient`,
run_brats2021_inference_singlethread.py:89; UPSTREAM nnU-Net v1 `crop_to_nonzero` + `GenericPreprocessor` with the
BraTS plans: "nonCT" z-score inside the non-zero mask, `use_mask_for_norm=True`, 1 mm isotropic spacing so no
resampling — SURVEY.md Appendix A.8) and its inverse for export (`save_segmentation_nifti_from_softmax`, App. A.7).

Cropping to the brain's bounding box cuts a median BraTS case from 18 sliding-window tiles to about 8.
Everything runs through libbrainseg_b200.so; there is no CPU implementation.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import voxelops as V


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def nonzero_mask(vol, fill_holes=True):
    """`create_nonzero_mask`: any modality != 0, then scipy.ndimage.binary_fill_holes.  vol: cuda fp32 (C, Z, Y, X);
    returns a cuda uint8 (Z, Y, X) mask."""
    assert vol.dim() == 4 and vol.is_cuda and vol.dtype == torch.float32 and vol.is_contiguous()
    Cn, Z, Y, X = vol.shape
    lib = L.lib()
    mask = torch.empty((Z, Y, X), dtype=torch.uint8, device=vol.device)
    L.check(lib.bsg_nonzero_mask(_ptr(vol), Cn, Z, Y, X, _ptr(mask), L.stream_ptr()))
    if fill_holes:
        n = Z * Y * X
        ws_bytes = lib.bsg_ccl26_workspace_bytes(Z, Y, X)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=vol.device)
        labels = torch.empty((Z, Y, X), dtype=torch.int32, device=vol.device)
        ncomp = torch.zeros(1, dtype=torch.int32, device=vol.device)
        flags = torch.empty(n // 2 + 2, dtype=torch.uint8, device=vol.device)
        L.check(lib.bsg_fill_holes_u8(_ptr(mask), Z, Y, X, _ptr(labels), _ptr(ncomp), _ptr(flags), flags.numel(), _ptr(ws),
                                      ws_bytes, L.stream_ptr()))
    return mask


def bounding_box(mask):
    """`get_bbox_from_mask(mask, 0)`: [[zmin, zmax+1], [ymin, ymax+1], [xmin, xmax+1]] (None for an empty mask)."""
    m = V.masked_moments(mask, [V.bits_of(1)])[0]
    if int(m["count"]) == 0:
        return None
    return [[int(m["mn0"]), int(m["mx0"]) + 1], [int(m["mn1"]), int(m["mx1"]) + 1], [int(m["mn2"]), int(m["mx2"]) + 1]]


def preprocess_case(data, use_mask_for_norm=True):
    """data: (C, Z, Y, X) float32, numpy / torch, host or device — the stacked modalities in nnU-Net array order.

    Returns (d, properties): `d` the cropped, per-channel z-scored cuda fp32 tensor handed to predict_3D, and
    `properties` with `crop_bbox`, `original_size_of_raw_data`, `size_after_cropping` (upstream key names) and the
    cropped `nonzero_mask` (cuda uint8) — what the export step needs to paste the result back."""
    dev = V.device()
    if isinstance(data, np.ndarray):
        data = torch.from_numpy(np.ascontiguousarray(data))
    vol = data.to(dev, torch.float32, non_blocking=True).contiguous()
    Cn, Z, Y, X = vol.shape
    lib = L.lib()
    mask = nonzero_mask(vol)
    bbox = bounding_box(mask)
    if bbox is None:  # all-zero image: upstream would fail in get_bbox_from_mask; keep the whole volume
        bbox = [[0, Z], [0, Y], [0, X]]
    (z0, z1), (y0, y1), (x0, x1) = bbox
    cz, cy, cx = z1 - z0, y1 - y0, x1 - x0
    n = Z * Y * X
    # with the mask driving the normalisation the statistics run over the mask (all inside the box); without it upstream
    # normalises over the whole cropped array.  `use_mask_for_norm` may differ per channel (the plans hold a dict).
    flags = [bool(use_mask_for_norm)] * Cn if isinstance(use_mask_for_norm, (bool, int)) else [bool(v) for v in use_mask_for_norm]
    if len(flags) != Cn:
        raise ValueError(f"use_mask_for_norm has {len(flags)} entries for {Cn} channels")
    box_mask = None
    if not all(flags):
        box_mask = torch.zeros_like(mask)
        box_mask[z0:z1, y0:y1, x0:x1] = 1
    out = torch.empty((Cn, cz, cy, cx), dtype=torch.float32, device=dev)
    mean_std_all = np.zeros((Cn, 2), dtype=np.float32)
    # channels sharing a setting go through one call each (all four BraTS modalities use the mask: one call)
    runs, c0 = [], 0
    for c in range(1, Cn + 1):
        if c == Cn or flags[c] != flags[c0]:
            runs.append((c0, c))
            c0 = c
    for (ca, cb) in runs:
        norm_mask = mask if flags[ca] else box_mask
        nc = cb - ca
        sums = torch.empty(nc * 3, dtype=torch.float64, device=dev)
        L.check(lib.bsg_masked_channel_stats(_ptr(vol[ca:cb]), nc, n, _ptr(norm_mask), _ptr(sums), L.stream_ptr()))
        s = sums.cpu().numpy().reshape(nc, 3)
        cnt = np.maximum(s[:, 2], 1.0)
        mean = s[:, 0] / cnt
        var = np.maximum(s[:, 1] / cnt - mean * mean, 0.0)
        ms = np.stack([mean, np.sqrt(var)], axis=1).astype(np.float32)
        mean_std_all[ca:cb] = ms
        mean_std = torch.from_numpy(ms).to(dev)
        L.check(lib.bsg_crop_normalize(_ptr(vol[ca:cb]), nc, Z, Y, X, _ptr(norm_mask), z0, y0, x0, cz, cy, cx, _ptr(mean_std),
                                       _ptr(out[ca:cb]), None, L.stream_ptr()))
    mask_c = mask[z0:z1, y0:y1, x0:x1].contiguous()
    props = {"crop_bbox": bbox, "original_size_of_raw_data": np.array([Z, Y, X]),
             "size_after_cropping": (cz, cy, cx), "nonzero_mask": mask_c, "channel_mean_std": mean_std_all}
    return out, props


def uncrop_segmentation(seg, properties):
    """Paste a (z, y, x) label volume of the cropped geometry back into zeros of the original size
    (`save_segmentation_nifti_from_softmax`: `seg_old_size[bbox] = seg`)."""
    Z, Y, X = (int(v) for v in properties["original_size_of_raw_data"])
    (z0, z1), (y0, y1), (x0, x1) = properties["crop_bbox"]
    if not torch.is_tensor(seg):
        seg = torch.from_numpy(np.ascontiguousarray(seg))
    out = torch.zeros((Z, Y, X), dtype=seg.dtype, device=seg.device)
    out[z0:z1, y0:y1, x0:x1] = seg
    return out
