"""CPU restatement (numpy / scipy) of nnU-Net v1 case preprocessing for the BraTS plans — TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the code lives in the un-vendored `Brats21_KAIST_MRI_Lab/nnunet` (nnunet/preprocessing/cropping.py
`create_nonzero_mask`, `get_bbox_from_mask`, `crop_to_nonzero`; nnunet/preprocessing/preprocessing.py
`GenericPreprocessor.resample_and_normalize`, "nonCT" branch).  Restated from the published
algorithm; the call site in the reference is run_brats2021_inference_singlethread.py:89 (`trainer.preprocess_patient`).
"""
import numpy as np
from scipy.ndimage import binary_fill_holes


def create_nonzero_mask(data):
    """data (C, Z, Y, X): union over modalities of != 0, holes filled (scipy default 6-connected structure)."""
    mask = np.zeros(data.shape[1:], dtype=bool)
    for c in range(data.shape[0]):
        mask = mask | (data[c] != 0)
    return binary_fill_holes(mask)


def get_bbox_from_mask(mask, outside_value=0):
    idx = np.where(mask != outside_value)
    return [[int(np.min(idx[a])), int(np.max(idx[a])) + 1] for a in range(3)]


def crop_to_nonzero(data):
    """Returns (cropped data, seg with -1 outside the mask / 0 inside, bbox)."""
    mask = create_nonzero_mask(data)
    bbox = get_bbox_from_mask(mask, 0)
    sl = tuple(slice(a, b) for a, b in bbox)
    cropped = np.stack([data[c][sl] for c in range(data.shape[0])])
    m = mask[sl]
    seg = np.where(m, 0, -1).astype(np.int16)[None]
    return cropped, seg, bbox


def normalize_nonct(data, seg, use_mask_for_norm=True):
    """GenericPreprocessor.resample_and_normalize, scheme != CT / CT2: per-channel z-score over the mask, float32."""
    data = data.astype(np.float32, copy=True)
    for c in range(data.shape[0]):
        mask = (seg[-1] >= 0) if use_mask_for_norm else np.ones(seg.shape[1:], dtype=bool)
        data[c][mask] = (data[c][mask] - data[c][mask].mean()) / (data[c][mask].std() + 1e-8)
        data[c][mask == 0] = 0
    return data


def preprocess_case(data, use_mask_for_norm=True):
    cropped, seg, bbox = crop_to_nonzero(data)
    return normalize_nonct(cropped, seg, use_mask_for_norm), seg, bbox


def synthetic_head(seed=0, shape=(64, 80, 72), channels=4):
    """A brain-like test volume: an ellipsoid of noisy tissue (some exact zeros inside = holes to fill, an enclosed
    cavity, a notch open to the outside) in a zero background."""
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(*[np.linspace(-1, 1, s) for s in shape], indexing="ij")
    inside = (z / 0.8) ** 2 + (y / 0.7) ** 2 + (x / 0.75) ** 2 < 1.0
    data = np.zeros((channels,) + shape, dtype=np.float32)
    for c in range(channels):
        v = (rng.standard_normal(shape) * 50 + 300 * (c + 1)).astype(np.float32)
        v[rng.random(shape) < 0.02] = 0.0  # speckle zeros (only a hole if zero in every modality)
        data[c] = np.where(inside, v, 0.0)
    cav = (z + 0.35) ** 2 + (y + 0.05) ** 2 + x ** 2 < 0.03  # enclosed cavity: filled
    data[:, cav] = 0.0
    data[:, (rng.random(shape) < 0.003) & inside] = 0.0  # isolated voxels that are zero in every modality: filled
    notch = (np.abs(y) < 0.05) & (np.abs(x) < 0.05) & (z > 0.2)  # channel to the outside: stays background
    data[:, notch] = 0.0
    return data
