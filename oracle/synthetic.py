"""Seeded synthetic inputs shared by the tests, smoke() and bench.py."""
import numpy as np
import torch
from scipy.ndimage import gaussian_filter


def case_volume(seed=0, shape=(4, 155, 240, 240)):
    """BASELINE config 1/2 input: randn(4,155,240,240) fp32, array order (C, z, y, x)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32).numpy()


def label_volume(seed=0, shape=(240, 240, 155), sigma=6.0):
    """Blobby label volume: smoothed noise thresholded at 1.5/2.0/2.5 sigma -> labels 1/2/3 (≈6 % tumour)."""
    rng = np.random.default_rng(seed)
    f = gaussian_filter(rng.standard_normal(shape), sigma)
    f = (f - f.mean()) / f.std()
    lab = np.zeros(shape, dtype=np.uint8)
    lab[f > 1.5] = 1
    lab[f > 2.0] = 2
    lab[f > 2.5] = 3
    return lab


def label_pair(seed=0, shape=(240, 240, 155)):
    """(prediction, ground truth) pair: the prediction is the GT rolled by 3 voxels along axis 0."""
    gt = label_volume(seed, shape)
    return np.roll(gt, 3, axis=0).copy(), gt


def mri_volumes(seed, seg):
    """Four MRI-like modalities for a label volume: an ellipsoidal "head" of smooth positive texture (zero outside),
    the tumour labels scale the signal per modality.  Integer-valued float32 (like int16 NIfTI data), so order
    statistics meet ties and every value is exact in float32 and float64."""
    rng = np.random.default_rng(1000 + seed)
    shape = seg.shape
    grids = np.meshgrid(*[np.linspace(-1.0, 1.0, s) for s in shape], indexing="ij")
    head = sum(g ** 2 for g in grids) < 2.2
    gains = {"t1": (0.6, 1.0, 0.9), "t1ce": (0.7, 1.0, 1.8), "t2": (1.9, 1.5, 1.2), "flair": (0.8, 1.6, 1.3)}
    out = {}
    for name, per_label in gains.items():
        tex = gaussian_filter(rng.standard_normal(shape), 2.0)
        tex = 400.0 + 120.0 * tex / tex.std()
        factor = np.ones(shape)
        for lab, gain in zip((1, 2, 3), per_label):
            factor[seg == lab] = gain
        out[name] = (np.round(np.clip(tex * factor, 1.0, None)) * head).astype(np.float32)
    return out
