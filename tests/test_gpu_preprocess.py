"""GPU parity of the case preprocessing (crop to the non-zero box, hole filling, masked z-score) against the CPU oracle.

Mask, hole filling and bounding box: bit-exact.  Normalised intensities: float32, within 2e-5 absolute (z-scores are
O(1); numpy reduces mean / std pairwise in float32, the kernel in fp64 — the oracle's own rounding is the larger term).
"""
import numpy as np
import pytest
import torch

from oracle import preprocess as OPP

pytestmark = pytest.mark.gpu

NORM_TOL = 2e-5


@pytest.fixture(scope="module")
def PP():
    from brainseg_b200 import preprocessing
    return preprocessing


@pytest.mark.parametrize("seed,shape", [(0, (64, 80, 72)), (1, (37, 45, 130)), (2, (155, 240, 240))])
def test_preprocess_matches_oracle(PP, seed, shape):
    data = OPP.synthetic_head(seed, shape)
    ref, seg, bbox = OPP.preprocess_case(data)
    got, props = PP.preprocess_case(data)
    assert props["crop_bbox"] == bbox
    assert tuple(props["size_after_cropping"]) == ref.shape[1:]
    assert np.array_equal(props["nonzero_mask"].cpu().numpy().astype(bool), seg[0] >= 0)
    g = got.cpu().numpy()
    assert g.dtype == np.float32 and g.shape == ref.shape
    err = float(np.abs(g - ref).max())
    print(f"bbox {bbox}, normalised max abs err {err:.3g}")
    assert err < NORM_TOL
    outside = np.broadcast_to(seg[0] < 0, g.shape)
    assert not g[outside].any() and not ref[outside].any()  # voxels outside the mask are exactly zero


def test_fill_holes_adversarial(PP):
    rng = np.random.default_rng(5)
    from scipy.ndimage import binary_fill_holes
    cases = [rng.random((20, 33, 70)) < 0.6, rng.random((9, 9, 9)) < 0.3, np.ones((5, 6, 7), bool), np.zeros((4, 4, 4), bool)]
    shell = np.zeros((12, 12, 12), bool)
    shell[2:10, 2:10, 2:10] = True
    shell[3:9, 3:9, 3:9] = False       # hollow box: interior is one hole
    shell[5, 5, 2] = False             # ...unless punctured: then nothing is filled
    cases += [shell.copy()]
    shell[5, 5, 2] = True
    cases += [shell]
    diag = np.zeros((6, 6, 6), bool)   # background cells touching only diagonally are NOT 6-connected
    diag[1:5, 1:5, 1:5] = True
    diag[2, 2, 2] = False
    diag[1, 1, 1] = False
    cases += [diag]
    for m in cases:
        vol = torch.from_numpy(m.astype(np.float32)[None]).cuda().contiguous()
        got = PP.nonzero_mask(vol).cpu().numpy().astype(bool)
        assert np.array_equal(got, binary_fill_holes(m))


def test_uncrop_and_mask_free_normalisation(PP):
    data = OPP.synthetic_head(3, (40, 50, 44))
    ref, seg, bbox = OPP.preprocess_case(data, use_mask_for_norm=False)
    got, props = PP.preprocess_case(data, use_mask_for_norm=False)
    assert float(np.abs(got.cpu().numpy() - ref).max()) < NORM_TOL
    lab = torch.randint(0, 4, props["size_after_cropping"], dtype=torch.uint8, device="cuda")
    full = PP.uncrop_segmentation(lab, props).cpu().numpy()
    assert full.shape == data.shape[1:]
    sl = tuple(slice(a, b) for a, b in bbox)
    assert np.array_equal(full[sl], lab.cpu().numpy())
    full[sl] = 0
    assert not full.any()
