"""GPU parity of the morphology / EDT / intensity-statistics kernels (csrc/morph.cu, through the C ABI) against SciPy,
NumPy, the CPU oracle and the reference's own results (tests/golden/voxelops.*).

Bit-exact: morphology, labellings, isotropic EDT, order statistics / percentiles / median / min / max / counts.
fp64 reductions (means, standard deviations, and the scores built from them): relative 1e-11 — NumPy sums pairwise
in raster order, the kernels sum per thread and then atomically; both carry ~1e-16 * sqrt(n) rounding noise.
Anisotropic EDT: relative 1e-14 (SciPy squares and adds the same three terms; ties between equidistant background
voxels may round differently)."""
import json
import os

import numpy as np
import pytest
import torch
from scipy import ndimage as ndi

from oracle import intensity as OI
from oracle import postproc as OP
from oracle import synthetic as SY
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu

SUM_RTOL = 1e-11


def tree_close(a, b, path=""):
    if isinstance(a, dict):
        assert set(a.keys()) == set(b.keys()), f"{path}: {set(a.keys()) ^ set(b.keys())}"
        for k in a:
            tree_close(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            tree_close(x, y, f"{path}[{i}]")
    elif isinstance(a, (float, np.floating)) and b is not None:
        assert abs(float(a) - float(b)) <= SUM_RTOL * max(abs(float(a)), abs(float(b))), f"{path}: {a!r} != {b!r}"
    else:
        assert a == b, f"{path}: {a!r} != {b!r}"


@pytest.fixture(scope="module")
def V():
    from brainseg_b200 import voxelops
    return voxelops


def golden_case(seed):
    vols = np.load(os.path.join(GOLDEN, "voxelops.npz"))
    with open(os.path.join(GOLDEN, "voxelops.json")) as f:
        ref = json.load(f)[str(seed)]
    seg = vols[f"seg{seed}"]
    mri = {k: vols[f"{k}{seed}"].astype(np.float32) for k in ("t1", "t1ce", "t2", "flair")}
    return vols, ref, seg, mri


def random_mask(seed, shape, p=0.5, smooth=1.5):
    rng = np.random.default_rng(seed)
    f = ndi.gaussian_filter(rng.standard_normal(shape), smooth) if smooth else rng.standard_normal(shape)
    return f > np.quantile(f, 1 - p)


MORPH_SHAPES = [(48, 40, 36), (33, 17, 70), (1, 9, 12), (7, 1, 1), (5, 5, 5), (64, 64, 130)]


@pytest.mark.parametrize("shape", MORPH_SHAPES)
def test_binary_morphology_matches_scipy(V, shape):
    for seed, p in ((0, 0.5), (1, 0.1), (2, 0.95)):
        m = random_mask(seed, shape, p)
        g = V.as_mask(m)
        for it in (1, 2, 3, 5):
            assert np.array_equal(V.binary_erosion(g, it).cpu().numpy().astype(bool), ndi.binary_erosion(m, iterations=it))
            assert np.array_equal(V.binary_dilation(g, it).cpu().numpy().astype(bool), ndi.binary_dilation(m, iterations=it))
    ones = np.ones(shape, bool)
    assert np.array_equal(V.binary_erosion(V.as_mask(ones)).cpu().numpy().astype(bool), ndi.binary_erosion(ones))
    assert int(V.binary_dilation(V.as_mask(np.zeros(shape, bool))).sum()) == 0
    a, b = random_mask(3, shape), random_mask(4, shape)
    assert np.array_equal(V.mask_andnot(V.as_mask(a), V.as_mask(b)).cpu().numpy().astype(bool), a & ~b)
    with pytest.raises(ValueError):
        V.binary_erosion(g, 0)


@pytest.mark.parametrize("seed", [0, 1])
def test_morphology_edt_labels_golden(V, seed):
    vols, ref, seg, _ = golden_case(seed)
    wt = V.as_mask(seg > 0)
    for k, want in ref["erode"].items():
        assert int(V.binary_erosion(wt, int(k)).sum()) == want
    for k, want in ref["dilate"].items():
        assert int(V.binary_dilation(wt, int(k)).sum()) == want
    assert np.array_equal(V.binary_dilation(wt, 5).cpu().numpy().astype(bool), vols[f"dilate5_{seed}"])
    assert np.array_equal(V.binary_erosion(wt, 2).cpu().numpy().astype(bool), vols[f"erode2_{seed}"])
    edt = V.distance_transform_edt(wt).cpu().numpy()
    assert np.array_equal(edt, np.sqrt(vols[f"edt_in_sq_{seed}"].astype(np.float64)))  # bit-exact
    if seed == 0:
        out = V.distance_transform_edt(V.as_mask(seg == 0), sampling=(0.9, 1.1, 1.25)).cpu().numpy()
        np.testing.assert_allclose(out, vols["edt_out_aniso_0"], rtol=1e-14, atol=0)
    vol = torch.from_numpy(seg).cuda()
    for conn in (6, 18, 26):
        lab, n = V.ccl(vol, V.bits_of(3), conn)
        assert n == ref[f"label{conn}_count"]
        assert np.array_equal(lab.cpu().numpy(), vols[f"label{conn}_{seed}"])


@pytest.mark.parametrize("shape", [(33, 17, 70), (1, 9, 12), (7, 1, 1), (40, 300, 20), (64, 64, 130)])
def test_edt_matches_scipy(V, shape):
    for seed, p in ((0, 0.5), (1, 0.9), (2, 0.999)):
        m = random_mask(seed, shape, p)
        if m.all():
            m.flat[m.size // 2] = False
        got = V.distance_transform_edt(V.as_mask(m)).cpu().numpy()
        assert np.array_equal(got, ndi.distance_transform_edt(m)), f"{shape} p={p}"
        got = V.distance_transform_edt(V.as_mask(m), sampling=(2.0, 0.5, 1.0)).cpu().numpy()
        assert np.array_equal(got, ndi.distance_transform_edt(m, sampling=(2.0, 0.5, 1.0)))  # exact binary fractions
        got = V.distance_transform_edt(V.as_mask(m), sampling=(0.7, 1.3, 1.9)).cpu().numpy()
        np.testing.assert_allclose(got, ndi.distance_transform_edt(m, sampling=(0.7, 1.3, 1.9)), rtol=1e-14, atol=0)
    assert float(V.distance_transform_edt(V.as_mask(np.zeros(shape, bool))).abs().max()) == 0.0
    assert torch.isinf(V.distance_transform_edt(V.as_mask(np.ones(shape, bool)))).all()  # documented: no background


def test_edt_and_labels_full_size(V):
    seg = SY.label_volume(3, (240, 240, 155))
    wt = seg > 0
    g = V.as_mask(wt)
    assert np.array_equal(V.distance_transform_edt(g).cpu().numpy(), ndi.distance_transform_edt(wt))
    out = V.distance_transform_edt(V.as_mask(~wt)).cpu().numpy()
    assert np.array_equal(out, ndi.distance_transform_edt(~wt))
    assert np.array_equal(V.binary_dilation(g, 5).cpu().numpy().astype(bool), ndi.binary_dilation(wt, iterations=5))
    vol = torch.from_numpy(seg).cuda()
    for conn, rank in ((6, 1), (18, 2)):
        lab, n = V.ccl(vol, V.MASK_GT0, conn)
        want, wn = ndi.label(wt, structure=ndi.generate_binary_structure(3, rank))
        assert n == wn and np.array_equal(lab.cpu().numpy(), want)


def test_order_statistics_exact(V):
    rng = np.random.default_rng(5)
    cases = [
        rng.standard_normal(100003).astype(np.float32) * 50,                       # negatives, no ties
        np.round(rng.standard_normal(250000) * 20).astype(np.float32),              # heavy ties, both signs, zeros
        np.abs(rng.standard_normal(7)).astype(np.float32),                          # tiny
        np.full(1000, 3.25, np.float32),                                           # all equal
        np.array([5.0], np.float32),                                               # one value
        np.concatenate([np.zeros(10, np.float32), -np.zeros(10, np.float32), [1e-40, -1e-40, 3e38, -3e38]]).astype(np.float32),
    ]
    qs = [0, 5, 10, 20, 25, 37.5, 50, 75, 85, 99.9, 100]
    for k, a in enumerate(cases):
        d = torch.from_numpy(a).cuda()
        mask = torch.ones(a.shape, dtype=torch.uint8, device="cuda")
        sel = V.MaskedValues(d, mask)
        assert sel.count == a.size
        a64 = a.astype(np.float64)
        want = [float(np.percentile(a64, q)) for q in qs]
        got = sel.percentiles(qs)
        assert got == want, f"case {k}: {got} != {want}"
        assert sel.median() == float(np.median(a64))
        s = np.sort(a)
        ranks = sorted(set([0, a.size // 3, a.size // 2, a.size - 1]))
        assert sel.order_stats(ranks) == [float(s[r]) for r in ranks]
        cnt, mean, std, lo, hi = V.intensity_moments(d, mask)
        assert cnt == a.size and lo == float(a.min()) and hi == float(a.max())
        assert abs(mean - a64.mean()) <= SUM_RTOL * max(abs(a64.mean()), np.abs(a64).mean())
        assert abs(std - a64.std()) <= SUM_RTOL * max(a64.std(), 1e-30) or a64.std() == 0 and std < 1e-12
    # data > 0 selection (mask None) and a sparse mask on a volume
    vol = np.round(rng.standard_normal((40, 50, 60)) * 100).astype(np.float32)
    d = torch.from_numpy(vol).cuda()
    sel = V.MaskedValues(d)
    pos = vol[vol > 0].astype(np.float64)
    assert sel.count == pos.size and sel.percentiles([5, 50]) == [float(np.percentile(pos, 5)), float(np.percentile(pos, 50))]
    m = rng.random(vol.shape) < 0.01
    sel = V.MaskedValues(d, torch.from_numpy(m.astype(np.uint8)).cuda())
    assert sel.percentiles([25, 75]) == [float(np.percentile(vol[m].astype(np.float64), q)) for q in (25, 75)]
    t = (12.5, -3.0, 40.25)
    want = int(((vol[m] < t[0]) & (vol[m] > t[1]) & (vol[m] < t[2])).sum())
    assert V.masked_threshold_count(torch.from_numpy(m.astype(np.uint8)).cuda(), d, t[0], d, t[1], d, t[2]) == want
    with pytest.raises(ValueError):
        V.MaskedValues(d, torch.zeros(vol.shape, dtype=torch.uint8, device="cuda")).percentiles([50])


@pytest.mark.parametrize("seed", [0, 1])
def test_reference_functions_golden(seed):
    """The drop-in feature_extraction functions against the reference's own outputs."""
    from brainseg_b200.feature_extraction import step4_morphology as S4
    from brainseg_b200.feature_extraction import utils as U

    _, ref, seg, mri = golden_case(seed)
    lv = U.LabelVolume(seg)
    masks = U.get_tumor_masks(lv)
    for tag, vd in (("iso", (1.0, 1.0, 1.0)), ("aniso", (0.9, 1.1, 1.25))):
        tree_close(S4.analyze_border_regularity(masks["wt"], vd), ref[f"border_{tag}"], f"border_{tag}")
        tree_close(S4.analyze_margin_definition(mri["t1ce"], lv, masks, vd), ref[f"margin_{tag}"], f"margin_{tag}")
        tree_close(S4.analyze_cystic_vs_solid(mri["t1"], mri["t2"], mri["flair"], lv, masks, vd), ref[f"cystic_{tag}"],
                   f"cystic_{tag}")
    for key, want in ref["stats"].items():
        mod, reg = key.split("_")
        got = U.get_intensity_stats(mri[mod], masks[reg])
        tree_close(got, want, key)
        for k in ("min", "max", "median", "q25", "q75", "voxel_count"):  # exact, not merely close
            assert got[k] == want[k], (key, k)
    tree_close(U.get_intensity_stats(mri["t1"], np.zeros(seg.shape, bool)), ref["stats_empty"], "empty")
    for mod, want in ref["normal_brain"].items():
        tree_close(U.get_normal_brain_stats(mri[mod], lv), want, f"normal_{mod}")
    for p, want in ref["brain_mask_count"].items():
        assert int(U.get_brain_mask(mri["t2"], float(p)).sum()) == want


def test_reference_functions_vs_oracle_odd_shapes():
    """Same functions against the CPU oracle on other shapes, incl. empty / tiny tumours and a tumour at the border."""
    from brainseg_b200.feature_extraction import step4_morphology as S4
    from brainseg_b200.feature_extraction import utils as U

    vd = (1.0, 1.0, 1.0)
    for seed, shape in ((4, (33, 47, 29)), (5, (64, 64, 40))):
        seg, _ = SY.label_pair(seed, shape)
        seg[:2, :5, :5] = 1  # tumour touching the volume corner
        mri = SY.mri_volumes(seed, seg)
        m64 = {k: v.astype(np.float64) for k, v in mri.items()}
        om = OP.get_tumor_masks(seg.astype(np.float64))
        lv = U.LabelVolume(seg)
        gm = U.get_tumor_masks(lv)
        tree_close(S4.analyze_border_regularity(gm["wt"], vd), OI.analyze_border_regularity(om["wt"], vd))
        tree_close(S4.analyze_margin_definition(mri["t1ce"], lv, gm, vd),
                   OI.analyze_margin_definition(m64["t1ce"], seg, om, vd))
        tree_close(S4.analyze_cystic_vs_solid(mri["t1"], mri["t2"], mri["flair"], lv, gm, vd),
                   OI.analyze_cystic_vs_solid(m64["t1"], m64["t2"], m64["flair"], seg, om, vd))
        tree_close(U.get_normal_brain_stats(mri["flair"], lv), OI.get_normal_brain_stats(m64["flair"], seg))
    empty = np.zeros((16, 16, 16), np.uint8)
    mri = SY.mri_volumes(0, empty)
    lv = U.LabelVolume(empty)
    gm, om = U.get_tumor_masks(lv), OP.get_tumor_masks(empty.astype(np.float64))
    assert S4.analyze_border_regularity(gm["wt"], vd) == OI.analyze_border_regularity(om["wt"], vd)
    assert S4.analyze_margin_definition(mri["t1ce"], lv, gm, vd) == OI.analyze_margin_definition(
        mri["t1ce"].astype(np.float64), empty, om, vd)
    assert S4.analyze_cystic_vs_solid(mri["t1"], mri["t2"], mri["flair"], lv, gm, vd) == OI.analyze_cystic_vs_solid(
        *(mri[k].astype(np.float64) for k in ("t1", "t2", "flair")), empty, om, vd)
    tiny = empty.copy()
    tiny[8, 8, 8:10] = 3  # 2 voxels: fewer than 10 surface voxels
    lv = U.LabelVolume(tiny)
    assert S4.analyze_border_regularity(U.get_tumor_masks(lv)["wt"], vd) == OI.analyze_border_regularity(tiny > 0, vd)
