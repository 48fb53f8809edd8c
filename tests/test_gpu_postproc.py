"""GPU parity of the voxel post-processing kernels (through the C ABI) against the CPU oracle — bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import postproc as OP
from oracle import synthetic as SY

pytestmark = pytest.mark.gpu


def tree_equal(a, b, path="", float_rtol=0.0):
    if isinstance(a, dict):
        assert set(a.keys()) == set(b.keys()), f"{path}: {set(a.keys()) ^ set(b.keys())}"
        for k in a:
            tree_equal(a[k], b[k], f"{path}/{k}", float_rtol)
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            tree_equal(x, y, f"{path}[{i}]", float_rtol)
    elif isinstance(a, (float, np.floating)):
        if float_rtol == 0.0:
            assert float(a) == float(b), f"{path}: {a!r} != {b!r}"
        else:
            assert abs(float(a) - float(b)) <= float_rtol * max(abs(float(a)), abs(float(b)), 1e-300), f"{path}: {a} {b}"
    else:
        assert a == b, f"{path}: {a!r} != {b!r}"


@pytest.fixture(scope="module")
def mods():
    from brainseg_b200 import convert_labels_to_brats as CL
    from brainseg_b200 import evaluate_segmentation as EV
    from brainseg_b200 import voxelops as V
    from brainseg_b200.feature_extraction import step3_multiplicity as S3
    from brainseg_b200.feature_extraction import step4_morphology as S4
    from brainseg_b200.feature_extraction import utils as U
    return dict(CL=CL, EV=EV, V=V, S3=S3, S4=S4, U=U)


SHAPES = [(48, 40, 36), (33, 17, 70), (240, 240, 155)]


def _pair(seed, shape):
    pred, gt = SY.label_pair(seed, shape)
    rng = np.random.default_rng(seed + 7)
    for _ in range(20):  # isolated fragments + ties
        p = tuple(rng.integers(0, s) for s in shape)
        pred[p] = rng.integers(1, 4)
    return pred, gt


@pytest.mark.parametrize("shape", SHAPES)
def test_remap_and_ensemble(mods, shape):
    pred, gt = _pair(1, shape)
    CL, V = mods["CL"], mods["V"]
    assert np.array_equal(CL.convert_labels_to_brats2025(pred.astype(np.float64)),
                          OP.convert_labels_to_brats2025(pred.astype(np.float64)))
    assert np.array_equal(CL.convert_labels_to_brats2021(pred), OP.convert_labels_to_brats2021(pred))
    weird = pred.copy()
    weird[0, 0, :5] = [4, 5, 200, 255, 3]
    assert np.array_equal(CL.convert_labels_to_brats2025(weird), OP.convert_labels_to_brats2025(weird))
    a, b = V.as_label_volume(pred), V.as_label_volume(gt)
    assert np.array_equal(V.ensemble_round(a, b).cpu().numpy(), OP.ensemble_labels_round(pred, gt))
    fused = V.ensemble_round(a, b, post_lut=CL.LUT_BRATS2021).cpu().numpy()
    assert np.array_equal(fused, OP.convert_labels_to_brats2021(OP.ensemble_labels_round(pred, gt)))


@pytest.mark.parametrize("shape", SHAPES + [(5, 7, 3)])
def test_fused_ensemble_remap_histogram(mods, shape):
    """bsg_label_pair_round_hist_u8 == ensemble + remap followed by the joint histogram against the ground truth, and the
    Dice metrics built from its bins equal the oracle's (evaluate_segmentation.py:84-162), bit for bit."""
    CL, V, EV = mods["CL"], mods["V"], mods["EV"]
    pred, gt = _pair(3, shape)
    other = np.roll(pred, 2, axis=1).copy()
    a, b, g = V.as_label_volume(pred), V.as_label_volume(other), V.as_label_volume(gt)
    out, buf = V.ensemble_remap_hist(a, b, g, post_lut=CL.LUT_BRATS2025)
    ref = OP.convert_labels_to_brats2025(OP.ensemble_labels_round(pred, other).astype(np.float64))
    assert np.array_equal(out.cpu().numpy(), ref)
    hist = V.joint_hist_from_buffer(buf)
    assert np.array_equal(hist, V.joint_hist(out, g))
    assert hist.sum() == pred.size
    ev, ev_ref = EV.evaluate_arrays(out, g, _hist=hist), OP.evaluate_arrays(ref, gt)
    assert float(ev["mean_dice"]) == float(ev_ref["mean_dice"])
    assert sorted(ev["labels"]) == sorted(ev_ref["labels"])


def test_ensemble_all_byte_pairs(mods):
    V = mods["V"]
    a = np.repeat(np.arange(256, dtype=np.uint8), 256).reshape(16, 64, 64)
    b = np.tile(np.arange(256, dtype=np.uint8), 256).reshape(16, 64, 64)
    got = V.ensemble_round(V.as_label_volume(a), V.as_label_volume(b)).cpu().numpy()
    assert np.array_equal(got, OP.ensemble_labels_round(a, b))


def test_round_to_u8(mods):
    V = mods["V"]
    x = np.array([0.0, 0.49, 0.5, 1.5, 2.5, 2.51, 3.0000001, 0.9999999], dtype=np.float64).reshape(2, 2, 2)
    assert np.array_equal(V.as_label_volume(x).cpu().numpy(), np.round(x).astype(np.uint8))
    assert np.array_equal(V.as_label_volume(x.astype(np.float32)).cpu().numpy(), np.round(x.astype(np.float32)).astype(np.uint8))


@pytest.mark.parametrize("shape", SHAPES)
def test_dice_metrics_bit_exact(mods, shape):
    pred, gt = _pair(2, shape)
    EV = mods["EV"]
    predf, gtf = pred.astype(np.float64), gt.astype(np.float64)
    ref = OP.evaluate_arrays(predf, gtf)
    got = EV.evaluate_arrays(predf, gtf)
    tree_equal({int(k): v for k, v in ref["labels"].items()}, got["labels"], "labels")
    tree_equal(ref["wt"], got["wt"], "wt")
    tree_equal(ref["tc"], got["tc"], "tc")
    assert float(ref["mean_dice"]) == float(got["mean_dice"])
    for lab in (1, 2, 3):
        tree_equal(OP.calculate_metrics(predf, gtf, lab), EV.calculate_metrics(predf, gtf, lab), f"label{lab}")
    assert EV.evaluate_arrays(np.zeros((2, 2, 2)), np.zeros((2, 2, 3))) is None


@pytest.mark.parametrize("shape", SHAPES)
def test_ccl_labels_scipy_order(mods, shape):
    pred, _ = _pair(3, shape)
    labels, n = mods["S3"].label_components(pred)
    ref, nref = OP.label_components(pred > 0)
    assert n == nref
    assert np.array_equal(labels.cpu().numpy(), ref)
    labels3, n3 = mods["S3"].label_components(pred, mods["V"].bits_of(3))
    ref3, nref3 = OP.label_components(pred == 3)
    assert n3 == nref3 and np.array_equal(labels3.cpu().numpy(), ref3)


def test_ccl_adversarial_shapes(mods):
    S3 = mods["S3"]
    rng = np.random.default_rng(0)
    cases = []
    cases.append((rng.random((9, 21, 130)) < 0.5).astype(np.uint8))     # dense random noise: long equivalence chains
    cases.append((rng.random((40, 40, 40)) < 0.2).astype(np.uint8))
    cases.append(np.ones((5, 9, 66), dtype=np.uint8))                   # one component spanning every tile face
    z = np.zeros((12, 20, 140), dtype=np.uint8)
    z[::2, ::2, ::2] = 1                                                # isolated voxels only
    cases.append(z)
    s = np.zeros((8, 16, 128), dtype=np.uint8)                          # a serpentine path crossing tiles diagonally
    for i in range(8):
        s[i, 2 * i, 16 * i:16 * i + 17 if i < 7 else 128] = 1
        s[i, 2 * i + 1, min(16 * i + 17, 127)] = 1
    cases.append(s)
    cases.append(np.zeros((4, 4, 4), dtype=np.uint8))
    for vol in cases:
        labels, n = S3.label_components(vol)
        ref, nref = OP.label_components(vol > 0)
        assert n == nref
        assert np.array_equal(labels.cpu().numpy(), ref)


@pytest.mark.parametrize("shape", SHAPES[:2] + [(240, 240, 155)])
@pytest.mark.parametrize("vd", [(1.0, 1.0, 1.0), (0.9, 1.1, 1.25)])
def test_component_statistics(mods, shape, vd):
    if shape == (240, 240, 155) and vd != (1.0, 1.0, 1.0):
        pytest.skip("full-size oracle loop is slow; one voxel size is enough")
    pred, _ = _pair(4, shape)
    seg = pred.astype(np.int32)
    tree_equal(OP.detect_connected_components(seg, vd), mods["S3"].detect_connected_components(seg, vd), "components")
    tree_equal(OP.analyze_enhancing_components(seg, vd), mods["S3"].analyze_enhancing_components(seg, vd), "enhancing")


def test_more_components_than_the_statistics_table(mods):
    # 4800 isolated enhancing voxels + noise: more components than ccl26's default 4096-row statistics table, so the
    # first pass only counts and the wrapper re-runs with room for every component (this used to dead-lock)
    rng = np.random.default_rng(7)
    seg = np.zeros((24, 40, 40), dtype=np.int32)
    seg[::2, ::2, ::2] = 3
    seg[1::2, 1::2, :] = (rng.random((12, 20, 40)) < 0.05) * 2
    vd = (1.0, 1.0, 1.0)
    ref = OP.analyze_enhancing_components(seg, vd)
    assert ref["num_enhancing_foci"] == 4800
    tree_equal(ref, mods["S3"].analyze_enhancing_components(seg, vd), "enhancing")
    tree_equal(OP.detect_connected_components(seg, vd), mods["S3"].detect_connected_components(seg, vd), "components")
    _, n, st = mods["V"].ccl26(mods["V"].as_label_volume(seg), stats_cap=16)
    _, nref = OP.label_components(seg > 0)
    assert n == nref and len(st) == n and int(st["count"].sum()) == int((seg > 0).sum())


@pytest.mark.parametrize("shape", SHAPES)
def test_masks_volumes_morphology(mods, shape):
    pred, _ = _pair(5, shape)
    flat = pred.reshape(-1)
    idx = np.flatnonzero(flat == 3)[::5]
    flat[idx] = 4  # some BraTS-2021 style ET voxels
    U, S4 = mods["U"], mods["S4"]
    predf = pred.astype(np.float64)
    ref_masks = OP.get_tumor_masks(predf)
    got_masks = U.get_tumor_masks(predf)
    vd = (0.9, 1.1, 1.25)
    for k in ref_masks:
        assert int(ref_masks[k].sum()) == got_masks[k].sum(), k
        assert OP.calculate_volume(ref_masks[k], 0.001) == U.calculate_volume(got_masks[k], 0.001)
        if k != "background":
            tree_equal(OP.get_centroid(ref_masks[k]), U.get_centroid(got_masks[k]), f"centroid/{k}")
            tree_equal(OP.get_bounding_box(ref_masks[k]), U.get_bounding_box(got_masks[k]), f"bbox/{k}")
            assert OP.surface_voxel_count(ref_masks[k]) == int(got_masks[k].stats["surface"]), k
            assert np.array_equal(np.asarray(got_masks[k]), ref_masks[k])
    assert OP.calculate_surface_area(ref_masks["wt"], vd) == S4.calculate_surface_area(got_masks["wt"], vd)
    # floating-point descriptors: covariance assembled from exact integer moments vs np.cov of centred points
    tree_equal(OP.calculate_shape_descriptors(predf, ref_masks, vd), S4.calculate_shape_descriptors(predf, got_masks, vd),
               "shape", float_rtol=1e-9)
    tree_equal(OP.analyze_necrosis_pattern(predf, ref_masks, np.array(vd)),
               S4.analyze_necrosis_pattern(predf, got_masks, np.array(vd)), "necrosis", float_rtol=1e-12)


def test_empty_volume(mods):
    empty = np.zeros((8, 8, 8))
    vd = (1.0, 1.0, 1.0)
    U, S3, S4 = mods["U"], mods["S3"], mods["S4"]
    tree_equal(OP.detect_connected_components(empty.astype(np.int32), vd), S3.detect_connected_components(empty, vd))
    tree_equal(OP.analyze_enhancing_components(empty.astype(np.int32), vd), S3.analyze_enhancing_components(empty, vd))
    tree_equal(OP.calculate_shape_descriptors(empty, OP.get_tumor_masks(empty), vd),
               S4.calculate_shape_descriptors(empty, U.get_tumor_masks(empty), vd))
    tree_equal(OP.analyze_necrosis_pattern(empty, OP.get_tumor_masks(empty), np.array(vd)),
               S4.analyze_necrosis_pattern(empty, U.get_tumor_masks(empty), np.array(vd)))
    assert U.get_centroid(U.get_tumor_masks(empty)["wt"]) is None
    assert U.get_bounding_box(U.get_tumor_masks(empty)["wt"]) is None


def test_golden_fixture_through_gpu(mods, golden_dir):
    """The committed reference outputs (tests/golden/postproc.*) reproduced by the CUDA path directly."""
    vols = np.load(os.path.join(golden_dir, "postproc.npz"))
    with open(os.path.join(golden_dir, "postproc.json")) as f:
        ref = json.load(f)
    for seed in (0, 1, 2):
        pred, gt = vols[f"pred{seed}"], vols[f"gt{seed}"]
        r = ref[str(seed)]
        assert np.array_equal(mods["CL"].convert_labels_to_brats2025(pred), vols[f"remap2025_{seed}"])
        assert np.array_equal(mods["CL"].convert_labels_to_brats2021(pred), vols[f"remap2021_{seed}"])
        labels, n = mods["S3"].label_components(pred)
        assert n == r["cc_count"] and np.array_equal(labels.cpu().numpy(), vols[f"cc_labels_{seed}"])
        got = mods["EV"].evaluate_arrays(pred, gt)
        for lab, m in r["metrics"].items():
            for key, val in m.items():
                assert float(got["labels"][int(lab)][key]) == float(val), (seed, lab, key)
        comps = mods["S3"].detect_connected_components(pred, (1.0, 1.0, 1.0))
        tree_equal(json.loads(json.dumps(comps)), r["components_iso"], f"golden/components/{seed}")
