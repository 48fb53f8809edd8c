"""Pins the CPU oracle against outputs of the reference's own code (tests/golden, made by oracle/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import postproc as OP
from oracle import sliding_window as SW
from oracle import unet as OU


def _load_json(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return json.load(f)


def approx_equal_tree(a, b, path=""):
    """Exact for ints/strings/bools, exact-by-value for floats (fixtures hold the reference's own float results)."""
    if isinstance(a, dict):
        assert set(map(str, a.keys())) == set(map(str, b.keys())), path
        bb = {str(k): v for k, v in b.items()}
        for k, v in a.items():
            approx_equal_tree(v, bb[str(k)], f"{path}/{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            approx_equal_tree(x, y, f"{path}[{i}]")
    elif isinstance(a, (float, np.floating)) or isinstance(b, (float, np.floating)):
        assert float(a) == float(b) or (np.isnan(float(a)) and np.isnan(float(b))), f"{path}: {a} != {b}"
    else:
        if isinstance(a, np.generic):
            a = a.item()
        if isinstance(b, np.generic):
            b = b.item()
        assert a == b, f"{path}: {a} != {b}"


@pytest.mark.parametrize("variant", ["bn", "in", "gn"])
def test_unet_forward_matches_reference_module(golden_dir, variant):
    fx = torch.load(os.path.join(golden_dir, f"unet_{variant}.pt"), weights_only=False)
    y = OU.forward(fx["state_dict"], fx["arch"], fx["x"])
    # same torch ops in the same order as the reference module: bit-identical on one machine, tiny slack across CPUs
    assert torch.allclose(y, fx["logits"], rtol=0, atol=1e-5), (y - fx["logits"]).abs().max()


def test_unet_flops_and_keys(golden_dir):
    keys = _load_json(golden_dir, "unet_keys.json")
    assert keys["model1_bn"]["params"] == 31_199_360 or abs(keys["model1_bn"]["params"] / 1e6 - 31.20) < 0.01
    assert abs(keys["model1_bn"]["gflops_128"] - 965.5) < 1.0
    assert abs(keys["model2_gn_large"]["gflops_128"] - 3342.2) < 1.0
    k = keys["model2_gn_large"]["keys"]
    assert len(k) == 98
    assert k["conv_blocks_context.5.1.blocks.0.conv.weight"] == [512, 512, 3, 3, 3]
    assert k["tu.4.weight"] == [64, 64, 2, 2, 2] and k["seg_outputs.4.weight"] == [3, 32, 1, 1, 1]


def test_sliding_window_known_answers(golden_dir):
    facts = _load_json(golden_dir, "sliding_window.json")
    for name, expect in facts["survey_expected_steps"].items():
        assert facts["steps"][name] == expect, name
    assert SW.compute_steps_for_sliding_window((128,) * 3, (155, 240, 240), 0.5) == [[0, 27], [0, 56, 112], [0, 56, 112]]
    assert len(np.prod([len(s) for s in facts["steps"]["137x171x140_p128_s0.5"]], keepdims=True)) == 1
    assert int(np.prod([len(s) for s in facts["steps"]["137x171x140_p128_s0.5"]])) == 8
    g = SW.get_gaussian((128, 128, 128))
    assert g.dtype == np.float32 and g.max() == 1.0 and g[64, 64, 64] == 1.0
    assert g.min() == g[0, 0, 0]
    assert abs(float(g.min()) / facts["gaussian_128"]["survey_min"] - 1) < 1e-5
    assert facts["ensemble_lut"] == facts["survey_ensemble_lut"]
    lut = [[int(OP.ensemble_labels_round(np.array([a]), np.array([b]))[0]) for b in range(4)] for a in range(4)]
    assert lut == facts["survey_ensemble_lut"]


def test_pad_nd_image_and_decide():
    x = np.arange(2 * 3 * 4 * 5, dtype=np.float32).reshape(2, 3, 4, 5)
    p, sl = SW.pad_nd_image(x, (8, 4, 7))
    assert p.shape == (2, 8, 4, 7) and np.array_equal(p[sl], x)
    assert sl[1] == slice(2, 5) and sl[3] == slice(1, 6)
    probs = np.array([[[[0.6, 0.6, 0.2]]], [[[0.7, 0.1, 0.2]]], [[[0.9, 0.1, 0.1]]]], dtype=np.float32)
    assert SW.decide(probs, (1, 2, 3)).ravel().tolist() == [3.0, 1.0, 0.0]  # later regions overwrite
    assert SW.decide(probs, None).ravel().tolist() == [2, 0, 0]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_postproc_matches_reference_outputs(golden_dir, seed):
    vols = np.load(os.path.join(golden_dir, "postproc.npz"))
    ref = _load_json(golden_dir, "postproc.json")[str(seed)]
    pred, gt = vols[f"pred{seed}"], vols[f"gt{seed}"]
    predf, gtf = pred.astype(np.float64), gt.astype(np.float64)
    assert np.array_equal(OP.convert_labels_to_brats2025(predf), vols[f"remap2025_{seed}"])
    assert np.array_equal(OP.convert_labels_to_brats2021(predf), vols[f"remap2021_{seed}"])
    assert np.array_equal(OP.ensemble_labels_round(pred, gt), vols[f"ensemble_{seed}"])
    ev = OP.evaluate_arrays(predf, gtf)
    approx_equal_tree({str(int(k)): v for k, v in ev["labels"].items()}, ref["metrics"], "metrics")
    approx_equal_tree(ev["wt"], ref["wt"], "wt")
    approx_equal_tree(ev["tc"], ref["tc"], "tc")
    seg = np.round(predf).astype(np.int32)
    for tag, vd in (("iso", (1.0, 1.0, 1.0)), ("aniso", (0.9, 1.1, 1.25))):
        approx_equal_tree(OP.detect_connected_components(seg, vd), ref[f"components_{tag}"], f"components_{tag}")
        approx_equal_tree(OP.analyze_enhancing_components(seg, vd), ref[f"enhancing_{tag}"], f"enhancing_{tag}")
        masks = OP.get_tumor_masks(predf)
        approx_equal_tree(OP.calculate_shape_descriptors(predf, masks, vd), ref[f"shape_{tag}"], f"shape_{tag}")
        approx_equal_tree(OP.analyze_necrosis_pattern(predf, masks, np.array(vd)), ref[f"necrosis_{tag}"],
                          f"necrosis_{tag}")
    masks = OP.get_tumor_masks(predf)
    assert {k: int(v.sum()) for k, v in masks.items()} == ref["mask_counts"]
    approx_equal_tree({k: OP.get_centroid(v) for k, v in masks.items()}, ref["centroids"], "centroids")
    approx_equal_tree({k: OP.get_bounding_box(v) for k, v in masks.items()}, ref["bboxes"], "bboxes")
    assert OP.surface_voxel_count(masks["wt"]) == ref["surface_count_wt"]
    lab, n = OP.label_components(seg > 0)
    assert n == ref["cc_count"] and np.array_equal(lab, vols[f"cc_labels_{seed}"])


def test_postproc_empty_volume(golden_dir):
    ref = _load_json(golden_dir, "postproc.json")["empty"]
    empty = np.zeros((8, 8, 8))
    vd = (1.0, 1.0, 1.0)
    approx_equal_tree(OP.detect_connected_components(empty.astype(np.int32), vd), ref["components"])
    approx_equal_tree(OP.analyze_enhancing_components(empty.astype(np.int32), vd), ref["enhancing"])
    approx_equal_tree(OP.calculate_shape_descriptors(empty, OP.get_tumor_masks(empty), vd), ref["shape"])
    approx_equal_tree(OP.analyze_necrosis_pattern(empty, OP.get_tumor_masks(empty), np.array(vd)), ref["necrosis"])
    assert OP.evaluate_arrays(np.zeros((2, 2, 2)), np.zeros((2, 2, 3))) is None  # evaluate_segmentation.py:78-81


def test_scipy_label_order_and_erosion_border():
    # labels are numbered in C-order raster order of each component's first voxel
    m = np.zeros((4, 4, 4), dtype=bool)
    m[3, 0, 0] = m[0, 3, 3] = m[0, 0, 1] = True
    lab, n = OP.label_components(m)
    assert n == 3 and lab[0, 0, 1] == 1 and lab[0, 3, 3] == 2 and lab[3, 0, 0] == 3
    # a 5^3 all-ones block erodes to 27 voxels (border_value=0)
    assert OP.surface_voxel_count(np.ones((5, 5, 5), dtype=bool)) == 125 - 27


@pytest.mark.needs_reference
def test_oracle_unet_equals_live_reference_module():
    from oracle import ref_import as R
    net = R.build_reference_unet("gn", base=8, num_pool=3, groups=4, seed=3)
    x = torch.randn(1, 4, 16, 16, 16, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        y = net(x)
    y2 = OU.forward(net.state_dict(), OU.arch_from_module(net), x)
    assert torch.equal(y, y2)
