"""GPU parity of the conv stacks + sliding-window path against the CPU oracle (fp32 torch restatement).

Tolerance (BASELINE.json north_star): probabilities within 1e-2 absolute for the 16-bit path — asserted on EVERY
prediction, with or without mirror / overlap averaging (BatchNorm-folded stacks run in bf16, InstanceNorm / GroupNorm
stacks in fp16).  Raw logits of single forwards are additionally held to logit-relative bounds.  Label agreement is
asserted on voxels whose oracle probability is not within the tolerance of the decision threshold (random-init nets put
many voxels near 0.5, where a rounding-sized error legitimately flips the decision) and reported overall.
"""
import numpy as np
import pytest
import torch

from oracle import sliding_window as SW
from tests.helpers import build_dropin_unet, oracle_fns

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-2


@pytest.mark.parametrize("variant", ["bn", "in", "gn"])
def test_forward_logits_match_oracle(variant):
    net = build_dropin_unet(variant, base=16, num_pool=3, groups=4, seed=3)
    fwd, _, _ = oracle_fns(net)
    x = torch.randn(2, 4, 32, 32, 32, generator=torch.Generator().manual_seed(7))
    ref = fwd(x)
    got = net(x).cpu()
    assert got.shape == ref.shape
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    perr = (torch.sigmoid(got) - torch.sigmoid(ref)).abs().max().item()
    rms = (got - ref).pow(2).mean().sqrt().item()
    print(f"{variant}: logits max err {err:.4g} rms {rms:.4g} (scale {scale:.3g}), sigmoid max err {perr:.4g}")
    assert rms < 4e-3 * max(scale, 1.0)   # un-averaged bf16 noise: ~0.2 % rms of the logit range
    assert err < 3e-2 * max(scale, 1.0)   # ... and < 3 % max over 2x3x32^3 logits
    assert perr < PROB_TOL


def test_forward_batch_one_and_engine_reuse():
    net = build_dropin_unet("bn", base=16, num_pool=2, seed=5)
    fwd, _, _ = oracle_fns(net)
    g = torch.Generator().manual_seed(1)
    for _ in range(2):  # second call reuses the cached engine and must not see stale activations
        x = torch.randn(1, 4, 16, 32, 24, generator=g)
        assert (torch.sigmoid(net(x).cpu()) - torch.sigmoid(fwd(x))).abs().max().item() < PROB_TOL


def _check_predict(net, vol, patch, mirror_axes, do_mirroring, step, regions, nonlin_fn, use_gaussian=True,
                   tol=PROB_TOL, mixed_precision=True, ref=None):
    if ref is None:
        fwd, _, _ = oracle_fns(net)
        ref = SW.predict_3d_tiled(fwd, nonlin_fn, vol, net.num_classes, patch, do_mirroring, mirror_axes, step,
                                  use_gaussian, regions)
    seg_ref, probs_ref = ref
    seg, probs = net.predict_3D(vol, do_mirroring, mirror_axes, True, step, patch, regions, use_gaussian, "constant",
                                {"constant_values": 0}, False, False, mixed_precision)
    assert probs.dtype == np.float32 and probs.shape == probs_ref.shape and seg.shape == seg_ref.shape
    assert seg.dtype == seg_ref.dtype
    perr = np.abs(probs - probs_ref).max()
    agree = (seg == seg_ref).mean()
    if regions is not None:
        decisive = np.all(np.abs(probs_ref - 0.5) > tol, axis=0)
    else:
        top2 = np.sort(probs_ref, axis=0)[-2:]
        decisive = (top2[1] - top2[0]) > 2 * tol
    print(f"prob max err {perr:.4g}, label agreement {agree * 100:.3f}% "
          f"({decisive.mean() * 100:.1f}% decisive voxels)")
    assert perr < tol
    assert np.array_equal(seg[decisive], seg_ref[decisive])
    return perr, agree, ref


def test_predict_3d_regions_sigmoid_all_mirrors():
    net = build_dropin_unet("bn", base=16, num_pool=3, seed=11)
    vol = torch.randn(4, 40, 56, 48, generator=torch.Generator().manual_seed(2)).numpy()
    _check_predict(net, vol, (32, 32, 32), (0, 1, 2), True, 0.5, (1, 2, 3), torch.sigmoid)


def test_predict_3d_softmax_argmax_groupnorm():
    net = build_dropin_unet("gn", base=16, num_pool=2, groups=4, seed=12, nonlin="softmax")
    vol = torch.randn(4, 33, 40, 37, generator=torch.Generator().manual_seed(3)).numpy()
    _check_predict(net, vol, (32, 32, 32), (0, 1, 2), True, 0.5, None, lambda t: torch.softmax(t, 1))


def test_predict_3d_volume_smaller_than_patch_and_subset_mirrors():
    net = build_dropin_unet("in", base=16, num_pool=2, seed=13)
    vol = torch.randn(4, 20, 32, 27, generator=torch.Generator().manual_seed(4)).numpy()
    # 4 mirrors, one padded tile, InstanceNorm over tiny (4^3) deep levels
    _check_predict(net, vol, (32, 32, 32), (1, 2), True, 0.5, (1, 2, 3), torch.sigmoid)
    # 2 mirrors / no mirrors, step 0.25 / 1.0, no Gaussian
    _check_predict(net, vol, (16, 16, 16), (0,), True, 0.25, (1, 2, 3), torch.sigmoid)
    _check_predict(net, vol, (16, 16, 16), (0, 1, 2), False, 1.0, None, torch.sigmoid, use_gaussian=False)


def test_predict_3d_full_brats_geometry_tiny_net():
    """BASELINE size (4x155x240x240, patch 128^3, step 0.5 -> 18 tiles) with a net small enough for the CPU oracle."""
    net = build_dropin_unet("bn", base=16, num_pool=2, seed=14)
    vol = torch.randn(4, 155, 240, 240, generator=torch.Generator().manual_seed(0)).numpy()
    _check_predict(net, vol, (128, 128, 128), (0, 1, 2), False, 0.5, (1, 2, 3), torch.sigmoid)


def test_mirror_equivariance_full_size():
    """Size-independent property: with all 8 mirrors and a symmetric step grid, predicting the flipped volume gives
    the flipped prediction (up to bf16 noise).  Checked at the BASELINE volume size with the BraTS architecture."""
    net = build_dropin_unet("bn", base=32, num_pool=5, seed=1)
    vol = torch.randn(4, 155, 240, 240, generator=torch.Generator().manual_seed(0))
    _, p1 = net.predict_3D_device(vol, True, (0, 1, 2), 0.5, (128, 128, 128), (1, 2, 3), True)
    _, p2 = net.predict_3D_device(torch.flip(vol, (1, 2, 3)), True, (0, 1, 2), 0.5, (128, 128, 128), (1, 2, 3), True)
    err = (p1 - torch.flip(p2, (1, 2, 3))).abs().max().item()
    print("mirror equivariance max prob diff", err)
    assert err < PROB_TOL
    assert torch.isfinite(p1).all() and p1.min() >= 0 and p1.max() <= 1


@pytest.mark.parametrize("variant", ["bn_brats", "gn_large_brats"])
def test_forward_full_patch_brats_architectures(variant):
    """One 128^3 forward of each benchmark architecture (model 1: BN 31.2 M; model 2: GroupNorm 87.4 M) against the
    fp32 oracle: every conv-kernel variant the benchmark launches (brick, tile, stride-2, transposed) at its real size."""
    if variant == "bn_brats":
        net = build_dropin_unet("bn", base=32, num_pool=5, seed=1)
    else:
        net = build_dropin_unet("gn", base=32, num_pool=5, seed=2, groups=8, encoder_scale=2, max_num_features=512)
    fwd, _, _ = oracle_fns(net)
    x = torch.randn(1, 4, 128, 128, 128, generator=torch.Generator().manual_seed(11))
    ref = fwd(x)
    got = net(x).cpu()
    perr = (torch.sigmoid(got) - torch.sigmoid(ref)).abs().max().item()
    rms = (got - ref).pow(2).mean().sqrt().item()
    print(f"{variant}: logits rms err {rms:.4g} (scale {ref.abs().max().item():.3g}), sigmoid max err {perr:.4g}")
    assert perr < PROB_TOL


def test_case_pipeline_folds_and_two_model_ensemble():
    """BratsCasePipeline with two folds per model: device-side fold mean (bsg_finalize over K accumulators) + regions
    decision + label-round ensemble + BraTS remap against the oracle chain (reference :128, :144-156, :305)."""
    from brainseg_b200 import pipeline as PL
    from oracle import postproc as OP

    models = [[build_dropin_unet("bn", base=16, num_pool=2, seed=31), build_dropin_unet("bn", base=16, num_pool=2, seed=32)],
              [build_dropin_unet("gn", base=16, num_pool=2, groups=4, seed=33),
               build_dropin_unet("gn", base=16, num_pool=2, groups=4, seed=34)]]
    vol = torch.randn(4, 40, 48, 36, generator=torch.Generator().manual_seed(9)).numpy()
    patch = (32, 32, 32)
    pipe = PL.BratsCasePipeline(models, patch, 0.5, (0, 1, 2), True, True, (1, 2, 3), "brats2025", batch=8)
    out = pipe.run_case(vol, features=False)
    segs_ref, decisive = [], np.ones(vol.shape[1:], dtype=bool)
    for folds in models:
        probs = [SW.predict_3d_tiled(oracle_fns(n)[0], torch.sigmoid, vol, 3, patch, True, (0, 1, 2), 0.5, True, (1, 2, 3))[1]
                 for n in folds]
        mean = np.mean(probs, axis=0)
        seg = np.zeros(mean.shape[1:], dtype=np.uint8)
        for i, c in enumerate((1, 2, 3)):
            seg[mean[i] > 0.5] = c
        segs_ref.append(seg)
        decisive &= np.all(np.abs(mean - 0.5) > PROB_TOL, axis=0)
    for got, ref in zip(out["model_segmentations"], segs_ref):
        assert np.array_equal(got.cpu().numpy()[decisive], ref[decisive])
    final_ref = OP.convert_labels_to_brats2025(OP.ensemble_labels_round(segs_ref[0], segs_ref[1]).astype(np.float64))
    assert np.array_equal(out["segmentation"].cpu().numpy()[decisive], final_ref[decisive])
    print(f"decisive voxels {decisive.mean() * 100:.1f}%")


@pytest.mark.parametrize("patch,step,shape", [((160, 160, 160), 0.5, (4, 168, 176, 160)), ((128, 128, 128), 0.25, (4, 130, 140, 128))])
def test_predict_3d_sweep_geometries(patch, step, shape):
    """BASELINE configs[4]: the patch-size / overlap sweep geometries (160^3 patches: odd 5^3 bottleneck level, levels
    that are not multiples of the tile boxes; step 0.25: denser tile grid) on a small net against the oracle."""
    net = build_dropin_unet("in", base=16, num_pool=5, seed=17)
    vol = torch.randn(*shape, generator=torch.Generator().manual_seed(6)).numpy()
    _check_predict(net, vol, patch, (0, 1, 2), False, step, (1, 2, 3), torch.sigmoid)


def test_case_pipeline_kaist_probability_ensemble():
    """The original KAIST post-processing (archived/kaist_original_inference.py:29-33): mean of the two models'
    probabilities, regions decision, enhancing-tumour suppression below 200 voxels, BraTS-2021 label convention."""
    from brainseg_b200 import pipeline as PL
    from oracle import postproc as OP

    models = [build_dropin_unet("bn", base=16, num_pool=2, seed=41), build_dropin_unet("gn", base=16, num_pool=2, groups=4, seed=42)]
    vol = torch.randn(4, 36, 40, 44, generator=torch.Generator().manual_seed(10)).numpy()
    patch = (32, 32, 32)
    probs = [SW.predict_3d_tiled(oracle_fns(n)[0], torch.sigmoid, vol, 3, patch, True, (0, 1, 2), 0.5, True, (1, 2, 3))[1]
             for n in models]
    mean = np.mean(probs, axis=0)
    seg = np.zeros(mean.shape[1:], dtype=np.uint8)
    for i, c in enumerate((1, 2, 3)):
        seg[mean[i] > 0.5] = c
    decisive = np.all(np.abs(mean - 0.5) > PROB_TOL, axis=0)
    n_et = int((seg == 3).sum())
    for thr in (n_et // 2, 4 * n_et + 200):  # below / above the measured enhancing-tumour count
        pipe = PL.BratsCasePipeline(models, patch, 0.5, (0, 1, 2), True, True, (1, 2, 3), "brats2021", batch=8,
                                    ensemble="prob_mean", small_et_threshold=thr, small_et_replace=2)
        out = pipe.run_case(vol, features=False)
        ref = seg.copy()
        if n_et < thr:
            ref[ref == 3] = 2
        ref = OP.convert_labels_to_brats2021(ref.astype(np.float64))
        got = out["segmentation"].cpu().numpy()
        assert np.array_equal(got[decisive], ref[decisive])
        assert (4 in np.unique(got)) == (n_et >= thr) or abs(n_et - thr) < (~decisive).sum()


@pytest.mark.parametrize("variant,shape,min_size", [("bn", (37, 45, 50), None), ("gn", (20, 33, 26), (32, 32, 32)),
                                                    ("in", (40, 48, 56), (16, 16, 16))])
def test_predict_3d_without_sliding_window(variant, shape, min_size):
    """use_sliding_window=False (upstream _internal_predict_3D_3Dconv): the whole volume, padded to a multiple of
    input_shape_must_be_divisible_by (and to >= patch_size when one is given), in ONE mirrored forward."""
    net = build_dropin_unet(variant, base=16, num_pool=2, groups=4, seed=31)
    fwd, _, _ = oracle_fns(net)
    vol = torch.randn(4, *shape, generator=torch.Generator().manual_seed(6)).numpy()
    div = [int(d) for d in net.input_shape_must_be_divisible_by]
    seg_ref, probs_ref = SW.predict_3d_full(fwd, torch.sigmoid, vol, 3, min_size or (0, 0, 0), div, True, (0, 1, 2),
                                            (1, 2, 3))
    seg, probs = net.predict_3D(vol, True, (0, 1, 2), False, 0.5, min_size, (1, 2, 3), False, "constant",
                                {"constant_values": 0}, False, False, True)
    assert probs.shape == probs_ref.shape == (3,) + shape and seg.dtype == seg_ref.dtype
    perr = np.abs(probs - probs_ref).max()
    decisive = np.all(np.abs(probs_ref - 0.5) > PROB_TOL, axis=0)
    print(f"whole-volume forward: prob max err {perr:.4g}, {decisive.mean() * 100:.1f}% decisive voxels")
    assert perr < PROB_TOL
    assert np.array_equal(seg[decisive], seg_ref[decisive])


def test_config1_full_case_brats_architecture():
    """BASELINE configs[0] in full: the synthetic 4x155x240x240 case, ONE model of the BraTS-2021 shape (31.2 M
    parameters), no mirroring, patch 128^3, step 0.5 (18 tiles), Gaussian weighting, regions export — against the fp32
    oracle running the same 18 forwards on the host (about a minute of CPU time on the GPU box).  north_star's bars:
    probabilities within 1e-2 absolute, label volume agreement >= 99.9 %.  Then the same case with
    mixed_precision=False (fp32-equivalent engine) against the same oracle result at north_star's fp32 bar: 1e-4."""
    net = build_dropin_unet("bn", base=32, num_pool=5, seed=1)
    vol = torch.randn(4, 155, 240, 240, generator=torch.Generator().manual_seed(0)).numpy()
    perr, agree, ref = _check_predict(net, vol, (128, 128, 128), (0, 1, 2), False, 0.5, (1, 2, 3), torch.sigmoid)
    assert agree >= 0.999, f"label agreement {agree * 100:.4f}% < 99.9%"
    perr32, agree32, _ = _check_predict(net, vol, (128, 128, 128), (0, 1, 2), False, 0.5, (1, 2, 3), torch.sigmoid,
                                        tol=1e-4, mixed_precision=False, ref=ref)
    assert agree32 >= 0.9999, f"fp32 mode label agreement {agree32 * 100:.5f}%"


# ---------------------------------------------------------------------------------------------- fp32-equivalent mode
FP32_TOL = 1e-4  # BASELINE.json north_star: probabilities within 1e-4 of the reference's fp32 (CPU) arithmetic


@pytest.mark.parametrize("variant", ["bn", "in", "gn"])
def test_fp32_mode_forward_logits(variant):
    """engine dtype "fp32" (fp16x3 split operands, three MMAs per fp32 product): logits of a single forward against the
    fp32 oracle — every layer kind (first conv on the split input, stride 2, concat of two split tensors, transposed)."""
    net = build_dropin_unet(variant, base=16, num_pool=3, groups=4, seed=3)
    net.engine_dtype = "fp32"
    fwd, _, _ = oracle_fns(net)
    x = torch.randn(2, 4, 32, 32, 32, generator=torch.Generator().manual_seed(7))
    ref = fwd(x)
    got = net(x).cpu()
    scale = max(ref.abs().max().item(), 1.0)
    err = (got - ref).abs().max().item()
    perr = (torch.sigmoid(got) - torch.sigmoid(ref)).abs().max().item()
    print(f"fp32 mode {variant}: logits max err {err:.3g} (scale {scale:.3g}), sigmoid max err {perr:.3g}")
    assert err < 1e-4 * scale
    assert perr < FP32_TOL


def test_fp32_mode_predict_3d_mixed_precision_false():
    """predict_3D(mixed_precision=False) — upstream's no-autocast path — through the sliding window with all mirrors:
    probabilities within 1e-4 of the fp32 oracle, labels equal wherever the oracle is not within 1e-4 of the threshold."""
    net = build_dropin_unet("gn", base=16, num_pool=3, groups=4, seed=21)
    fwd, _, _ = oracle_fns(net)
    vol = torch.randn(4, 40, 56, 48, generator=torch.Generator().manual_seed(2)).numpy()
    seg_ref, probs_ref = SW.predict_3d_tiled(fwd, torch.sigmoid, vol, net.num_classes, (32, 32, 32), True, (0, 1, 2), 0.5,
                                             True, (1, 2, 3))
    seg, probs = net.predict_3D(vol, True, (0, 1, 2), True, 0.5, (32, 32, 32), (1, 2, 3), True, "constant",
                                {"constant_values": 0}, False, False, mixed_precision=False)
    perr = np.abs(probs - probs_ref).max()
    decisive = np.all(np.abs(probs_ref - 0.5) > FP32_TOL, axis=0)
    agree = (seg == seg_ref).mean()
    print(f"fp32 mode predict_3D: prob max err {perr:.3g}, label agreement {agree * 100:.4f}% "
          f"({decisive.mean() * 100:.2f}% decisive)")
    assert perr < FP32_TOL
    assert np.array_equal(seg[decisive], seg_ref[decisive])
    assert agree > 0.9999


def test_fp32_mode_brats_architecture_full_patch():
    """One 128^3 forward of model 1's architecture (5 pools, 320 features) in fp32-equivalent mode."""
    net = build_dropin_unet("bn", base=32, num_pool=5, seed=1)
    net.engine_dtype = "fp32"
    fwd, _, _ = oracle_fns(net)
    x = torch.randn(1, 4, 128, 128, 128, generator=torch.Generator().manual_seed(11))
    ref = fwd(x)
    got = net(x).cpu()
    perr = (torch.sigmoid(got) - torch.sigmoid(ref)).abs().max().item()
    print(f"fp32 mode bn_brats: sigmoid max err {perr:.3g}")
    assert perr < FP32_TOL
