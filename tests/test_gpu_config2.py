"""Parity of the BENCHMARKED configuration (BASELINE configs[1]) against the CPU oracle, at size.

* model 2 (GroupNorm large, 87.4 M parameters — 78 % of a benchmark case's FLOPs) through the full 18-tile sliding
  window with north_star's label bar (>= 99.9 % agreement) against the oracle run live on the host;
* both benchmark models x 8 mirrors through BratsCasePipeline on a two-tile sub-volume against the oracle chain
  (per-model regions decision, label-round ensemble, BraTS remap);
* the whole configs[1] case (2 models x 18 tiles x 8 mirrors) against the oracle result recorded once by
  oracle/make_config2_golden.py (tests/golden/config2_oracle.npz; 288 fp32 forwards are ~25 minutes of host time).
Reference call sites: run_brats2021_inference_singlethread.py:97-106, :144-156, :263-312.
"""
import numpy as np
import pytest
import torch

from oracle import postproc as OP
from oracle import sliding_window as SW
from tests.helpers import build_dropin_unet, oracle_fns

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-2
PATCH = (128, 128, 128)


def _gn_large():
    return build_dropin_unet("gn", base=32, num_pool=5, seed=2, groups=8, encoder_scale=2, max_num_features=512)


def test_model2_gn_large_full_sliding_window():
    """The large GroupNorm model over the full BASELINE volume (18 tiles, no mirroring, Gaussian weighting, regions):
    probabilities within 1e-2, label volume agreement >= 99.9 % against the fp32 oracle."""
    net = _gn_large()
    vol = torch.randn(4, 155, 240, 240, generator=torch.Generator().manual_seed(0)).numpy()
    fwd, _, _ = oracle_fns(net)
    seg_ref, probs_ref = SW.predict_3d_tiled(fwd, torch.sigmoid, vol, 3, PATCH, False, (0, 1, 2), 0.5, True, (1, 2, 3))
    seg, probs = net.predict_3D(vol, False, (0, 1, 2), True, 0.5, PATCH, (1, 2, 3), True, "constant",
                                {"constant_values": 0}, False, False, True)
    perr = float(np.abs(probs - probs_ref).max())
    agree = float((seg == seg_ref).mean())
    decisive = np.all(np.abs(probs_ref - 0.5) > PROB_TOL, axis=0)
    print(f"GN-large full window: prob max err {perr:.4g}, label agreement {agree * 100:.4f}% "
          f"({decisive.mean() * 100:.1f}% decisive voxels)")
    assert perr < PROB_TOL
    assert np.array_equal(seg[decisive], seg_ref[decisive])
    assert agree >= 0.999, f"label agreement {agree * 100:.4f}% < 99.9%"


def test_two_model_eight_mirror_chain_two_tiles():
    """Both benchmark architectures, all 8 mirrors, two overlapping tiles, through BratsCasePipeline: per-model label
    volumes, the label-round ensemble and the BraTS-2025 remap against the oracle chain."""
    from brainseg_b200 import pipeline as PL

    models = [build_dropin_unet("bn", base=32, num_pool=5, seed=1), _gn_large()]
    vol = torch.randn(4, 128, 128, 160, generator=torch.Generator().manual_seed(5)).numpy()
    pipe = PL.BratsCasePipeline(models, PATCH, 0.5, (0, 1, 2), True, True, (1, 2, 3), "brats2025", batch=16)
    out = pipe.run_case(vol, features=False)
    segs_ref, decisive = [], np.ones(vol.shape[1:], dtype=bool)
    for m, net in enumerate(models):
        fwd, _, _ = oracle_fns(net)
        seg_ref, probs_ref = SW.predict_3d_tiled(fwd, torch.sigmoid, vol, 3, PATCH, True, (0, 1, 2), 0.5, True, (1, 2, 3))
        got = out["model_segmentations"][m].cpu().numpy()
        dec = np.all(np.abs(probs_ref - 0.5) > PROB_TOL, axis=0)
        agree = float((got == seg_ref).mean())
        print(f"model {m + 1}: label agreement {agree * 100:.4f}% ({dec.mean() * 100:.1f}% decisive voxels)")
        assert np.array_equal(got[dec], seg_ref[dec])
        assert agree >= 0.999
        segs_ref.append(seg_ref.astype(np.uint8))
        decisive &= dec
    final_ref = OP.convert_labels_to_brats2025(OP.ensemble_labels_round(segs_ref[0], segs_ref[1]).astype(np.float64))
    final = out["segmentation"].cpu().numpy()
    assert np.array_equal(final[decisive], final_ref[decisive])
    assert float((final == final_ref).mean()) >= 0.999


def test_config2_full_case_against_recorded_oracle():
    """BASELINE configs[1] in full — exactly the case bench.py times — against the oracle's recorded result."""
    from brainseg_b200 import pipeline as PL
    import synthetic_case as SY

    golden = SY.load_config2_oracle()
    if golden is None:
        pytest.skip("tests/golden/config2_oracle.npz not generated (oracle/make_config2_golden.py)")
    m1, m2 = SY.build_benchmark_models("large")
    vol = SY.case_volume(0, (4, 155, 240, 240))
    L = golden["lattice"]
    for m, net in enumerate((m1, m2), 1):
        seg, probs = net.predict_3D_device(vol, True, (0, 1, 2), 0.5, PATCH, (1, 2, 3), True)
        seg = seg.cpu().numpy()
        lat = probs[:, ::L, ::L, ::L].cpu().numpy()
        perr = float(np.abs(lat - golden[f"probs{m}"]).max())
        agree = float((seg == golden[f"seg{m}"]).mean())
        dec = golden[f"decisive{m}"]
        print(f"configs[1] model {m}: prob max err on the 1/{L ** 3} lattice {perr:.4g}, label agreement "
              f"{agree * 100:.4f}% ({dec.mean() * 100:.1f}% decisive voxels)")
        assert perr < PROB_TOL
        assert np.array_equal(seg[dec], golden[f"seg{m}"][dec])
        assert agree >= 0.999, f"model {m}: label agreement {agree * 100:.4f}% < 99.9%"
    # and the pipeline the benchmark runs (two stream lanes x 8, case-stream API), ensemble + remap included
    pipe = PL.BratsCasePipeline([m1, m2], PATCH, 0.5, (0, 1, 2), True, True, (1, 2, 3), "brats2025", batch=16)
    out = pipe.run_case(vol, features=False)
    final_ref = OP.convert_labels_to_brats2025(OP.ensemble_labels_round(golden["seg1"], golden["seg2"]).astype(np.float64))
    final = out["segmentation"].cpu().numpy()
    both = golden["decisive1"] & golden["decisive2"]
    agree = float((final == final_ref).mean())
    print(f"configs[1] ensemble + remap: label agreement {agree * 100:.4f}%")
    assert np.array_equal(final[both], final_ref[both])
    assert agree >= 0.999


def test_fp16_range_guard_reruns_in_bf16():
    """A net whose first conv weights are scaled so that raw (pre-norm) outputs exceed fp16's 65504: the conv epilogue
    raises the overflow flag, the pipeline re-plans in bf16 and the result still matches the oracle (InstanceNorm
    removes the scale; bf16 tolerance)."""
    from brainseg_b200 import pipeline as PL

    net = build_dropin_unet("in", base=16, num_pool=2, seed=51)
    with torch.no_grad():
        net.conv_blocks_context[0].blocks[0].conv.weight.mul_(3.0e5)
        net.conv_blocks_context[0].blocks[0].conv.bias.mul_(3.0e5)
    net.refresh_engines()
    vol = torch.randn(4, 32, 40, 36, generator=torch.Generator().manual_seed(8)).numpy()
    fwd, _, _ = oracle_fns(net)
    seg_ref, probs_ref = SW.predict_3d_tiled(fwd, torch.sigmoid, vol, 3, (32, 32, 32), True, (0, 1, 2), 0.5, True, (1, 2, 3))
    pipe = PL.BratsCasePipeline([net], (32, 32, 32), 0.5, (0, 1, 2), True, True, (1, 2, 3), "brats2025", batch=8)
    assert pipe._guarded()
    out = pipe.run_case(vol, features=False)
    assert pipe.fp16_overflows == 1 and net.engine_dtype == "bf16"
    got = out["model_segmentations"][0].cpu().numpy()
    decisive = np.all(np.abs(probs_ref - 0.5) > 5e-2, axis=0)  # bf16 InstanceNorm stacks: ~1.5e-2 probability noise
    assert decisive.mean() > 0.1
    assert np.array_equal(got[decisive], seg_ref[decisive])
    # a second case goes straight through the bf16 engines
    out2 = pipe.run_case(vol, features=False)
    assert pipe.fp16_overflows == 1
    assert torch.equal(out2["segmentation"], out["segmentation"])


def test_norm_statistics_on_tiny_levels_with_batch():
    """Levels with fewer than 32 voxels per batch item (2^3 bottleneck of a 32^3 patch with 4 poolings) and batch > 1:
    the rows of one epilogue warp span several batch items; every item must get its own InstanceNorm statistics."""
    net = build_dropin_unet("in", base=16, num_pool=4, seed=61)
    fwd, _, _ = oracle_fns(net)
    x = torch.randn(6, 4, 32, 32, 32, generator=torch.Generator().manual_seed(3))
    x[1] *= 3.0  # items with different statistics
    x[4] += 1.5
    ref = fwd(x)
    got = net(x).cpu()
    perr = (torch.sigmoid(got) - torch.sigmoid(ref)).abs().amax(dim=(1, 2, 3, 4))
    print("per-item sigmoid max err", [f"{v:.4g}" for v in perr.tolist()])
    assert float(perr.max()) < PROB_TOL


def test_load_state_dict_repacks_cached_engines_in_place():
    """load_state_dict / load_checkpoint_ram per fold: the cached engine (buffers, plans, tensor maps) is kept and only
    its packed weights are rewritten."""
    net = build_dropin_unet("gn", base=16, num_pool=2, groups=4, seed=71)
    other = build_dropin_unet("gn", base=16, num_pool=2, groups=4, seed=72)
    x = torch.randn(1, 4, 32, 32, 32, generator=torch.Generator().manual_seed(2))
    net(x)
    eng = net.engine_for((32, 32, 32), 1)
    net.load_state_dict(other.state_dict())
    assert net.engine_for((32, 32, 32), 1) is eng
    fwd, _, _ = oracle_fns(other)
    assert (torch.sigmoid(net(x).cpu()) - torch.sigmoid(fwd(x))).abs().max().item() < PROB_TOL
