"""CPU tests of the host-side logic: the C ABI loads and exports what include/*.h declares, the drop-in classes keep
the reference's checkpoint layout, the sliding-window geometry matches the oracle, and the product refuses to run
without its CUDA path."""
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    import __graft_entry__ as G
    G.build()
    from brainseg_b200 import _lib as L
    lib = L.lib()
    header = open(os.path.join(ROOT, "include", "brainseg_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(bsg_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.bsg_version() >= 100
    assert lib.bsg_ccl26_workspace_bytes(240, 240, 155) >= 240 * 240 * 155 * 4


def test_struct_layouts_match_header():
    from brainseg_b200 import voxelops as V
    assert V.COMP_DTYPE.itemsize == 88 and V.MOM_DTYPE.itemsize == 120
    assert V.MOM_DTYPE.fields["surface"][1] == 80 and V.COMP_DTYPE.fields["mn0"][1] == 56


def test_no_cpu_fallback():
    from brainseg_b200 import _lib as L
    from brainseg_b200 import voxelops as V
    from tests.helpers import build_dropin_unet
    if torch.cuda.is_available():
        pytest.skip("this check is for CPU-only hosts")
    with pytest.raises(L.BsgError):
        V.as_label_volume(np.zeros((4, 4, 4), dtype=np.uint8))
    net = build_dropin_unet("bn", base=16, num_pool=2)
    with pytest.raises(L.BsgError):
        net(torch.zeros(1, 4, 16, 16, 16))
    with pytest.raises(RuntimeError):
        net.conv_blocks_context[0].blocks[0](torch.zeros(1, 4, 8, 8, 8))  # blocks only hold parameters


def test_dropin_unet_checkpoint_layout(golden_dir):
    from tests.helpers import build_dropin_unet
    keys = json.load(open(os.path.join(golden_dir, "unet_keys.json")))
    m1 = build_dropin_unet("bn", base=32, num_pool=5)
    m2 = build_dropin_unet("gn", base=32, num_pool=5, groups=8, encoder_scale=2, max_num_features=512)
    for net, ref in ((m1, keys["model1_bn"]), (m2, keys["model2_gn_large"])):
        sd = net.state_dict()
        assert list(sd.keys()) == list(ref["keys"].keys())  # same names in the same order
        assert all(list(sd[k].shape) == ref["keys"][k] for k in sd)
        assert sum(p.numel() for p in net.parameters()) == ref["params"]
    assert [int(v) for v in m1.input_shape_must_be_divisible_by] == [32, 32, 32]
    assert m1.num_classes == 3 and m1.conv_op == torch.nn.Conv3d and m1.do_ds is False


@pytest.mark.needs_reference
def test_dropin_unet_loads_reference_state_dict():
    """A state_dict produced by the REFERENCE class loads into the drop-in (strict) and vice versa."""
    from oracle import ref_import as R
    from tests.helpers import build_dropin_unet
    ref = R.build_reference_unet("gn", base=16, num_pool=3, groups=4, seed=3)
    ours = build_dropin_unet("gn", base=16, num_pool=3, groups=4, seed=99)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)
    # same construction order + same seed => identical random initial weights as the reference class
    a = R.build_reference_unet("bn", base=16, num_pool=2, seed=5, randomize_norm=False).state_dict()
    b = build_dropin_unet("bn", base=16, num_pool=2, seed=5)
    torch.manual_seed(5)
    from brainseg_b200 import generic_UNet as G
    from torch import nn
    c = G.Generic_UNet(4, 16, 3, 2, 2, 2, nn.Conv3d, nn.BatchNorm3d, {"eps": 1e-5, "affine": True}, nn.Dropout3d,
                       {"p": 0, "inplace": True}, nn.LeakyReLU, {"negative_slope": 1e-2, "inplace": True}, True, False,
                       lambda x: x, G.InitWeights_He(1e-2), [[2, 2, 2]] * 2, [[3, 3, 3]] * 3, False, True,
                       True).state_dict()
    assert all(torch.equal(a[k], c[k]) for k in a)
    assert b is not None


def test_unsupported_configurations_raise():
    from brainseg_b200 import generic_UNet as G
    from torch import nn
    with pytest.raises(NotImplementedError):
        G.Generic_UNet(4, 32, 3, 5)  # 2-D default conv_op
    with pytest.raises(NotImplementedError):
        G.Generic_UNet(4, 32, 3, 5, conv_op=nn.Conv3d, norm_op=nn.BatchNorm3d, convolutional_pooling=False,
                       convolutional_upsampling=False)


def test_sliding_geometry_matches_oracle():
    from brainseg_b200 import sliding as S
    from oracle import sliding_window as SW
    for img, patch, step in [((155, 240, 240), (128,) * 3, 0.5), ((155, 240, 240), (128,) * 3, 0.25),
                             ((256,) * 3, (160,) * 3, 0.5), ((137, 171, 140), (128,) * 3, 0.5),
                             ((128, 128, 128), (128,) * 3, 0.5), ((40, 56, 48), (32,) * 3, 1.0)]:
        assert S.compute_steps_for_sliding_window(patch, img, step) == SW.compute_steps_for_sliding_window(patch, img, step)
    for patch in [(128, 128, 128), (32, 48, 64), (16, 16, 16)]:
        g = S.gaussian_importance_map(patch, torch.device("cpu")).numpy()
        assert np.array_equal(g, SW.get_gaussian(patch))  # bit-exact with the SciPy-based upstream construction
    assert S.mirror_codes_for((0, 1, 2)) == list(range(8))
    assert S.mirror_codes_for((0,)) == [0, 4] and S.mirror_codes_for((1, 2)) == [0, 1, 2, 3]
    assert S.mirror_codes_for((0, 1, 2), do_mirroring=False) == [0]
    # codes <-> upstream flip dims (App. A.6): bit0 = dim 4, bit1 = dim 3, bit2 = dim 2
    for m, flips in enumerate(SW.MIRROR_FLIPS):
        assert sorted(flips) == sorted(d for b, d in ((1, 4), (2, 3), (4, 2)) if m & b)


def test_work_item_sharding_partitions_the_case():
    from brainseg_b200 import sliding as S
    full = S.shard_work_items(18, range(8))
    assert len(full) == 144
    for world in (2, 3, 4, 8):
        shards = [S.shard_work_items(18, range(8), r, world) for r in range(world)]
        assert sorted(sum(shards, [])) == sorted(full)
        assert max(map(len, shards)) - min(map(len, shards)) <= 1


def test_balanced_batch_wastes_no_forwards():
    """Sharded mode: the forwards in flight are sized to the rank's share of the 144 (tile, mirror) items per model."""
    from brainseg_b200 import sliding as S
    for world in (1, 2, 3, 4, 6, 8):
        n = -(-144 // world)
        total = S.balanced_batch(n, lanes=2)
        b = total // 2
        assert 4 <= b <= 12 and total == 2 * b
        rounds = -(-n // total)
        rem = n - (rounds - 1) * total
        assert -(-rem // b) * b - rem == 0, (world, n, b)  # every BraTS share divides evenly
    assert S.balanced_batch(18, lanes=2) == 18        # 8 GPUs: two lanes of 9, one round
    assert S.balanced_batch(7, lanes=1, lo=4, hi=8) == 7


def test_weight_packing_layouts():
    from brainseg_b200 import packing as P
    w = torch.arange(2 * 3 * 27, dtype=torch.float32).reshape(2, 3, 3, 3, 3)
    p = P.pack_conv3_weight(w)
    assert p.shape == (27, 32, 16) and p.dtype == torch.bfloat16
    kd, kh, kw = 2, 0, 1
    tap = (kd * 3 + kw) * 3 + kh  # tap order (kd, kw, kh)
    assert p[tap, 1, 2].item() == w[1, 2, kd, kh, kw].item()
    assert p[:, 2:, :].abs().sum() == 0 and p[:, :, 3:].abs().sum() == 0
    wt = torch.arange(3 * 2 * 8, dtype=torch.float32).reshape(3, 2, 2, 2, 2)
    pt = P.pack_convT2_weight(wt)
    assert pt.shape == (1, 8 * 32, 16)
    par = 1 * 4 + 0 * 2 + 1  # parity index (kd, kh, kw)
    assert pt[0, par * 32 + 1, 2].item() == wt[2, 1, 1, 0, 1].item()
    x = torch.randn(2, 4, 3, 5, 6)
    assert torch.allclose(P.from_ndhwc(P.to_ndhwc_bf16(x), 4), x.to(torch.bfloat16).float())


def test_fp16x3_split_arithmetic():
    """Engine dtype "fp32": an activation tensor stored as [hi | hi | lo] contracted against weights stacked
    [w_hi | w_lo | w_hi] reproduces the fp32 product to ~2^-20 relative, with every operand fp16-exact (what the 16-bit
    tensor pipe multiplies exactly and accumulates in fp32)."""
    from brainseg_b200 import packing as P
    g = torch.Generator().manual_seed(0)
    x = torch.randn(64, 24, generator=g, dtype=torch.float64) * 3.0   # (voxels, logical channels), two source parts
    w = torch.randn(8, 24, generator=g, dtype=torch.float64) * 0.05  # (cout, logical channels)
    parts = [(0, 16), (48, 8)]                                        # part 0: 16 ch at 0, part 1: 8 ch at 48
    phys = torch.zeros(64, 80)
    start = 0
    for off, c in parts:
        hi, lo = P.split_f16x3(x[:, start:start + c].float())
        phys[:, off:off + c] = hi.float()
        phys[:, off + c:off + 2 * c] = hi.float()
        phys[:, off + 2 * c:off + 3 * c] = lo.float()
        start += c
    wk = P.split_k_weight(w.float(), parts, 80, 1)
    assert torch.equal(wk, wk.to(torch.float16).float()) and torch.equal(phys, phys.to(torch.float16).float())
    assert wk[:, 72:].abs().sum() == 0
    got = phys.double() @ wk.double().t()
    ref = x.float().double() @ w.float().double().t()
    plain = x.float().to(torch.float16).double() @ w.float().to(torch.float16).double().t()
    scale = (x.abs() @ w.abs().t()).max().item()
    assert (got - ref).abs().max().item() < 4e-6 * scale
    assert (plain - ref).abs().max().item() > 50 * (got - ref).abs().max().item()  # vs plain fp16 operands


def test_remap_luts_and_dice_formulas_on_host():
    from brainseg_b200 import convert_labels_to_brats as CL
    from brainseg_b200 import evaluate_segmentation as EV
    assert CL.LUT_BRATS2025[:5].tolist() == [0, 2, 1, 3, 0] and CL.LUT_BRATS2021[:5].tolist() == [0, 2, 1, 4, 0]
    assert CL.LUT_BRATS2025[5:].sum() == 0
    m = EV._metrics(np.float32(10), np.float32(2), np.float32(3), np.float32(100))
    assert m["dice"].dtype == np.float32 and m["dice"] == np.float32(20) / (np.float32(20) + np.float32(2) + np.float32(3) + 1e-8)


def test_nifti_round_trip(tmp_path):
    from brainseg_b200 import nifti_io as N

    d = np.random.default_rng(0).standard_normal((5, 6, 7)).astype(np.float32)
    like = N.new_header(d.shape, (1.0, 1.5, 2.0))
    for ext in (".nii", ".nii.gz"):
        p = str(tmp_path / ("a" + ext))
        N.save(p, d, like)
        im = N.load(p)
        assert np.array_equal(im.data, d) and im.zooms == (1.0, 1.5, 2.0)
        N.save(p, (d > 0).astype(np.uint8), im)  # label map written with the source image's geometry
        im2 = N.load(p)
        assert im2.data.dtype == np.uint8 and np.array_equal(im2.get_fdata(), (d > 0).astype(np.float64))
    with pytest.raises(ValueError):
        N.save(str(tmp_path / "b.nii"), np.zeros((2, 2, 2), np.uint8), like)  # geometry mismatch


def test_network_config_is_inferred_from_the_checkpoint():
    from brainseg_b200 import nnunet_compat as NC
    from tests.helpers import build_dropin_unet

    for kw, name in ((dict(variant="bn", base=32, num_pool=5), "nnUNetTrainerV2BraTSRegions_DA4_BN_BD"),
                     (dict(variant="gn", base=32, num_pool=5, groups=8, encoder_scale=2, max_num_features=512),
                      "nnUNetTrainerV2BraTSRegions_DA4_BN_BD_largeUnet_Groupnorm"),
                     (dict(variant="in", base=16, num_pool=2), "")):
        sd = build_dropin_unet(**kw).state_dict()
        sd2 = NC.build_network(NC.infer_network_config(sd, name)).state_dict()
        assert list(sd) == list(sd2) and all(sd[k].shape == sd2[k].shape for k in sd)


def test_run_full_pipeline_host_logic(tmp_path, capsys):
    """Step 1 (BraTS-2025 -> BraTS-2021 file names, .nii compression) and the exit-code / marker contract of the
    orchestrator (reference run_full_pipeline.py:89-144, :460-470, :714-732) — the parts that need no GPU."""
    import gzip

    from brainseg_b200 import run_full_pipeline as RP
    from brainseg_b200.feature_extraction import utils as U

    case = tmp_path / "BraTS-GLI-00003-000"
    case.mkdir()
    for suffix in ("t1n", "t1c", "t2w"):
        (case / f"BraTS-GLI-00003-000-{suffix}.nii").write_bytes(b"raw-nifti-bytes")
    (case / "BraTS-GLI-00003-000-t2f.nii.gz").write_bytes(gzip.compress(b"flair"))
    (case / "notes.txt").write_text("ignored")
    assert RP.rename_brats2025_files(case) == ("BraTS-GLI-00003-000", 4, 0)
    names = sorted(p.name for p in case.iterdir())
    assert names == ["BraTS-GLI-00003-000_flair.nii.gz", "BraTS-GLI-00003-000_t1.nii.gz",
                     "BraTS-GLI-00003-000_t1ce.nii.gz", "BraTS-GLI-00003-000_t2.nii.gz", "notes.txt"]
    assert gzip.decompress((case / "BraTS-GLI-00003-000_t1.nii.gz").read_bytes()) == b"raw-nifti-bytes"
    assert RP.rename_brats2025_files(case) == ("BraTS-GLI-00003-000", 0, 4)  # second run: already converted
    assert U.get_case_id(case) == "BraTS-GLI-00003-000"
    assert sorted(U.get_mri_paths(case)) == ["flair", "t1", "t1ce", "t2"]
    capsys.readouterr()

    # missing ground truth: STAGE:error + ERROR: line, exit code 1; the segmentation stage is never announced
    with pytest.raises(SystemExit) as stop:
        RP.main([str(case), "--results-root", str(tmp_path / "results")])
    out = capsys.readouterr().out
    assert stop.value.code == 1
    assert "STAGE:error" in out and "ERROR:Ground truth segmentation not found" in out and "STAGE:segmenting" not in out
    # a folder that does not exist: exit code 1 without any stage marker (the reference checks before its try block)
    with pytest.raises(SystemExit) as stop:
        RP.main([str(tmp_path / "nowhere")])
    out = capsys.readouterr().out
    assert stop.value.code == 1 and "STAGE:" not in out and "Case folder not found" in out


def test_step3_lesion_bookkeeping():
    """calculate_inter_lesion_distances / detect_satellite_lesions / classify_distribution_pattern on hand-made
    component lists (reference step3_multiplicity.py:155-205, :266-375)."""
    from brainseg_b200.feature_extraction import step3_multiplicity as S3

    def comp(i, x, vol=1.0, enh=True):
        return {"id": i, "centroid_mm": {"x": float(x), "y": 0.0, "z": 0.0}, "volume_cm3": vol, "has_enhancement": enh}

    one = [comp(1, 0)]
    assert S3.calculate_inter_lesion_distances(one, (1, 1, 1)) == {"distances": [], "min_distance_mm": None,
                                                                 "max_distance_mm": None, "mean_distance_mm": None}
    assert S3.detect_satellite_lesions(one, one[0], (1, 1, 1))["description"] == "Single lesion, no satellites"
    three = [comp(1, 0), comp(2, 15), comp(3, 50, enh=False)]
    d = S3.calculate_inter_lesion_distances(three, (1, 1, 1))
    assert [p["distance_mm"] for p in d["distances"]] == [15.0, 50.0, 35.0]
    assert [p["relationship"] for p in d["distances"]] == ["Satellite/adjacent", "Distant/separate", "Regional spread"]
    assert d["min_distance_mm"] == 15.0 and d["max_distance_mm"] == 50.0 and abs(d["mean_distance_mm"] - 100 / 3) < 1e-12
    sat = S3.detect_satellite_lesions(three, three[0], (1, 1, 1))
    assert sat["satellite_count"] == 1 and sat["satellites"][0]["component_id"] == 2 and sat["has_satellites"]
    pattern = S3.classify_distribution_pattern({"num_components": 3}, d, sat, {"num_enhancing_foci": 5})
    assert pattern["pattern"] == "Primary with satellites" and pattern["lesion_count"] == 3
    assert pattern["enhancement_note"].startswith("Multiple enhancing foci")
    far = [comp(1, 0), comp(2, 30), comp(3, 90)]
    dfar = S3.calculate_inter_lesion_distances(far, (1, 1, 1))
    nosat = S3.detect_satellite_lesions(far, far[0], (1, 1, 1))
    assert S3.classify_distribution_pattern({"num_components": 3}, dfar, nosat, {"num_enhancing_foci": 0})["pattern"] == \
        "Distant multifocal"
    near = [comp(1, 0), comp(2, 25), comp(3, 39)]
    dnear = S3.calculate_inter_lesion_distances(near, (1, 1, 1))
    assert S3.classify_distribution_pattern({"num_components": 3}, dnear, S3.detect_satellite_lesions(near, near[0], (1, 1, 1)),
                                            {"num_enhancing_foci": 3})["pattern"] == "Regional multifocal"
    assert S3.classify_distribution_pattern({"num_components": 5}, dnear, nosat, {"num_enhancing_foci": 1})["pattern"] == \
        "Diffuse/scattered"
    assert S3.classify_distribution_pattern({"num_components": 0}, None, None, None)["pattern"] == "No tumor"


@pytest.mark.parametrize("seed", ["0", "1"])
def test_step3_bookkeeping_matches_reference_drivers(seed):
    """The same functions fed with the reference's own component lists reproduce the reference's distance, satellite and
    distribution sections (tests/golden/step_drivers.json, produced by the reference's analyze_multiplicity)."""
    import json
    import os

    from brainseg_b200.feature_extraction import step3_multiplicity as S3
    from tests.conftest import GOLDEN

    def approx_equal_tree(a, b, path):
        """exact for structure / strings / ints; floats to float32 resolution: the reference's centroid_mm values are
        np.float32 under nibabel + NumPy >= 2, so its distances are float32 arithmetic"""
        if isinstance(a, dict):
            assert set(a) == set(b), path
            for k in a:
                approx_equal_tree(a[k], b[k], f"{path}/{k}")
        elif isinstance(a, list):
            assert len(a) == len(b), path
            for i, (x, y) in enumerate(zip(a, b)):
                approx_equal_tree(x, y, f"{path}[{i}]")
        elif isinstance(a, float) and isinstance(b, (int, float)) and not isinstance(b, bool):
            assert abs(a - b) <= 2e-6 * max(abs(a), abs(b)), f"{path}: {a} != {b}"
        else:
            assert a == b, f"{path}: {a!r} != {b!r}"

    with open(os.path.join(GOLDEN, "step_drivers.json")) as f:
        ref = json.load(f)[seed]["step3"]
    comps = ref["component_analysis"]
    dims = ref["voxel_info"]["dimensions_mm"]
    dist = S3.calculate_inter_lesion_distances(comps["components"], dims)
    approx_equal_tree(dist, ref["distance_analysis"], "distance_analysis")
    sat = S3.detect_satellite_lesions(comps["components"], comps["components"][0], dims)
    approx_equal_tree(sat, ref["satellite_analysis"], "satellite_analysis")
    approx_equal_tree(S3.classify_distribution_pattern(comps, dist, sat, ref["enhancing_analysis"]),
                      ref["distribution_pattern"], "distribution_pattern")


def test_bench_reference_arm_contract():
    """`bench.py --impl reference`: ONE JSON line with the contract's keys, produced on the host cores only; ranks other
    than 0 exit 0 without output (torchrun launches the arm on every rank)."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    quiet = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"], cwd=root,
                           env=dict(env, RANK="1", WORLD_SIZE="2"), capture_output=True, text=True, timeout=120)
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""
    run = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--skip-config0"], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stderr[-2000:]
    lines = [l for l in run.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "cases/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("cases/sec") and line["n_gpus"] == 1 and line["vs_baseline"] is None
    # a step is a MEASURED sample (one forward of each model = 1/144 of a case's forwards); the case rate is the labelled
    # extrapolation 1 / (144 x step + post-processing chain)
    assert line["value"] > 0 and line["extrapolated"] is True and line["steps_per_case"] == 144
    t_case = 144 * line["ms_per_step"] / 1e3 + line["post_chain_s"]
    assert abs(line["value"] * t_case - 1.0) < 1e-6 and abs(line["ms_per_case_extrapolated"] / 1e3 - t_case) < 1e-6
    assert line["e2e"] == {"value": line["value"], "unit": "cases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    base = line["cpu_baseline"]
    assert base["kind"] == "port" and base["cores"] >= 1 and base["value"] == line["value"] and "sample" in base
    assert "workload" in line["config"] and "model" not in line["config"]
