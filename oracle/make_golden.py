"""Generates tests/golden/* by running the REFERENCE's own code (imported from /root/reference, see ref_import.py)
on small seeded inputs.  Run in the build container:  python -m oracle.make_golden

Fixtures (all small, committed):
  unet_{bn,in,gn}.pt      state_dict + arch + input + logits of the reference Generic_UNet (tiny widths)
  unet_keys.json          state_dict key list / shapes, parameter count and forward GFLOPs of the two BraTS shapes
  postproc.npz/.json      label volumes and the reference functions' outputs on them (remap, Dice, step3, step4)
  sliding_window.json     known-answer facts for the restated nnU-Net v1 tiler
  voxelops.npz/.json      MRI-like volumes + the reference's intensity statistics / border / margin / cystic results
                          and SciPy's morphology, EDT and 6/18/26-connected labellings on them (python -m
                          oracle.make_golden voxelops regenerates only these)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import as R  # noqa: E402
from oracle import unet as U  # noqa: E402
from oracle import sliding_window as SW  # noqa: E402
from oracle import synthetic as SY  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def jsonable(o):
    if isinstance(o, dict):
        return {str(k): jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, np.generic):
        # keep float32 results exactly: repr via float() is exact for float32 -> float64
        return o.item()
    if isinstance(o, np.ndarray):
        return o.tolist()
    return o


def unet_fixtures():
    for variant in ("bn", "in", "gn"):
        net = R.build_reference_unet(variant, base=4, num_pool=3, groups=2, seed=11)
        g = torch.Generator().manual_seed(5)
        x = torch.randn(1, 4, 16, 16, 16, generator=g)
        with torch.no_grad():
            y = net(x)
        torch.save({"state_dict": {k: v.clone() for k, v in net.state_dict().items()},
                    "arch": U.arch_from_module(net), "x": x, "logits": y}, os.path.join(OUT, f"unet_{variant}.pt"))
    keys = {}
    for name, variant, kw in (("model1_bn", "bn", {}),
                              ("model2_gn_large", "gn", dict(encoder_scale=2, max_num_features=512, groups=8))):
        net = R.build_reference_unet(variant, **kw)
        sd = net.state_dict()
        arch = U.arch_from_module(net)
        keys[name] = {"keys": {k: list(v.shape) for k, v in sd.items()},
                      "params": int(sum(p.numel() for p in net.parameters())),
                      "gflops_128": U.conv_flops(sd, arch, (128, 128, 128)) / 1e9, "arch": jsonable(arch)}
    with open(os.path.join(OUT, "unet_keys.json"), "w") as f:
        json.dump(keys, f, indent=1)


def postproc_fixtures():
    ns = R.load_reference()
    shape = (48, 40, 36)
    vols, results = {}, {}
    voxel_dims = (1.0, 1.0, 1.0)
    aniso = (0.9, 1.1, 1.25)
    for seed in (0, 1, 2):
        pred, gt = SY.label_pair(seed, shape)
        # sprinkle a few isolated voxels / label-4 voxels so fragments, ties and "other" labels are exercised
        rng = np.random.default_rng(100 + seed)
        for _ in range(12):
            p = tuple(rng.integers(0, s) for s in shape)
            pred[p] = rng.integers(1, 4)
        if seed == 2:
            pred[pred == 3] = 4  # BraTS-2021 convention volume
        vols[f"pred{seed}"] = pred
        vols[f"gt{seed}"] = gt
        predf, gtf = pred.astype(np.float64), gt.astype(np.float64)
        r = {}
        r["remap2025_sha"] = ns.convert_labels.convert_labels_to_brats2025(predf).tolist() if False else None
        vols[f"remap2025_{seed}"] = ns.convert_labels.convert_labels_to_brats2025(predf)
        vols[f"remap2021_{seed}"] = ns.convert_labels.convert_labels_to_brats2021(predf)
        vols[f"ensemble_{seed}"] = np.round((predf + gtf) / 2.0).astype(np.uint8)  # run_brats...py:305 expression
        labs = sorted(set(np.unique(predf)) | set(np.unique(gtf)))
        r["metrics"] = {str(int(l)): ns.evaluate.calculate_metrics(predf, gtf, l) for l in labs if l != 0}
        r["wt"] = ns.evaluate.calculate_metrics_binary(np.isin(predf, [1, 2, 3]).astype(np.float32),
                                                       np.isin(gtf, [1, 2, 3]).astype(np.float32))
        r["tc"] = ns.evaluate.calculate_metrics_binary(np.isin(predf, [1, 3]).astype(np.float32),
                                                       np.isin(gtf, [1, 3]).astype(np.float32))
        seg = np.round(predf).astype(np.int32)  # step3_multiplicity.py:460
        for tag, vd in (("iso", voxel_dims), ("aniso", aniso)):
            r[f"components_{tag}"] = ns.step3.detect_connected_components(seg, vd)
            r[f"enhancing_{tag}"] = ns.step3.analyze_enhancing_components(seg, vd)
            masks = ns.utils.get_tumor_masks(predf)
            r[f"shape_{tag}"] = ns.step4.calculate_shape_descriptors(predf, masks, vd)
            r[f"necrosis_{tag}"] = ns.step4.analyze_necrosis_pattern(predf, masks, np.array(vd))
        masks = ns.utils.get_tumor_masks(predf)
        r["mask_counts"] = {k: int(v.sum()) for k, v in masks.items()}
        r["centroids"] = {k: ns.utils.get_centroid(v) for k, v in masks.items()}
        r["bboxes"] = {k: ns.utils.get_bounding_box(v) for k, v in masks.items()}
        r["surface_count_wt"] = int((masks["wt"] & ~__import__("scipy.ndimage").ndimage.binary_erosion(masks["wt"])).sum())
        lab, n = __import__("scipy.ndimage").ndimage.label(seg > 0, structure=np.ones((3, 3, 3)))
        vols[f"cc_labels_{seed}"] = lab.astype(np.int32)
        r["cc_count"] = int(n)
        r.pop("remap2025_sha")
        results[str(seed)] = jsonable(r)
    # empty-volume behaviour
    empty = np.zeros((8, 8, 8))
    results["empty"] = jsonable({
        "components": ns.step3.detect_connected_components(empty.astype(np.int32), voxel_dims),
        "enhancing": ns.step3.analyze_enhancing_components(empty.astype(np.int32), voxel_dims),
        "shape": ns.step4.calculate_shape_descriptors(empty, ns.utils.get_tumor_masks(empty), voxel_dims),
        "necrosis": ns.step4.analyze_necrosis_pattern(empty, ns.utils.get_tumor_masks(empty), np.array(voxel_dims)),
    })
    np.savez_compressed(os.path.join(OUT, "postproc.npz"), **vols)
    with open(os.path.join(OUT, "postproc.json"), "w") as f:
        json.dump(results, f, indent=1)


def voxelops_fixtures():
    """feature_extraction/utils.py:27-68 and step4_morphology.py:133-397 run on seeded synthetic cases, plus the SciPy
    primitives underneath them (the kernels in csrc/morph.cu are checked against both)."""
    from scipy import ndimage as ndi
    ns = R.load_reference()
    shape = (48, 40, 36)
    vols, results = {}, {}
    for seed in (0, 1):
        pred, _ = SY.label_pair(seed, shape)
        mri = SY.mri_volumes(seed, pred)
        vols[f"seg{seed}"] = pred
        for k, v in mri.items():
            assert np.array_equal(v, v.astype(np.int16).astype(np.float32))
            vols[f"{k}{seed}"] = v.astype(np.int16)  # integer-valued, like int16 NIfTI data
        d = {k: v.astype(np.float64) for k, v in mri.items()}  # what nibabel's get_fdata() hands the reference
        segf = pred.astype(np.float64)
        masks = ns.utils.get_tumor_masks(segf)
        r = {}
        for tag, vd in (("iso", (1.0, 1.0, 1.0)), ("aniso", (0.9, 1.1, 1.25))):
            r[f"border_{tag}"] = ns.step4.analyze_border_regularity(masks["wt"], vd)
            r[f"margin_{tag}"] = ns.step4.analyze_margin_definition(d["t1ce"], segf, masks, vd)
            r[f"cystic_{tag}"] = ns.step4.analyze_cystic_vs_solid(d["t1"], d["t2"], d["flair"], segf, masks, vd)
        r["stats"] = {f"{mod}_{reg}": ns.utils.get_intensity_stats(d[mod], masks[reg])
                      for mod in ("t1", "t1ce", "t2", "flair") for reg in ("wt", "ncr", "et")}
        r["stats_empty"] = ns.utils.get_intensity_stats(d["t1"], np.zeros(shape, bool))
        r["normal_brain"] = {mod: ns.utils.get_normal_brain_stats(d[mod], segf) for mod in ("t1", "flair")}
        r["brain_mask_count"] = {str(p): int(ns.utils.get_brain_mask(d["t2"], p).sum()) for p in (5, 37.5)}
        wt = np.asarray(masks["wt"])
        r["erode"] = {str(k): int(ndi.binary_erosion(wt, iterations=k).sum()) for k in (1, 2, 3)}
        r["dilate"] = {str(k): int(ndi.binary_dilation(wt, iterations=k).sum()) for k in (1, 2, 5)}
        vols[f"dilate5_{seed}"] = ndi.binary_dilation(wt, iterations=5)
        vols[f"erode2_{seed}"] = ndi.binary_erosion(wt, iterations=2)
        edt = ndi.distance_transform_edt(wt)
        sq = np.rint(edt ** 2).astype(np.int32)  # isotropic: SciPy's value is sqrt(exact integer) — store the integer
        assert np.array_equal(np.sqrt(sq.astype(np.float64)), edt)
        vols[f"edt_in_sq_{seed}"] = sq
        if seed == 0:
            vols["edt_out_aniso_0"] = ndi.distance_transform_edt(~wt, sampling=(0.9, 1.1, 1.25))
        for conn, rank in ((6, 1), (18, 2), (26, 3)):
            lab, n = ndi.label(pred == 3, structure=ndi.generate_binary_structure(3, rank))
            vols[f"label{conn}_{seed}"] = lab.astype(np.int32)
            r[f"label{conn}_count"] = int(n)
        results[str(seed)] = jsonable(r)
    np.savez_compressed(os.path.join(OUT, "voxelops.npz"), **vols)
    with open(os.path.join(OUT, "voxelops.json"), "w") as f:
        json.dump(results, f, indent=1)


def write_case_folder(folder, case_id, seg_xyz, mri_xyz, zooms=(0.9, 1.1, 1.25)):
    """A BraTS-2021 style case folder from (x, y, z) arrays (nibabel's axis order), written with the package's own
    NIfTI writer; returns the label file's path.  Shared with tests/test_gpu_drivers.py."""
    from brainseg_b200 import nifti_io

    os.makedirs(folder, exist_ok=True)
    like = nifti_io.new_header(seg_xyz.shape[::-1], zooms)
    for mod, vol in mri_xyz.items():
        nifti_io.save(os.path.join(folder, f"{case_id}_{mod}.nii.gz"), np.ascontiguousarray(vol.transpose(2, 1, 0)), like)
    seg_path = os.path.join(folder, f"{case_id}_seg.nii.gz")
    nifti_io.save(seg_path, np.ascontiguousarray(seg_xyz.transpose(2, 1, 0)).astype(np.uint8),
                  nifti_io.new_header(seg_xyz.shape[::-1], zooms, dtype=np.uint8))
    return seg_path


def driver_fixtures():
    """The reference's file-level step drivers (step3_multiplicity.py:445-546, step4_morphology.py:602-687) on a case
    folder built from the voxelops fixture volumes.  nibabel is absent: the reference modules' `load_nifti` is
    replaced by a reader over the package's NIfTI parser that returns what nibabel would ((x, y, z) float64 data,
    header.get_zooms()); everything after the load is the reference's own code.  `text_summary` (report text) is
    dropped."""
    import contextlib
    import io
    import tempfile

    from brainseg_b200.feature_extraction import utils as BU

    ns = R.load_reference()
    vols = np.load(os.path.join(OUT, "voxelops.npz"))
    out = {}
    for seed in (0, 1):
        with tempfile.TemporaryDirectory() as tmp:
            case = f"BraTS2021_{seed:05d}"
            folder = os.path.join(tmp, case)
            mri = {k: vols[f"{k}{seed}"].astype(np.float32) for k in ("t1", "t1ce", "t2", "flair")}
            zooms = (1.0, 1.0, 1.0) if seed == 0 else (0.9, 1.1, 1.25)  # BraTS spacing / an anisotropic header
            seg_path = write_case_folder(folder, case, vols[f"seg{seed}"], mri, zooms)
            for mod in (ns.step3, ns.step4):
                mod.load_nifti = BU.load_nifti
            with contextlib.redirect_stdout(io.StringIO()):
                r3 = ns.step3.analyze_multiplicity(folder, seg_path)
                r4 = ns.step4.analyze_morphology(folder, seg_path)
            r3.pop("text_summary")
            r4.pop("text_summary")
            out[str(seed)] = jsonable({"step3": r3, "step4": r4})
    with open(os.path.join(OUT, "step_drivers.json"), "w") as f:
        json.dump(out, f, indent=1)


def sliding_window_facts():
    g = SW.get_gaussian((128, 128, 128), 1.0 / 8)
    facts = {
        "source": "restated nnU-Net v1 (UPSTREAM, not in /root/reference); values from SURVEY.md §8c/§8d",
        "steps": {
            "155x240x240_p128_s0.5": SW.compute_steps_for_sliding_window((128,) * 3, (155, 240, 240), 0.5),
            "155x240x240_p128_s0.25": SW.compute_steps_for_sliding_window((128,) * 3, (155, 240, 240), 0.25),
            "256_p128_s0.5": SW.compute_steps_for_sliding_window((128,) * 3, (256,) * 3, 0.5),
            "256_p128_s0.25": SW.compute_steps_for_sliding_window((128,) * 3, (256,) * 3, 0.25),
            "256_p160_s0.5": SW.compute_steps_for_sliding_window((160,) * 3, (256,) * 3, 0.5),
            "256_p160_s0.25": SW.compute_steps_for_sliding_window((160,) * 3, (256,) * 3, 0.25),
            "137x171x140_p128_s0.5": SW.compute_steps_for_sliding_window((128,) * 3, (137, 171, 140), 0.5),
        },
        "survey_expected_steps": {
            "155x240x240_p128_s0.5": [[0, 27], [0, 56, 112], [0, 56, 112]],
            "155x240x240_p128_s0.25": [[0, 27], [0, 28, 56, 84, 112], [0, 28, 56, 84, 112]],
            "256_p128_s0.5": [[0, 64, 128]] * 3,
            "256_p128_s0.25": [[0, 32, 64, 96, 128]] * 3,
            "256_p160_s0.5": [[0, 48, 96]] * 3,
            "256_p160_s0.25": [[0, 32, 64, 96]] * 3,
        },
        "gaussian_128": {"min": float(g.min()), "max": float(g.max()), "corner": float(g[0, 0, 0]),
                         "center": float(g[64, 64, 64]), "survey_min": 3.7751344e-11},
        "ensemble_lut": [[int(np.round((a + b) / 2.0)) for b in range(4)] for a in range(4)],
        "survey_ensemble_lut": [[0, 0, 1, 2], [0, 1, 2, 2], [1, 2, 2, 2], [2, 2, 2, 3]],
    }
    with open(os.path.join(OUT, "sliding_window.json"), "w") as f:
        json.dump(facts, f, indent=1)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["voxelops"]:
        voxelops_fixtures()
    elif sys.argv[1:] == ["drivers"]:
        driver_fixtures()
    else:
        unet_fixtures()
        postproc_fixtures()
        sliding_window_facts()
        voxelops_fixtures()
        driver_fixtures()
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))
