"""End-to-end drop-in test of run_brats2021_inference_singlethread.main on a synthetic case folder and an nnU-Net style
RESULTS_FOLDER with random-init checkpoints: file layout, NIfTI geometry, preprocessing -> per-fold prediction -> fold
mean -> regions export -> uncrop -> two-model label ensemble -> volumes, against the CPU oracle chain."""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import postproc as OP
from oracle import preprocess as OPP
from oracle import sliding_window as SW
from tests.helpers import build_dropin_unet, oracle_fns

pytestmark = pytest.mark.gpu

PATCH = (32, 32, 32)
TOL = 1e-2


def _write_results_folder(root, folds):
    from brainseg_b200 import run_brats2021_inference_singlethread as R

    nets = {}
    for name, kw in ((R.MODEL1, dict(variant="bn", base=16, num_pool=2)), (R.MODEL2, dict(variant="gn", base=16, num_pool=2, groups=4))):
        folder = os.path.join(root, "3d_fullres", "Task500_BraTS2021", name)
        os.makedirs(folder)
        plans = {"plans_per_stage": {0: {"patch_size": np.array(PATCH)}}, "use_mask_for_norm": {0: True, 1: True, 2: True, 3: True}}
        with open(os.path.join(folder, "plans.pkl"), "wb") as f:
            pickle.dump(plans, f)
        nets[name] = []
        for k in folds:
            net = build_dropin_unet(seed=100 + 10 * len(nets) + k, **kw)
            os.makedirs(os.path.join(folder, f"fold_{k}"))
            torch.save({"state_dict": net.state_dict(), "epoch": 1000},
                       os.path.join(folder, f"fold_{k}", "model_final_checkpoint.model"))
            nets[name].append(net)
    return nets


def test_cli_end_to_end(tmp_path, monkeypatch):
    from brainseg_b200 import nifti_io, nnunet_compat
    from brainseg_b200 import run_brats2021_inference_singlethread as R

    # GroupNorm group count is not recoverable from a state_dict: the compat layer defaults to 8, the test nets use 4
    orig = nnunet_compat.infer_network_config
    monkeypatch.setattr(nnunet_compat, "infer_network_config",
                        lambda sd, name="", ng=None: orig(sd, name, 4 if "Groupnorm" in name else ng))
    folds = (0, 1)
    nets = _write_results_folder(str(tmp_path / "results"), folds)
    data = OPP.synthetic_head(7, (40, 52, 46))
    like = nifti_io.new_header(data.shape[1:], (1.0, 1.0, 1.0))
    inp = tmp_path / "input"
    inp.mkdir()
    for c, mod in enumerate(("t1", "t1ce", "t2", "flair")):
        nifti_io.save(str(inp / f"BraTS_00001_{mod}.nii.gz"), data[c], like)
    out = tmp_path / "out"
    R.main(["--input", str(inp), "--output", str(out), "--results", str(tmp_path / "results"), "--folds", "0", "1"])

    # ---- oracle chain
    d_ref, seg_mask, bbox = OPP.preprocess_case(data)
    sl = tuple(slice(a, b) for a, b in bbox)
    segs, decisive = [], []
    for name in (R.MODEL1, R.MODEL2):
        probs = []
        for net in nets[name]:
            fwd, _, _ = oracle_fns(net)
            probs.append(SW.predict_3d_tiled(fwd, torch.sigmoid, d_ref, 3, PATCH, True, (0, 1, 2), 0.5, True, (1, 2, 3))[1])
        mean = np.mean(probs, axis=0)
        seg = np.zeros(mean.shape[1:], dtype=np.uint8)
        for i, c in enumerate((1, 2, 3)):
            seg[mean[i] > 0.5] = c
        full = np.zeros(data.shape[1:], dtype=np.uint8)
        full[sl] = seg
        dec = np.ones(data.shape[1:], dtype=bool)  # outside the crop box both sides write 0
        dec[sl] = np.all(np.abs(mean - 0.5) > TOL, axis=0)
        segs.append(full)
        decisive.append(dec)

    for i, name in enumerate(("temp_model1", "temp_model2")):
        im = nifti_io.load(str(out / name / "BraTS_00001.nii.gz"))
        assert im.data.dtype == np.uint8 and im.data.shape == data.shape[1:] and im.zooms == (1.0, 1.0, 1.0)
        agree = (im.data == segs[i])
        print(f"{name}: agreement {agree.mean() * 100:.3f}% ({decisive[i].mean() * 100:.1f}% decisive)")
        assert agree[decisive[i]].all()
    final = nifti_io.load(str(out / "BraTS_00001.nii.gz"))
    both = decisive[0] & decisive[1]
    assert np.array_equal(final.data[both], OP.ensemble_labels_round(segs[0], segs[1])[both])
    # volumes are computed from the file that was written
    vols = R.calculate_volumes(str(out / "BraTS_00001.nii.gz"))
    f = final.data
    vv = float(np.prod(final.zooms)) / 1000.0  # the reference's formula: count * (prod(zooms) / 1000)
    assert vols["NCR"] == (f == 1).sum() * vv and vols["ED"] == (f == 2).sum() * vv
    assert vols["ET"] == (f == 4).sum() * vv and vols["WT"] == ((f == 1).sum() + (f == 2).sum() + (f == 4).sum()) * vv
