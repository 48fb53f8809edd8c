"""File-level drivers on the GPU: feature_extraction step 3 / step 4 (`analyze_multiplicity`, `analyze_morphology`),
`feature_extraction/run_all.py` and `run_full_pipeline.py` on synthetic case folders.

Golden results (tests/golden/step_drivers.json) come from the reference's own drivers run on the same folders
(oracle/make_golden.py drivers).  nibabel hands the reference its voxel sizes as float32 scalars, and under NumPy >= 2
(NEP 50) `python_float * np.float32` stays float32 — the reference's physical-unit fields (centroid_mm, volumes,
diameters ...) are therefore rounded to float32 even for a 1 mm header.  The package computes them in float64 from the
same zooms, so those fields agree to float32 resolution (2e-6 relative); counts, labels, classifications and every
voxel-unit field agree exactly."""
import contextlib
import io
import json
import os

import numpy as np
import pytest

from oracle.make_golden import write_case_folder
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


def tree_close(a, b, rtol, path=""):
    if isinstance(a, dict):
        assert set(a.keys()) == set(b.keys()), f"{path}: {set(a.keys()) ^ set(b.keys())}"
        for k in a:
            tree_close(a[k], b[k], rtol, f"{path}/{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            tree_close(x, y, rtol, f"{path}[{i}]")
    elif isinstance(a, (float, np.floating)) and isinstance(b, (int, float, np.floating, np.integer)):
        assert abs(float(a) - float(b)) <= rtol * max(abs(float(a)), abs(float(b))), f"{path}: {a!r} != {b!r}"
    else:
        assert a == b, f"{path}: {a!r} != {b!r}"


def _case(tmp_path, seed, zooms, names2025=False):
    vols = np.load(os.path.join(GOLDEN, "voxelops.npz"))
    case = f"BraTS2021_{seed:05d}"
    mri = {k: vols[f"{k}{seed}"].astype(np.float32) for k in ("t1", "t1ce", "t2", "flair")}
    folder = str(tmp_path / case)
    seg_path = write_case_folder(folder, case, vols[f"seg{seed}"], mri, zooms)
    return folder, seg_path, case


@pytest.mark.parametrize("seed,zooms,rtol", [(0, (1.0, 1.0, 1.0), 2e-6), (1, (0.9, 1.1, 1.25), 2e-6)])
def test_step_drivers_match_reference(tmp_path, seed, zooms, rtol):
    from brainseg_b200.feature_extraction import step3_multiplicity as S3
    from brainseg_b200.feature_extraction import step4_morphology as S4
    from brainseg_b200.feature_extraction import utils as U

    with open(os.path.join(GOLDEN, "step_drivers.json")) as f:
        ref = json.load(f)[str(seed)]
    folder, seg_path, case = _case(tmp_path, seed, zooms)
    with contextlib.redirect_stdout(io.StringIO()):
        r3 = S3.analyze_multiplicity(folder, seg_path, str(tmp_path / "s3.json"))
        r4 = S4.analyze_morphology(folder, seg_path, str(tmp_path / "s4.json"))
    assert r3["case_id"] == case
    # through JSON, as the consumers read it (numpy scalars -> floats / ints)
    tree_close(U.load_results(str(tmp_path / "s3.json")), ref["step3"], rtol, "step3")
    tree_close(U.load_results(str(tmp_path / "s4.json")), ref["step4"], rtol, "step4")
    assert r3["component_analysis"]["num_components"] == ref["step3"]["component_analysis"]["num_components"]


def test_run_full_pipeline_end_to_end(tmp_path, monkeypatch, capsys):
    from brainseg_b200 import nnunet_compat
    from brainseg_b200 import run_full_pipeline as RP
    from brainseg_b200.feature_extraction import step3_multiplicity as S3
    from brainseg_b200.feature_extraction import utils as U
    from brainseg_b200 import evaluate_segmentation as EV
    from tests.test_gpu_cli import _write_results_folder

    orig = nnunet_compat.infer_network_config
    monkeypatch.setattr(nnunet_compat, "infer_network_config",
                        lambda sd, name="", ng=None: orig(sd, name, 4 if "Groupnorm" in name else ng))
    _write_results_folder(str(tmp_path / "models"), (0,))
    folder, seg_path, case = _case(tmp_path, 0, (1.0, 1.0, 1.0))
    # BraTS-2025 naming on disk: step 1 must rename it (the folder name is the case id)
    for old, new in (("t1", "t1n"), ("t1ce", "t1c"), ("t2", "t2w"), ("flair", "t2f"), ("seg", "seg")):
        os.rename(os.path.join(folder, f"{case}_{old}.nii.gz"), os.path.join(folder, f"{case}-{new}.nii.gz"))
    with pytest.raises(SystemExit) as stop:
        RP.main([folder, "--results-root", str(tmp_path / "results"), "--models", str(tmp_path / "models"), "--folds", "0"])
    out = capsys.readouterr().out
    assert stop.value.code == 0, out[-2000:]
    stages = [line for line in out.splitlines() if line.startswith("STAGE:")]
    assert stages == ["STAGE:segmenting", "STAGE:extracting", "STAGE:generating", "STAGE:exporting", "STAGE:done"]
    res = tmp_path / "results" / case
    for name in (f"{case}.nii.gz", f"{case}_brats.nii.gz", "pipeline_summary.json", "feature_extraction/step3_multiplicity.json",
                 "feature_extraction/step4_morphology.json", "feature_extraction/comprehensive_analysis.json"):
        assert (res / name).exists(), name
    summary = U.load_results(str(res / "pipeline_summary.json"))
    assert set(summary) == {"case_id", "timestamp", "pipeline_duration_minutes", "input_folder", "output_folder",
                            "segmentation_file", "converted_file", "ground_truth_file", "feature_extraction_folder",
                            "gemini_report", "pdf_report", "metrics"}
    # the metrics are the evaluation's own numbers at the 2 decimals the text carries
    with contextlib.redirect_stdout(io.StringIO()):
        ev = EV.evaluate_segmentation(str(res / f"{case}_brats.nii.gz"), os.path.join(folder, f"{case}_seg.nii.gz"))
    assert ev is not None and set(summary["metrics"]) <= {"mean_dice", "wt_dice", "tc_dice", "et_dice"}
    assert "mean_dice" in summary["metrics"]
    # step 3 written by the pipeline == step 3 computed directly on the converted file
    with contextlib.redirect_stdout(io.StringIO()):
        direct = S3.analyze_multiplicity(folder, str(res / f"{case}_brats.nii.gz"))
    tree_close(U.load_results(str(res / "feature_extraction" / "step3_multiplicity.json")),
               json.loads(json.dumps(direct, cls=U.NumpyEncoder)), 0.0)

    # error paths: missing ground truth -> exit 1 after STAGE:error / ERROR:; missing models -> exit 2
    os.remove(os.path.join(folder, f"{case}_seg.nii.gz"))
    with pytest.raises(SystemExit) as stop:
        RP.main([folder, "--results-root", str(tmp_path / "results2"), "--models", str(tmp_path / "models")])
    out = capsys.readouterr().out
    assert stop.value.code == 1 and "STAGE:error" in out and "ERROR:Ground truth segmentation not found" in out
    folder2, _, _ = _case(tmp_path / "again", 1, (1.0, 1.0, 1.0))
    with pytest.raises(SystemExit) as stop:
        RP.main([folder2, "--results-root", str(tmp_path / "results3"), "--models", str(tmp_path / "no_models")])
    out = capsys.readouterr().out
    assert stop.value.code == 2 and "STAGE:error" in out
