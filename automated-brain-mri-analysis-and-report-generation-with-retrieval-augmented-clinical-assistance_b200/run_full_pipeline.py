"""Drop-in for the CLI of the reference's run_full_pipeline.py (`run_full_pipeline.py <case_folder>`): the same
stages, `STAGE:` / `ERROR:` markers (read by the reference's api.py), exit codes (0 / 1 missing input / 2 runtime
failure / 130 interrupt) and results layout, with steps 2-5 — segmentation, label conversion, evaluation, feature
extraction steps 3 and 4 — executed IN THIS PROCESS on the GPU instead of four subprocesses (reference :146-298).

    python -m brainseg_b200.run_full_pipeline <case_folder> [--results-root DIR] [--models RESULTS_FOLDER] [--folds ...]

    <results-root>/<CaseID>/<CaseID>.nii.gz            raw segmentation (two-model ensemble)
                            <CaseID>_brats.nii.gz      BraTS label convention
                            pipeline_summary.json      paths + Dice metrics (the reference's keys)
                            feature_extraction/        step3_multiplicity.json, step4_morphology.json, ...

Steps 6-8 of the reference (LLM radiology report, PDF export, RAG assistant) are out of scope here; their stage
markers are still printed so a front end tracking progress sees the same sequence.
"""
import argparse
import contextlib
import gzip
import io
import json
import re
import shutil
import sys
import time
from datetime import datetime
from pathlib import Path

MODALITIES = ("t1", "t1ce", "t2", "flair")
# BraTS-2025 file suffixes -> BraTS-2021 names (reference run_full_pipeline.py SUFFIX_MAPPING)
_NEW_TO_OLD = {"t1n": "t1", "t1c": "t1ce", "t2w": "t2", "t2f": "flair", "seg": "seg"}
_BRATS2025_NAME = re.compile(r"^(.+)-(t1n|t1c|t2w|t2f|seg)\.(nii\.gz|nii)$")


def _header(text):
    print(f"\n{'=' * 70}\n{text}\n{'=' * 70}")


def _step(number, text):
    print(f"\n{'─' * 70}\nSTEP {number}: {text}\n{'─' * 70}")


def rename_brats2025_files(case_folder):
    """`<id>-t1n.nii[.gz]` ... -> `<folder>_t1.nii.gz` ... in place (plain .nii files are gzip-compressed).
    Returns (case_id, files_renamed, already_converted)."""
    folder = Path(case_folder)
    case_id = folder.name
    renamed = already = 0
    for path in sorted(folder.iterdir()):
        if not path.is_file():
            continue
        hit = _BRATS2025_NAME.match(path.name)
        if hit is None:
            if re.match(rf"^{re.escape(case_id)}_({'|'.join(MODALITIES)}|seg)\.nii\.gz$", path.name):
                already += 1
            continue
        target = folder / f"{case_id}_{_NEW_TO_OLD[hit.group(2)]}.nii.gz"
        if target.exists():
            print(f"  ⚠ Target exists, skipping: {target.name}")
            continue
        if hit.group(3) == "nii":
            print(f"  📦 Compressing: {path.name} → {target.name}")
            with open(path, "rb") as src, gzip.open(target, "wb") as dst:
                shutil.copyfileobj(src, dst)
            path.unlink()
        else:
            print(f"  📝 Renaming: {path.name} → {target.name}")
            path.rename(target)
        renamed += 1
    return case_id, renamed, already


def run_segmentation(case_folder, output_folder, models=None, folds=None):
    """Both models + ensemble through the drop-in inference script's main() (same files as the reference's step 2)."""
    from . import run_brats2021_inference_singlethread as inference

    case_folder, output_folder = Path(case_folder), Path(output_folder)
    argv = ["--input", str(case_folder), "--output", str(output_folder)]
    if models:
        argv += ["--results", str(models)]
    if folds:
        argv += ["--folds"] + [str(k) for k in folds]
    print(f"  🔄 Running inference on the GPU...\n  📂 Input: {case_folder}\n  📂 Output: {output_folder}")
    started = time.time()
    try:
        inference.main(argv)
    except SystemExit as stop:  # the script exits with 1 when the model folders are missing (reference :169-171)
        if stop.code not in (0, None):
            raise RuntimeError(f"Segmentation failed with return code {stop.code}") from None
    print(f"  ⏱ Inference completed in {time.time() - started:.1f} seconds")
    produced = output_folder / f"{case_folder.name}.nii.gz"
    if not produced.exists():
        raise FileNotFoundError(f"Expected output file not found: {produced}")
    return produced


def convert_labels(input_file, output_file):
    from . import convert_labels_to_brats as remap

    print("  🔄 Converting labels...")
    text = io.StringIO()
    with contextlib.redirect_stdout(text):
        remap.convert_file(str(input_file), str(output_file))
    for line in text.getvalue().split("\n"):
        if "Labels" in line or "SUCCESS" in line or "mapping" in line.lower():
            print(f"  {line}")
    return Path(output_file)


_METRIC_LINES = (  # (key, must contain, must not contain) — the reference greps its evaluation script's stdout
    ("mean_dice", "Mean Dice Score", None),
    ("wt_dice", "Whole Tumor", "Dice"),
    ("tc_dice", "Tumor Core", "Dice"),
    ("et_dice", "Enhancing Tumor", "Label"),
)


def evaluate_segmentation(pred_file, gt_file):
    """Dice summary in percent, parsed from the evaluation text exactly as the reference parses it (:252-269)."""
    from . import evaluate_segmentation as evaluation

    print("  🔄 Evaluating segmentation...")
    text = io.StringIO()
    with contextlib.redirect_stdout(text):
        evaluation.evaluate_segmentation(str(pred_file), str(gt_file))
    report = text.getvalue()
    print(report)
    metrics = {}
    for line in report.split("\n"):
        for key, needle, veto in _METRIC_LINES:
            if needle in line and (veto is None or veto not in line):
                number = re.search(r"(\d+\.\d+)%", line)
                if number:
                    metrics[key] = float(number.group(1))
                break
    return metrics


def run_feature_extraction(mri_folder, segmentation_file, output_folder):
    from .feature_extraction import run_all

    output_folder = Path(output_folder)
    output_folder.mkdir(parents=True, exist_ok=True)
    print("  🔄 Running feature extraction pipeline...")
    run_all.run_all_steps(mri_folder, segmentation_file, output_folder)
    return output_folder


def run_pipeline(case_folder, results_root=None, models=None, folds=None):
    case_folder = Path(case_folder).absolute()
    if not case_folder.exists():
        raise FileNotFoundError(f"Case folder not found: {case_folder}")
    case_id = case_folder.name
    results_folder = Path(results_root or Path.cwd() / "results") / case_id
    _header("BRAIN MRI ANALYSIS PIPELINE")
    print(f"\n📋 Case ID: {case_id}\n📂 Input folder: {case_folder}\n📂 Results folder: {results_folder}")
    print(f"🕐 Started: {datetime.now().strftime('%Y-%m-%d %H:%M:%S')}")
    started = time.time()
    try:
        _step(1, "RENAMING FILES (BraTS 2025 → BraTS 2021 format)")
        case_id, renamed, already = rename_brats2025_files(case_folder)
        if renamed > 0:
            print(f"\n  ✅ Renamed {renamed} files")
        elif already > 0:
            print(f"\n  ✅ Files already in correct format ({already} files)")
        else:
            print("\n  ⚠ No files found to rename")
        missing = [m for m in MODALITIES if not (case_folder / f"{case_id}_{m}.nii.gz").exists()]
        if missing:
            raise FileNotFoundError(f"Missing required MRI files: {missing}")
        gt_file = case_folder / f"{case_id}_seg.nii.gz"
        if not gt_file.exists():
            raise FileNotFoundError(f"Ground truth segmentation not found: {gt_file}")
        print(f"  ✅ All required files present\n  ✅ Ground truth found: {gt_file.name}")

        print("STAGE:segmenting")
        _step(2, "RUNNING SEGMENTATION (BraTS 2021 KAIST Model)")
        results_folder.mkdir(parents=True, exist_ok=True)
        seg_output = run_segmentation(case_folder, results_folder, models, folds)
        print(f"\n  ✅ Segmentation complete: {seg_output.name}")

        _step(3, "CONVERTING LABELS")
        converted = convert_labels(seg_output, results_folder / f"{case_id}_brats.nii.gz")
        print(f"\n  ✅ Labels converted: {converted.name}")

        _step(4, "EVALUATING SEGMENTATION")
        metrics = evaluate_segmentation(converted, gt_file)
        if metrics:
            print("\n  📊 Summary:")
            for key, label in (("mean_dice", "Mean Dice"), ("wt_dice", "Whole Tumor"), ("tc_dice", "Tumor Core"),
                               ("et_dice", "Enhancing Tumor")):
                if key in metrics:
                    print(f"     {label}: {metrics[key]:.2f}%")

        print("STAGE:extracting")
        _step(5, "RUNNING FEATURE EXTRACTION PIPELINE")
        features = run_feature_extraction(case_folder, converted, results_folder / "feature_extraction")
        print(f"\n  ✅ Feature extraction complete\n  📂 Output folder: {features}")

        print("STAGE:generating")
        _step(6, "GENERATING RADIOLOGY REPORT")
        print("\n  ⚠ Skipped: report generation is not part of brainseg_b200")
        print("STAGE:exporting")
        _step(7, "GENERATING PROFESSIONAL PDF REPORT")
        print("\n  ⚠ Skipped (text report required first)")

        elapsed = time.time() - started
        _header("PIPELINE COMPLETE")
        print(f"\n⏱ Total time: {elapsed:.1f} seconds\n🕐 Completed: {datetime.now().strftime('%Y-%m-%d %H:%M:%S')}")
        summary = {
            "case_id": case_id,
            "timestamp": datetime.now().isoformat(),
            "pipeline_duration_minutes": round(elapsed / 60, 2),
            "input_folder": str(case_folder),
            "output_folder": str(results_folder),
            "segmentation_file": str(seg_output),
            "converted_file": str(converted),
            "ground_truth_file": str(gt_file),
            "feature_extraction_folder": str(features),
            "gemini_report": None,
            "pdf_report": None,
            "metrics": metrics,
        }
        summary_file = results_folder / "pipeline_summary.json"
        with open(summary_file, "w") as f:
            json.dump(summary, f, indent=2)
        print(f"\n📄 Pipeline summary saved: {summary_file}")
        print("STAGE:done")
        return summary
    except Exception as failure:
        print("STAGE:error")
        print(f"ERROR:{failure}")
        raise


def main(argv=None):
    cli = argparse.ArgumentParser(description="Automated Brain MRI Analysis Pipeline (GPU, in-process steps 1-5)")
    cli.add_argument("case_folder", help="Path to the case folder (e.g., BraTS-GLI-00003-000)")
    cli.add_argument("--results-root", default=None, help="where results/<CaseID>/ is created (default: ./results)")
    cli.add_argument("--models", default=None, help="nnU-Net RESULTS_FOLDER with the two BraTS-2021 trainers' folds")
    cli.add_argument("--folds", type=int, nargs="+", default=None, help="folds to ensemble per model (default: 0-4)")
    args = cli.parse_args(argv)
    try:
        run_pipeline(args.case_folder, args.results_root, args.models, args.folds)
        sys.exit(0)
    except FileNotFoundError as failure:
        print(f"\n❌ Error: {failure}")
        sys.exit(1)
    except RuntimeError as failure:
        print(f"\n❌ Error: {failure}")
        sys.exit(2)
    except KeyboardInterrupt:
        print("\n\n⚠ Pipeline interrupted by user")
        sys.exit(130)
    except Exception as failure:  # noqa: BLE001 — the reference maps every other failure to exit code 1
        print(f"\n❌ Unexpected error: {failure}")
        import traceback

        traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
