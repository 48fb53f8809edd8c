#!/usr/bin/env python
"""Drop-in for the reference's run_brats2021_inference_singlethread.py: same function names, arguments, on-disk layout
and console wording, with the work done by the sm_100a engine.

    python -m brainseg_b200.run_brats2021_inference_singlethread --input <dir> --output <dir> [--results <RESULTS_FOLDER>]

writes `<out>/temp_model{1,2}/<case>.nii.gz` and the ensembled `<out>/<case>.nii.gz` (reference :272-308).
"""
import argparse
import os
import re
import shutil
import sys
from pathlib import Path

import numpy as np
import torch

from . import nifti_io
from . import voxelops as V
from .nnunet_compat import load_model_and_checkpoint_files, save_segmentation_nifti_from_softmax

MODEL1 = "nnUNetTrainerV2BraTSRegions_DA4_BN_BD__nnUNetPlansv2.1"
MODEL2 = "nnUNetTrainerV2BraTSRegions_DA4_BN_BD_largeUnet_Groupnorm__nnUNetPlansv2.1"
# BraTS modality suffix -> nnU-Net channel index (reference :46-51)
CHANNEL_OF = (("t1", "0000"), ("t1ce", "0001"), ("t2", "0002"), ("flair", "0003"))
_BRATS_NAME = re.compile(r"^(?P<case>.+)_(?P<mod>t1|t1ce|t2|flair|seg)\.nii\.gz$")
RULE = "=" * 70


def _banner(*lines, lead="\n"):
    print(lead + "\n".join((RULE,) + lines + (RULE,)))


def prepare_input(sample_dir, output_dir):
    """Copies `<case>_{t1,t1ce,t2,flair}.nii.gz` to the nnU-Net names `<case>_{0000..0003}.nii.gz` (reference :25-78).
    Returns [(case, [four file names])] for the cases that have every modality."""
    src_dir, dst_dir = Path(sample_dir), Path(output_dir)
    dst_dir.mkdir(parents=True, exist_ok=True)
    cases = {m.group("case") for m in (_BRATS_NAME.match(p.name) for p in src_dir.glob("*.nii.gz")) if m}
    print(f"Found {len(cases)} cases: {cases}")
    ready = []
    for case in sorted(cases):
        staged = []
        for mod, idx in CHANNEL_OF:
            src = src_dir / f"{case}_{mod}.nii.gz"
            if not src.exists():
                print(f"[WARNING] Missing {mod} for {case}")
                staged = None
                break
            dst = dst_dir / f"{case}_{idx}.nii.gz"
            if not dst.exists():
                shutil.copy(src, dst)
            staged.append(str(dst))
        if staged:
            ready.append((case, staged))
    return ready


def predict_case_single_threaded(trainer, list_of_files, output_file, params, do_tta=True, mixed_precision=True,
                                 step_size=0.5, all_in_gpu=False):
    """One case through every fold of one model (reference :81-158): preprocess, per-fold probabilities, np.mean over
    the folds (on the host, like the reference), regions export to `output_file`."""
    print(f"Preprocessing {output_file}")
    d, _, dct = trainer.preprocess_patient(list_of_files)
    print(f"Data shape after preprocessing: {tuple(d.shape)}")
    print(f"Predicting {output_file}")
    kwargs = dict(do_mirroring=do_tta, mirror_axes=trainer.data_aug_params["mirror_axes"], use_sliding_window=True,
                  step_size=step_size, use_gaussian=True, all_in_gpu=all_in_gpu, mixed_precision=mixed_precision)
    per_fold = []
    for weights in params:
        trainer.load_checkpoint_ram(weights, False)
        per_fold.append(trainer.predict_preprocessed_data_return_seg_and_softmax(d, **kwargs)[1])
    print(f"Ensembling {len(per_fold)} folds")
    softmax_mean = np.mean(per_fold, axis=0)
    export = trainer.plans.get("segmentation_export_params", {}) if isinstance(trainer.plans, dict) else {}
    print(f"Saving segmentation to {output_file}")
    save_segmentation_nifti_from_softmax(softmax_mean, output_file, dct, export.get("interpolation_order", 1), (1, 2, 3),
                                         None, None, None, None, force_separate_z=export.get("force_separate_z"),
                                         interpolation_order_z=export.get("interpolation_order_z", 0))
    return output_file


def run_model_single_threaded(model_folder, input_folder, output_folder, folds=(0, 1, 2, 3, 4)):
    """Every case of `input_folder` through one model (reference :161-214)."""
    model_folder, out_dir = Path(model_folder), Path(output_folder)
    if not model_folder.exists():
        print(f"[ERROR] Model not found: {model_folder}")
        sys.exit(1)
    print(f"Model path: {model_folder}")
    print(f"Loading model with folds: {folds}")
    torch.cuda.empty_cache()
    trainer, params = load_model_and_checkpoint_files(str(model_folder), folds, mixed_precision=True,
                                                      checkpoint_name="model_final_checkpoint")
    print(f"Loaded {len(params)} fold checkpoints")
    todo = prepare_input(Path(input_folder), out_dir / "temp_input")
    if not todo:
        print("[ERROR] No valid cases found!")
        return
    out_dir.mkdir(parents=True, exist_ok=True)
    for case_name, case_files in todo:
        target = out_dir / f"{case_name}.nii.gz"
        _banner(f"Processing case: {case_name}")
        predict_case_single_threaded(trainer=trainer, list_of_files=case_files, output_file=str(target), params=params,
                                     do_tta=True, mixed_precision=True, step_size=0.5, all_in_gpu=False)
        print(f"✓ Completed: {target}")


def calculate_volumes(seg_path):
    """Tumour sub-region volumes in cm³ of a BraTS-labelled file: voxel counts of labels 1 / 2 / 4 times the voxel
    volume (reference :217-243), counted on the device."""
    img = nifti_io.load(seg_path)
    cm3 = float(np.prod(img.zooms)) / 1000.0
    labels = V.as_label_volume(img.data if img.data.dtype == np.uint8 else img.get_fdata())
    ncr, ed, et = (int(rec["count"]) for rec in V.masked_moments(labels, [V.bits_of(1), V.bits_of(2), V.bits_of(4)]))
    return {"NCR": ncr * cm3, "ED": ed * cm3, "ET": et * cm3, "TC": (ncr + et) * cm3, "WT": (ncr + ed + et) * cm3}


def ensemble_case(seg1_path, seg2_path, final_output):
    """np.round((seg1 + seg2) / 2.0).astype(np.uint8) of the two models' label files (reference :299-308)."""
    first, second = nifti_io.load(seg1_path), nifti_io.load(seg2_path)
    merged = V.ensemble_round(V.as_label_volume(first.get_fdata()), V.as_label_volume(second.get_fdata()))
    nifti_io.save(final_output, merged.cpu().numpy(), first)
    return final_output


VOLUME_LINES = (("NCR", "NCR (Necrotic Core):       "), ("ED", "ED (Peritumoral Edema):    "), ("ET", "ET (Enhancing Tumor):      "),
                ("TC", "TC (Tumor Core):           "), ("WT", "WT (Whole Tumor):          "))


def main(argv=None):
    cli = argparse.ArgumentParser(description="BraTS 2021 Brain Tumor Segmentation (Single-threaded)")
    cli.add_argument("--input", type=str, required=True, help="Input directory with BraTS sample data")
    cli.add_argument("--output", type=str, required=True, help="Output directory for segmentation results")
    cli.add_argument("--results", type=str, default=None, help="RESULTS_FOLDER (default: $RESULTS_FOLDER or ./nnUNet_results)")
    cli.add_argument("--folds", type=int, nargs="*", default=[0, 1, 2, 3, 4])
    args = cli.parse_args(argv)
    results = Path(args.results or os.environ.get("RESULTS_FOLDER") or Path.cwd() / "nnUNet_results")
    os.environ["RESULTS_FOLDER"] = str(results)
    _banner("BraTS 2021 TUMOR SEGMENTATION (SINGLE-THREADED)", lead="")
    print(f"RESULTS_FOLDER: {results}\n")
    task_dir, out_dir = results / "3d_fullres" / "Task500_BraTS2021", Path(args.output)
    member_dirs = []
    for i, name in enumerate((MODEL1, MODEL2), start=1):
        _banner(f"MODEL {i}: {name.split('__')[0]}")
        member_dirs.append(out_dir / f"temp_model{i}")
        run_model_single_threaded(task_dir / name, args.input, member_dirs[-1], tuple(args.folds))
    _banner("ENSEMBLING MODEL PREDICTIONS")
    for seg1 in sorted(member_dirs[0].glob("*.nii.gz")):
        case_name = seg1.name[:-len(".nii.gz")]
        seg2 = member_dirs[1] / seg1.name
        if not seg2.exists():
            print(f"[WARNING] Missing model2 prediction for {case_name}")
            continue
        print(f"Ensembling {case_name}")
        final_output = ensemble_case(seg1, seg2, out_dir / f"{case_name}.nii.gz")
        print(f"✓ Saved: {final_output}")
        volumes = calculate_volumes(final_output)
        print(f"\nTumor Volume Analysis for {case_name}:")
        for key, label in VOLUME_LINES:
            print(f"  {label} {volumes[key]:.2f} cm³")
    _banner("SEGMENTATION COMPLETE!")
    print(f"Results saved to: {out_dir}")


if __name__ == "__main__":
    main()
