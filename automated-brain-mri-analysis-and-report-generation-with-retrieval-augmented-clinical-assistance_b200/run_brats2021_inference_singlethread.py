#!/usr/bin/env python
"""Drop-in for the reference's run_brats2021_inference_singlethread.py (same functions, arguments, file layout and
console wording), running on the sm_100a engine:

    python -m brainseg_b200.run_brats2021_inference_singlethread --input <dir> --output <dir> [--results <RESULTS_FOLDER>]

produces `<out>/temp_model{1,2}/<case>.nii.gz` and the ensembled `<out>/<case>.nii.gz` (reference :272-308).
"""
import argparse
import os
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


def prepare_input(sample_dir, output_dir):
    """BraTS names (case_t1/t1ce/t2/flair.nii.gz) -> nnU-Net names (case_0000..0003.nii.gz) (reference :25-78)."""
    sample_dir, output_dir = Path(sample_dir), Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    cases = set()
    for file in sample_dir.glob("*.nii.gz"):
        parts = file.stem.replace(".nii", "").split("_")
        if parts[-1] in ["t1", "t1ce", "t2", "flair", "seg"]:
            cases.add("_".join(parts[:-1]))
    print(f"Found {len(cases)} cases: {cases}")
    modality_map = {"t1": "0000", "t1ce": "0001", "t2": "0002", "flair": "0003"}
    prepared_cases = []
    for case in sorted(cases):
        case_files, all_found = [], True
        for mod, idx in modality_map.items():
            src, dst = sample_dir / f"{case}_{mod}.nii.gz", output_dir / f"{case}_{idx}.nii.gz"
            if src.exists():
                if not dst.exists():
                    shutil.copy(src, dst)
                case_files.append(str(dst))
            else:
                print(f"[WARNING] Missing {mod} for {case}")
                all_found = False
                break
        if all_found:
            prepared_cases.append((case, case_files))
    return prepared_cases


def predict_case_single_threaded(trainer, list_of_files, output_file, params, do_tta=True, mixed_precision=True,
                                 step_size=0.5, all_in_gpu=False):
    """One case, all folds of one model (reference :81-158): preprocess, per-fold softmax, mean over folds, regions
    export.  The fold mean is np.mean over the per-fold probability volumes, taken on the host like the reference."""
    print(f"Preprocessing {output_file}")
    d, _, dct = trainer.preprocess_patient(list_of_files)
    print(f"Data shape after preprocessing: {tuple(d.shape)}")
    print(f"Predicting {output_file}")
    all_softmax = []
    for p in params:
        trainer.load_checkpoint_ram(p, False)
        all_softmax.append(trainer.predict_preprocessed_data_return_seg_and_softmax(
            d, do_mirroring=do_tta, mirror_axes=trainer.data_aug_params["mirror_axes"], use_sliding_window=True,
            step_size=step_size, use_gaussian=True, all_in_gpu=all_in_gpu, mixed_precision=mixed_precision)[1])
    print(f"Ensembling {len(all_softmax)} folds")
    softmax_mean = np.mean(all_softmax, axis=0)
    export = trainer.plans.get("segmentation_export_params", {}) if isinstance(trainer.plans, dict) else {}
    print(f"Saving segmentation to {output_file}")
    save_segmentation_nifti_from_softmax(softmax_mean, output_file, dct, export.get("interpolation_order", 1), (1, 2, 3),
                                         None, None, None, None, force_separate_z=export.get("force_separate_z"),
                                         interpolation_order_z=export.get("interpolation_order_z", 0))
    return output_file


def run_model_single_threaded(model_folder, input_folder, output_folder, folds=(0, 1, 2, 3, 4)):
    """All cases of a folder through one model (reference :161-214)."""
    model_folder, input_folder, output_folder = Path(model_folder), Path(input_folder), Path(output_folder)
    if not model_folder.exists():
        print(f"[ERROR] Model not found: {model_folder}")
        sys.exit(1)
    print(f"Model path: {model_folder}")
    print(f"Loading model with folds: {folds}")
    torch.cuda.empty_cache()
    trainer, params = load_model_and_checkpoint_files(str(model_folder), folds, mixed_precision=True,
                                                      checkpoint_name="model_final_checkpoint")
    print(f"Loaded {len(params)} fold checkpoints")
    prepared_cases = prepare_input(input_folder, output_folder / "temp_input")
    if not prepared_cases:
        print("[ERROR] No valid cases found!")
        return
    for case_name, case_files in prepared_cases:
        output_file = output_folder / f"{case_name}.nii.gz"
        output_folder.mkdir(parents=True, exist_ok=True)
        print(f"\n{'=' * 70}\nProcessing case: {case_name}\n{'=' * 70}")
        predict_case_single_threaded(trainer=trainer, list_of_files=case_files, output_file=str(output_file), params=params,
                                     do_tta=True, mixed_precision=True, step_size=0.5, all_in_gpu=False)
        print(f"✓ Completed: {output_file}")


def calculate_volumes(seg_path):
    """Tumour volumes in cm³ from a BraTS-labelled segmentation file (reference :217-243)."""
    img = nifti_io.load(seg_path)
    voxel_volume_cm3 = float(np.prod(img.zooms)) / 1000.0
    lv = V.as_label_volume(img.data if img.data.dtype == np.uint8 else img.get_fdata())
    m = V.masked_moments(lv, [V.bits_of(1), V.bits_of(2), V.bits_of(4)])
    ncr, ed, et = (int(m[i]["count"]) for i in range(3))
    return {"NCR": ncr * voxel_volume_cm3, "ED": ed * voxel_volume_cm3, "ET": et * voxel_volume_cm3,
            "TC": (ncr + et) * voxel_volume_cm3, "WT": (ncr + ed + et) * voxel_volume_cm3}


def ensemble_case(seg1_path, seg2_path, final_output):
    """np.round((seg1 + seg2) / 2.0).astype(np.uint8) of the two models' label files (reference :299-308)."""
    im1, im2 = nifti_io.load(seg1_path), nifti_io.load(seg2_path)
    ens = V.ensemble_round(V.as_label_volume(im1.get_fdata()), V.as_label_volume(im2.get_fdata()))
    nifti_io.save(final_output, ens.cpu().numpy(), im1)
    return final_output


def main(argv=None):
    parser = argparse.ArgumentParser(description="BraTS 2021 Brain Tumor Segmentation (Single-threaded)")
    parser.add_argument("--input", type=str, required=True, help="Input directory with BraTS sample data")
    parser.add_argument("--output", type=str, required=True, help="Output directory for segmentation results")
    parser.add_argument("--results", type=str, default=None, help="RESULTS_FOLDER (default: ./nnUNet_results)")
    parser.add_argument("--folds", type=int, nargs="*", default=[0, 1, 2, 3, 4])
    args = parser.parse_args(argv)
    results_folder = Path(args.results or os.environ.get("RESULTS_FOLDER") or Path.cwd() / "nnUNet_results")
    os.environ["RESULTS_FOLDER"] = str(results_folder)
    print("=" * 70 + "\nBraTS 2021 TUMOR SEGMENTATION (SINGLE-THREADED)\n" + "=" * 70)
    print(f"RESULTS_FOLDER: {results_folder}\n")
    base = results_folder / "3d_fullres" / "Task500_BraTS2021"
    output_folder = Path(args.output)
    for i, name in enumerate((MODEL1, MODEL2), start=1):
        print("\n" + "=" * 70 + f"\nMODEL {i}: {name.split('__')[0]}\n" + "=" * 70)
        run_model_single_threaded(base / name, args.input, output_folder / f"temp_model{i}", tuple(args.folds))
    print("\n" + "=" * 70 + "\nENSEMBLING MODEL PREDICTIONS\n" + "=" * 70)
    model1_output, model2_output = output_folder / "temp_model1", output_folder / "temp_model2"
    for seg1_path in sorted(model1_output.glob("*.nii.gz")):
        case_name = seg1_path.stem.replace(".nii", "")
        seg2_path = model2_output / seg1_path.name
        if not seg2_path.exists():
            print(f"[WARNING] Missing model2 prediction for {case_name}")
            continue
        print(f"Ensembling {case_name}")
        final_output = output_folder / f"{case_name}.nii.gz"
        ensemble_case(seg1_path, seg2_path, final_output)
        print(f"✓ Saved: {final_output}")
        volumes = calculate_volumes(final_output)
        print(f"\nTumor Volume Analysis for {case_name}:")
        print(f"  NCR (Necrotic Core):        {volumes['NCR']:.2f} cm³")
        print(f"  ED (Peritumoral Edema):     {volumes['ED']:.2f} cm³")
        print(f"  ET (Enhancing Tumor):       {volumes['ET']:.2f} cm³")
        print(f"  TC (Tumor Core):            {volumes['TC']:.2f} cm³")
        print(f"  WT (Whole Tumor):           {volumes['WT']:.2f} cm³")
    print("\n" + "=" * 70 + "\nSEGMENTATION COMPLETE!\n" + "=" * 70)
    print(f"Results saved to: {output_folder}")


if __name__ == "__main__":
    main()
