"""Drop-in for the label-remap functions of the reference's convert_labels_to_brats.py (:34-55).

Same names, argument meaning and results; the 4 boolean-mask passes become one LUT kernel (bsg_label_lut_u8).
numpy in -> numpy uint8 out (like the reference); cuda tensor in -> cuda uint8 tensor out (stays on device).
"""
import numpy as np
import torch

from . import voxelops as V


def _lut(et_label):
    lut = np.zeros(256, dtype=np.uint8)  # "any other value -> 0" (np.zeros_like, :39/:51)
    lut[1], lut[2], lut[3] = 2, 1, et_label  # ED 1->2, NCR 2->1, ET 3->3|4
    return lut


LUT_BRATS2025 = _lut(3)
LUT_BRATS2021 = _lut(4)


def _convert(seg, lut):
    out = V.label_lut(V.as_label_volume(seg), lut)
    if torch.is_tensor(seg):
        return out
    return out.cpu().numpy()


def convert_labels_to_brats2025(seg):
    """nnU-Net labels (0,1,2,3) -> BraTS 2025 labels (0,2,1,3); reference convert_labels_to_brats.py:34-43."""
    return _convert(seg, LUT_BRATS2025)


def convert_labels_to_brats2021(seg):
    """nnU-Net labels (0,1,2,3) -> BraTS 2021 labels (0,2,1,4); reference convert_labels_to_brats.py:46-55."""
    return _convert(seg, LUT_BRATS2021)


# ---------------------------------------------------------------------------------------------- file-level CLI
def convert_file(input_path, output_path, format="brats2025"):
    """Convert a single NIfTI file (reference convert_labels_to_brats.py:58-108, same console wording)."""
    from . import nifti_io

    print(f"\n{'=' * 70}")
    print(f"Converting: {input_path}")
    print(f"Format: {format.upper()}")
    print(f"{'=' * 70}")
    img = nifti_io.load(str(input_path))
    data = img.get_fdata()
    unique_before = np.unique(data)
    print(f"\nLabels before conversion: {unique_before}")
    if format == "brats2025":
        data_converted, expected_labels, et_label = convert_labels_to_brats2025(data), {0, 1, 2, 3}, 3
    else:
        data_converted, expected_labels, et_label = convert_labels_to_brats2021(data), {0, 1, 2, 4}, 4
    unique_after = np.unique(data_converted)
    print(f"Labels after conversion:  {unique_after}")
    print(f"\nLabel mapping applied ({format.upper()}):")
    if 1 in unique_before:
        print("  1 (ED) -> 2 (ED)")
    if 2 in unique_before:
        print("  2 (NCR) -> 1 (NCR)")
    if 3 in unique_before:
        print(f"  3 (ET) -> {et_label} (ET)  [CRITICAL CONVERSION]")
    nifti_io.save(str(output_path), data_converted, img)
    print(f"\n[OK] Saved converted segmentation to: {output_path}")
    print(f"\nExpected {format.upper()} labels: {sorted(expected_labels)}")
    print(f"Actual labels in output: {unique_after}")
    if set(unique_after) == expected_labels:
        print(f"[OK] SUCCESS: All {format.upper()} labels present!")
    elif et_label not in unique_after:
        print(f"[WARNING] Label {et_label} missing - check if input had label 3")


def main(argv=None):
    import argparse
    import sys
    from pathlib import Path

    parser = argparse.ArgumentParser(description="Convert nnU-Net labels to BraTS format")
    parser.add_argument("input", help="Input NIfTI file with nnU-Net labels [0,1,2,3]")
    parser.add_argument("output", nargs="?", help="Output NIfTI file (optional, defaults to input_brats.nii.gz)")
    parser.add_argument("--format", choices=["brats2025", "brats2021"], default="brats2025",
                        help="Output format: brats2025 (default, ET=3) or brats2021 (legacy, ET=4)")
    args = parser.parse_args(argv)
    input_file = Path(args.input)
    output_file = Path(args.output) if args.output else input_file.parent / (
        input_file.stem.replace(".nii", "_brats.nii") + ".gz")
    if not input_file.exists():
        print(f"[ERROR] Input file not found: {input_file}")
        sys.exit(1)
    convert_file(input_file, output_file, args.format)


if __name__ == "__main__":
    main()
