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
# per output format: (converter, labels expected in the output, value the enhancing tumour maps to)
_FORMATS = {"brats2025": (convert_labels_to_brats2025, (0, 1, 2, 3), 3),
            "brats2021": (convert_labels_to_brats2021, (0, 1, 2, 4), 4)}


def convert_file(input_path, output_path, format="brats2025"):
    """Remap one NIfTI label file (reference convert_labels_to_brats.py:58-108).  The console text is the reference's,
    line for line (tests/golden/cli_convert_*.txt)."""
    from . import nifti_io

    convert, expected, et_label = _FORMATS[format if format in _FORMATS else "brats2021"]
    tag, rule = format.upper(), "=" * 70
    print(f"\n{rule}\nConverting: {input_path}\nFormat: {tag}\n{rule}")
    source = nifti_io.load(str(input_path))
    before = source.get_fdata()
    seen = np.unique(before)
    print(f"\nLabels before conversion: {seen}")
    after = convert(before)
    produced = np.unique(after)
    print(f"Labels after conversion:  {produced}")
    print(f"\nLabel mapping applied ({tag}):")
    for old, text in ((1, "  1 (ED) -> 2 (ED)"), (2, "  2 (NCR) -> 1 (NCR)"),
                      (3, f"  3 (ET) -> {et_label} (ET)  [CRITICAL CONVERSION]")):
        if old in seen:
            print(text)
    nifti_io.save(str(output_path), after, source)
    print(f"\n[OK] Saved converted segmentation to: {output_path}")
    print(f"\nExpected {tag} labels: {sorted(expected)}")
    print(f"Actual labels in output: {produced}")
    if set(produced) == set(expected):
        print(f"[OK] SUCCESS: All {tag} labels present!")
    elif et_label not in produced:
        print(f"[WARNING] Label {et_label} missing - check if input had label 3")


def main(argv=None):
    import argparse
    import sys
    from pathlib import Path

    cli = argparse.ArgumentParser(description="Convert nnU-Net labels to BraTS format")
    cli.add_argument("input", help="Input NIfTI file with nnU-Net labels [0,1,2,3]")
    cli.add_argument("output", nargs="?", help="Output NIfTI file (optional, defaults to input_brats.nii.gz)")
    cli.add_argument("--format", choices=sorted(_FORMATS, reverse=True), default="brats2025",
                     help="Output format: brats2025 (default, ET=3) or brats2021 (legacy, ET=4)")
    args = cli.parse_args(argv)
    src = Path(args.input)
    if not src.exists():
        print(f"[ERROR] Input file not found: {src}")
        sys.exit(1)
    dst = Path(args.output) if args.output else src.with_name(src.name.replace(".nii", "_brats.nii", 1))
    convert_file(src, dst, args.format)


if __name__ == "__main__":
    main()
