"""Golden console output of the REFERENCE's file-level CLIs (build container only; needs /root/reference):
evaluate_segmentation.evaluate_segmentation(pred, gt) and convert_labels_to_brats.convert_file(in, out, format), run on
seeded label volumes through a nibabel stand-in that serves arrays from memory.  run_full_pipeline.py:252-269 parses
this text, so the drop-ins must print it verbatim.  Writes tests/golden/cli_*.txt.

    python oracle/make_golden_cli.py
"""
import contextlib
import io
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import synthetic as SY  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SHAPE = (48, 40, 36)  # nibabel order (x, y, z)


def install_fake_nibabel(files):
    nib = types.ModuleType("nibabel")

    class Img:
        def __init__(self, data, affine=None, header=None):
            self._d, self.affine, self.header = data, affine, header

        def get_fdata(self):
            return np.asarray(self._d, dtype=np.float64)

    nib.load = lambda p: Img(files[os.path.basename(str(p))])
    nib.Nifti1Image = Img
    nib.save = lambda img, p: files.__setitem__(os.path.basename(str(p)), np.asarray(img._d))
    sys.modules["nibabel"] = nib


def main():
    pred, gt = SY.label_pair(0, SHAPE)
    files = {"pred.nii.gz": pred, "gt.nii.gz": gt}
    install_fake_nibabel(files)
    from oracle import ref_import

    ref = ref_import.load_reference()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref.evaluate.evaluate_segmentation("pred.nii.gz", "gt.nii.gz")
    with open(os.path.join(OUT, "cli_evaluate.txt"), "w") as f:
        f.write(buf.getvalue())
    for fmt in ("brats2025", "brats2021"):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ref.convert_labels.convert_file("pred.nii.gz", f"pred_{fmt}.nii.gz", fmt)
        with open(os.path.join(OUT, f"cli_convert_{fmt}.txt"), "w") as f:
            f.write(buf.getvalue())
    print("wrote", [n for n in os.listdir(OUT) if n.startswith("cli_")])


if __name__ == "__main__":
    main()
