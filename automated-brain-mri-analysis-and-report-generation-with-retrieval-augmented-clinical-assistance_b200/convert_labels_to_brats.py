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
