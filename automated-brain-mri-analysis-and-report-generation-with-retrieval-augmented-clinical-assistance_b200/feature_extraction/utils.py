"""Drop-in for the mask helpers of the reference's feature_extraction/utils.py (:167-216).

The reference materialises six boolean volumes; here a mask is a (label volume, label set) pair whose count, centroid,
bounding box, second moments and surface count all come from ONE device pass (bsg_masked_moments) shared by the
masks of a volume.
"""
import numpy as np
import torch

from .. import voxelops as V

_REGIONS = {  # utils.py:171-178
    "ncr": V.bits_of(1),
    "ed": V.bits_of(2),
    "et": V.bits_of(3, 4),
    "tc": V.bits_of(1, 3, 4),
    "wt": V.MASK_GT0,
}


class LabelVolume:
    def __init__(self, seg):
        self.vol = V.as_label_volume(seg)  # np.round(seg).astype(int) happens on the device
        if self.vol.dim() != 3:
            raise ValueError("expected a 3-D label volume")
        self._moments = {}

    def moments(self, bits):
        if bits not in self._moments:
            todo = [b for b in dict.fromkeys(list(_REGIONS.values()) + [bits]) if b not in self._moments][:8]
            if bits not in todo:
                todo = [bits]
            res = V.masked_moments(self.vol, todo, surface_flags=(1 << len(todo)) - 1)
            for b, r in zip(todo, res):
                self._moments[b] = r
        return self._moments[bits]


class LabelMask:
    """Lazy boolean mask `isin(volume, labels)`; `.sum()` and `np.asarray()` behave like the reference's arrays."""

    def __init__(self, parent, bits, background=False):
        self.parent, self.bits, self.background = parent, bits, background

    @property
    def shape(self):
        return tuple(self.parent.vol.shape)

    @property
    def stats(self):
        if self.background:
            raise ValueError("only the voxel count is defined for the background mask")
        return self.parent.moments(self.bits)

    def sum(self):
        if self.background:
            return int(self.parent.vol.numel()) - int(self.parent.moments(V.MASK_GT0)["count"])
        return int(self.stats["count"])

    def tensor(self):
        v = self.parent.vol.to(torch.int64)
        m = ((torch.tensor(self.bits, dtype=torch.int64, device=v.device) >> v.clamp(max=31)) & 1).bool() & (v < 32)
        return ~m if self.background else m

    def __array__(self, dtype=None, copy=None):
        a = self.tensor().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a


def as_mask(mask):
    """LabelMask as is; any other boolean/0-1 array becomes a single-label volume."""
    if isinstance(mask, LabelMask):
        return mask
    if isinstance(mask, np.ndarray):
        mask = mask.astype(np.uint8) if mask.dtype == bool else (mask != 0).astype(np.uint8)
    elif torch.is_tensor(mask):
        mask = (mask != 0).to(torch.uint8)
    return LabelMask(LabelVolume(mask), V.bits_of(1))


def get_tumor_masks(seg_data):
    """Masks for the tumour regions (reference utils.py:167-178): background, ncr, ed, et (3|4), tc (1|3|4), wt (>0)."""
    parent = seg_data if isinstance(seg_data, LabelVolume) else LabelVolume(seg_data)
    masks = {"background": LabelMask(parent, V.MASK_GT0, background=True)}
    for name, bits in _REGIONS.items():
        masks[name] = LabelMask(parent, bits)
    return masks


def calculate_volume(mask, voxel_volume_cm3):
    """Volume in cm³ of a mask (reference utils.py:181-183)."""
    return float(as_mask(mask).sum() * voxel_volume_cm3)


def get_centroid(mask):
    """Centroid of a mask in voxel coordinates (reference utils.py:186-197); None when empty."""
    m = as_mask(mask)
    if m.sum() == 0:
        return None
    s = m.stats
    n = int(s["count"])
    # np.mean of integer coordinates == exact integer sum / count in float64
    return {"x": float(int(s["s0"]) / n), "y": float(int(s["s1"]) / n), "z": float(int(s["s2"]) / n)}


def get_bounding_box(mask):
    """Bounding box of a mask (reference utils.py:200-216); None when empty."""
    m = as_mask(mask)
    if m.sum() == 0:
        return None
    s = m.stats
    mn = [int(s["mn0"]), int(s["mn1"]), int(s["mn2"])]
    mx = [int(s["mx0"]), int(s["mx1"]), int(s["mx2"])]
    return {"min_x": mn[0], "max_x": mx[0], "min_y": mn[1], "max_y": mx[1], "min_z": mn[2], "max_z": mx[2],
            "size_x": mx[0] - mn[0] + 1, "size_y": mx[1] - mn[1] + 1, "size_z": mx[2] - mn[2] + 1}


# ------------------------------------------------------------------------------------------------------------------
# Intensity statistics (reference utils.py:27-68)
# ------------------------------------------------------------------------------------------------------------------

def mask_u8(mask):
    """LabelMask / bool array / tensor -> cuda uint8 volume (non-zero = set)."""
    if isinstance(mask, LabelMask):
        return mask.tensor().to(torch.uint8)
    return V.as_mask(mask)


_EMPTY_STATS = ("mean", "std", "min", "max", "median", "q25", "q75")


def get_intensity_stats(data, mask):
    """mean / std / min / max / median / quartiles of data[mask > 0] (reference utils.py:27-51).  Moments are fp64
    device sums; the order statistics are exact (radix select) and interpolated the way np.percentile does."""
    d, m = V.as_intensity(data), mask_u8(mask)
    cnt, mean, std, lo, hi = V.intensity_moments(d, m)
    if cnt == 0:
        out = {k: None for k in _EMPTY_STATS}
        out["voxel_count"] = 0
        return out
    sel = V.MaskedValues(d, m)
    q25, q75 = sel.percentiles([25, 75])
    return {"mean": float(mean), "std": float(std), "min": float(lo), "max": float(hi), "median": float(sel.median()),
            "q25": q25, "q75": q75, "voxel_count": int(cnt)}


def _gt_threshold(data, thr):
    """data > thr for float32 data and a float64 threshold, decided as numpy decides it for float64 data."""
    t32 = np.float32(thr)
    if np.float64(t32) > thr:  # rounding went up: x > thr  <=>  x >= t32
        return data >= float(t32)
    return data > float(t32)


def get_brain_mask(data, threshold_percentile=5):
    """data > percentile(data[data > 0], p) (reference utils.py:62-67); cuda bool tensor."""
    d = V.as_intensity(data)
    sel = V.MaskedValues(d, None)
    if sel.count == 0:  # data.max() <= 0 (the reference tests == 0): nothing is brain
        return d > 0
    return _gt_threshold(d, sel.percentiles([threshold_percentile])[0])


def get_normal_brain_stats(data, seg_mask):
    """Intensity statistics of non-tumour brain tissue (reference utils.py:54-60)."""
    d = V.as_intensity(data)
    seg = seg_mask.vol if isinstance(seg_mask, LabelVolume) else V.as_label_volume(seg_mask)
    normal = get_brain_mask(d, 5) & (seg == 0)
    return get_intensity_stats(d, normal)


# ------------------------------------------------------------------------------------------------------------------
# File-level helpers of the step drivers (reference utils.py:15-24, 71-125, 218-247) on brainseg_b200.nifti_io
# ------------------------------------------------------------------------------------------------------------------

class NiftiHeaderView:
    """The two header queries the reference makes on a nibabel header."""

    def __init__(self, image):
        self._image = image

    def get_zooms(self):
        return tuple(np.float32(v) for v in self._image.zooms)  # (x, y, z) mm, float32 scalars as nibabel returns them

    def get_data_shape(self):
        return tuple(reversed(self._image.data.shape))


def load_nifti(filepath):
    """(data, affine, header) like the reference's nibabel loader: data is float64 in nibabel's (x, y, z) axis order;
    affine is None (nothing on this path reads it)."""
    from .. import nifti_io

    image = nifti_io.load(str(filepath))
    return np.ascontiguousarray(image.get_fdata().transpose(2, 1, 0)), None, NiftiHeaderView(image)


def get_case_id(input_folder):
    """Case id from the T1 file name (BraTS-2021 `<id>_t1.nii.gz`, else BraTS-2025 `<id>-t1n.nii.gz`), else the
    folder name."""
    from pathlib import Path

    folder = Path(input_folder)
    for pattern, cut in (("*_t1.nii.gz", "_t1"), ("*-t1n.nii.gz", "-t1")):
        hits = list(folder.glob(pattern))
        if hits:
            return hits[0].name.split(cut)[0]
    return folder.name


_MODALITY_SUFFIXES = (  # (probe, {modality: suffix})
    ("_t1.nii.gz", {"t1": "_t1", "t1ce": "_t1ce", "t2": "_t2", "flair": "_flair"}),
    ("-t1n.nii.gz", {"t1": "-t1n", "t1ce": "-t1c", "t2": "-t2w", "flair": "-t2f"}),
)


def get_mri_paths(input_folder, case_id=None):
    from pathlib import Path

    folder = Path(input_folder)
    case_id = get_case_id(folder) if case_id is None else case_id
    for probe, names in _MODALITY_SUFFIXES:
        if (folder / f"{case_id}{probe}").exists():
            return {mod: folder / f"{case_id}{suffix}.nii.gz" for mod, suffix in names.items()}
    raise ValueError(f"Could not find MRI files in {folder}")


def get_voxel_dimensions(header):
    dims = header.get_zooms()[:3]
    return {"dimensions_mm": list(dims), "volume_mm3": float(np.prod(dims)), "volume_cm3": float(np.prod(dims) / 1000)}


class NumpyEncoder(__import__("json").JSONEncoder):
    def default(self, obj):
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        if isinstance(obj, np.bool_):
            return bool(obj)
        return super().default(obj)


def save_results(results, output_path):
    import json
    from pathlib import Path

    target = Path(output_path)
    target.parent.mkdir(parents=True, exist_ok=True)
    with open(target, "w") as f:
        json.dump(results, f, indent=2, cls=NumpyEncoder)
    print(f"Results saved to: {target}")


def load_results(json_path):
    import json

    with open(json_path) as f:
        return json.load(f)
