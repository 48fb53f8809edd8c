"""numpy / scipy restatement of the reference's voxel post-processing (the consumers of predict_3D's output).

Each function follows the cited reference lines; dict keys and arithmetic types (float32 Dice sums, float64 means,
Python-int counts) are kept so results can be compared bit-for-bit with the reference functions themselves
(tests/test_oracle_golden.py does that against fixtures generated from the imported reference).
"""
import numpy as np
from scipy import ndimage
from scipy.ndimage import label, binary_erosion

MIN_LESION_VOLUME_CM3 = 0.1  # feature_extraction/step3_multiplicity.py:38


# ------------------------------------------------------------------ ensemble + remap
def ensemble_labels_round(seg1, seg2):
    """run_brats2021_inference_singlethread.py:305 — np.round((seg1+seg2)/2) (half-to-even) -> uint8."""
    return np.round((seg1.astype(np.float64) + seg2.astype(np.float64)) / 2.0).astype(np.uint8)


def convert_labels_to_brats2025(seg):
    """convert_labels_to_brats.py:34-43."""
    seg = np.round(seg).astype(np.uint8)
    new_seg = np.zeros_like(seg)
    new_seg[seg == 1] = 2
    new_seg[seg == 2] = 1
    new_seg[seg == 3] = 3
    return new_seg


def convert_labels_to_brats2021(seg):
    """convert_labels_to_brats.py:46-55."""
    seg = np.round(seg).astype(np.uint8)
    new_seg = np.zeros_like(seg)
    new_seg[seg == 1] = 2
    new_seg[seg == 2] = 1
    new_seg[seg == 3] = 4
    return new_seg


def calculate_volumes(seg, voxel_dims=(1.0, 1.0, 1.0)):
    """run_brats2021_inference_singlethread.py:217-243 (labels 1, 2, 4)."""
    voxel_volume_cm3 = np.prod(voxel_dims) / 1000.0
    ncr = np.sum(seg == 1)
    ed = np.sum(seg == 2)
    et = np.sum(seg == 4)
    return {"NCR": ncr * voxel_volume_cm3, "ED": ed * voxel_volume_cm3, "ET": et * voxel_volume_cm3,
            "TC": (ncr + et) * voxel_volume_cm3, "WT": (ncr + ed + et) * voxel_volume_cm3}


# ------------------------------------------------------------------ Dice
def calculate_metrics(pred, gt, label_value):
    """evaluate_segmentation.py:12-49 — float32 mask products and sums."""
    pred_mask = (pred == label_value).astype(np.float32)
    gt_mask = (gt == label_value).astype(np.float32)
    tp = np.sum(pred_mask * gt_mask)
    fp = np.sum(pred_mask * (1 - gt_mask))
    fn = np.sum((1 - pred_mask) * gt_mask)
    tn = np.sum((1 - pred_mask) * (1 - gt_mask))
    dice = (2 * tp) / (2 * tp + fp + fn + 1e-8)
    iou = tp / (tp + fp + fn + 1e-8)
    sensitivity = tp / (tp + fn + 1e-8)
    specificity = tn / (tn + fp + 1e-8)
    return {"dice": dice, "iou": iou, "sensitivity": sensitivity, "specificity": specificity,
            "tp": tp, "fp": fp, "fn": fn, "tn": tn}


def calculate_metrics_binary(pred_mask, gt_mask):
    """evaluate_segmentation.py:181-195."""
    tp = np.sum(pred_mask * gt_mask)
    fp = np.sum(pred_mask * (1 - gt_mask))
    fn = np.sum((1 - pred_mask) * gt_mask)
    dice = (2 * tp) / (2 * tp + fp + fn + 1e-8)
    iou = tp / (tp + fp + fn + 1e-8)
    sensitivity = tp / (tp + fn + 1e-8)
    return {"dice": dice, "iou": iou, "sensitivity": sensitivity}


def evaluate_arrays(pred_data, gt_data):
    """Arithmetic of evaluate_segmentation.py:84-162 without file I/O or printing.

    Returns {"labels": {label: metrics}, "wt": ..., "tc": ..., "et": metrics|None, "mean_dice": float}."""
    if pred_data.shape != gt_data.shape:
        return None  # :78-81
    pred_labels = np.unique(pred_data)
    gt_labels = np.unique(gt_data)
    all_metrics = {}
    for lab in sorted(set(pred_labels) | set(gt_labels)):
        if lab == 0:
            continue
        all_metrics[lab] = calculate_metrics(pred_data, gt_data, lab)
    wt = calculate_metrics_binary(np.isin(pred_data, [1, 2, 3]).astype(np.float32),
                                  np.isin(gt_data, [1, 2, 3]).astype(np.float32))
    tc = calculate_metrics_binary(np.isin(pred_data, [1, 3]).astype(np.float32),
                                  np.isin(gt_data, [1, 3]).astype(np.float32))
    et = all_metrics.get(3)
    mean_dice = np.mean([wt["dice"], tc["dice"], et["dice"] if et is not None else 0])
    return {"labels": all_metrics, "wt": wt, "tc": tc, "et": et, "mean_dice": mean_dice}


# ------------------------------------------------------------------ step3: connected components
def detect_connected_components(seg_data, voxel_dims):
    """feature_extraction/step3_multiplicity.py:41-152."""
    tumor_mask = seg_data > 0
    if tumor_mask.sum() == 0:
        return {"num_components": 0, "components": [], "is_single_lesion": True, "description": "No tumor detected"}
    structure = ndimage.generate_binary_structure(3, 3)
    labeled_array, num_components = label(tumor_mask, structure=structure)
    components = []
    for comp_id in range(1, num_components + 1):
        comp_mask = labeled_array == comp_id
        comp_voxels = comp_mask.sum()
        volume_cm3 = comp_voxels * np.prod(voxel_dims) / 1000
        coords = np.where(comp_mask)
        centroid = {"x": float(np.mean(coords[0])), "y": float(np.mean(coords[1])), "z": float(np.mean(coords[2]))}
        centroid_mm = {"x": centroid["x"] * voxel_dims[0], "y": centroid["y"] * voxel_dims[1],
                       "z": centroid["z"] * voxel_dims[2]}
        bbox = {"x_min": int(coords[0].min()), "x_max": int(coords[0].max()),
                "y_min": int(coords[1].min()), "y_max": int(coords[1].max()),
                "z_min": int(coords[2].min()), "z_max": int(coords[2].max())}
        max_diameter_mm = max((bbox["x_max"] - bbox["x_min"]) * voxel_dims[0],
                              (bbox["y_max"] - bbox["y_min"]) * voxel_dims[1],
                              (bbox["z_max"] - bbox["z_min"]) * voxel_dims[2])
        comp_labels = seg_data[comp_mask]
        composition = {"ncr": int((comp_labels == 1).sum()), "ed": int((comp_labels == 2).sum()),
                       "et": int((comp_labels == 3).sum())}
        components.append({"id": comp_id, "voxel_count": int(comp_voxels), "volume_cm3": float(volume_cm3),
                           "centroid_voxel": centroid, "centroid_mm": centroid_mm, "bounding_box": bbox,
                           "max_diameter_mm": float(max_diameter_mm), "composition": composition,
                           "has_enhancement": composition["et"] > 0})
    significant = [c for c in components if c["volume_cm3"] >= MIN_LESION_VOLUME_CM3]
    noise = [c for c in components if c["volume_cm3"] < MIN_LESION_VOLUME_CM3]
    significant.sort(key=lambda c: c["volume_cm3"], reverse=True)
    for i, comp in enumerate(significant):
        comp["rank"] = i + 1
        comp["classification"] = "Primary lesion" if i == 0 else f"Secondary lesion #{i}"
    n_sig = len(significant)
    noise_note = (f" ({len(noise)} sub-threshold fragments excluded, <{MIN_LESION_VOLUME_CM3} cm³)" if noise else "")
    return {"num_components": n_sig, "components": significant, "is_single_lesion": n_sig == 1,
            "description": f"{n_sig} lesion(s) detected{noise_note}", "excluded_fragments": len(noise),
            "minimum_volume_threshold_cm3": MIN_LESION_VOLUME_CM3}


def analyze_enhancing_components(seg_data, voxel_dims):
    """feature_extraction/step3_multiplicity.py:207-263."""
    et_mask = seg_data == 3
    if et_mask.sum() == 0:
        return {"num_enhancing_foci": 0, "enhancing_components": [], "pattern": "Non-enhancing",
                "description": "No enhancing tumor components detected"}
    structure = ndimage.generate_binary_structure(3, 3)
    labeled_et, n_et = label(et_mask, structure=structure)
    comps = []
    for comp_id in range(1, n_et + 1):
        comp_mask = labeled_et == comp_id
        volume_cm3 = comp_mask.sum() * np.prod(voxel_dims) / 1000
        coords = np.where(comp_mask)
        centroid_mm = {"x": float(np.mean(coords[0]) * voxel_dims[0]), "y": float(np.mean(coords[1]) * voxel_dims[1]),
                       "z": float(np.mean(coords[2]) * voxel_dims[2])}
        comps.append({"id": comp_id, "volume_cm3": float(volume_cm3), "centroid_mm": centroid_mm})
    comps.sort(key=lambda c: c["volume_cm3"], reverse=True)
    if n_et == 0:
        pattern = "Non-enhancing"
    elif n_et == 1:
        pattern = "Single enhancing focus"
    elif n_et <= 3:
        pattern = "Few enhancing foci"
    else:
        pattern = "Multiple/scattered enhancing foci"
    return {"num_enhancing_foci": n_et, "enhancing_components": comps, "pattern": pattern,
            "total_enhancing_volume_cm3": float(sum(c["volume_cm3"] for c in comps)),
            "description": f"{n_et} separate enhancing focus/foci detected"}


def label_components(mask):
    """scipy.ndimage.label with the 26-connected structure (step3:58-59): int32 labels in raster order."""
    return label(mask, structure=ndimage.generate_binary_structure(3, 3))


# ------------------------------------------------------------------ utils + step4
def get_tumor_masks(seg_data):
    """feature_extraction/utils.py:167-178."""
    seg_data = np.round(seg_data).astype(np.int32)
    return {"background": seg_data == 0, "ncr": seg_data == 1, "ed": seg_data == 2,
            "et": (seg_data == 3) | (seg_data == 4), "tc": (seg_data == 1) | (seg_data == 3) | (seg_data == 4),
            "wt": seg_data > 0}


def calculate_volume(mask, voxel_volume_cm3):
    """feature_extraction/utils.py:181-183."""
    return float(mask.sum() * voxel_volume_cm3)


def get_centroid(mask):
    """feature_extraction/utils.py:186-197."""
    if mask.sum() == 0:
        return None
    coords = np.array(np.where(mask)).T
    c = coords.mean(axis=0)
    return {"x": float(c[0]), "y": float(c[1]), "z": float(c[2])}


def get_bounding_box(mask):
    """feature_extraction/utils.py:200-216."""
    if mask.sum() == 0:
        return None
    coords = np.where(mask)
    return {"min_x": int(coords[0].min()), "max_x": int(coords[0].max()),
            "min_y": int(coords[1].min()), "max_y": int(coords[1].max()),
            "min_z": int(coords[2].min()), "max_z": int(coords[2].max()),
            "size_x": int(coords[0].max() - coords[0].min() + 1),
            "size_y": int(coords[1].max() - coords[1].min() + 1),
            "size_z": int(coords[2].max() - coords[2].min() + 1)}


def calculate_surface_area(mask, voxel_dims):
    """feature_extraction/step4_morphology.py:33-55 — 6-connected erosion, border_value=0."""
    if mask.sum() == 0:
        return 0.0
    eroded = binary_erosion(mask)
    surface_voxels = mask & ~eroded
    avg_face_area = (voxel_dims[0] * voxel_dims[1] + voxel_dims[1] * voxel_dims[2] + voxel_dims[0] * voxel_dims[2]) / 3
    return float(surface_voxels.sum() * avg_face_area)


def surface_voxel_count(mask):
    """Integer core of calculate_surface_area (step4:42-45)."""
    return int((mask & ~binary_erosion(mask)).sum())


def calculate_sphericity(volume_mm3, surface_area_mm2):
    """step4_morphology.py:58-75."""
    if surface_area_mm2 == 0 or volume_mm3 == 0:
        return 0.0
    radius = (3 * volume_mm3 / (4 * np.pi)) ** (1 / 3)
    sphere_surface = 4 * np.pi * radius ** 2
    return float(min(1.0, max(0.0, sphere_surface / surface_area_mm2)))


def calculate_compactness(volume_mm3, surface_area_mm2):
    """step4_morphology.py:118-130."""
    if surface_area_mm2 == 0:
        return 0.0
    return float(min(1.0, (36 * np.pi * volume_mm3 ** 2) / (surface_area_mm2 ** 3)))


def calculate_elongation(mask, voxel_dims):
    """step4_morphology.py:78-115 — np.cov (ddof=1) of mm-scaled coordinates, eigvalsh."""
    coords = np.where(mask)
    if len(coords[0]) < 10:
        return 1.0, [1.0, 1.0, 1.0]
    points = np.array([coords[0] * voxel_dims[0], coords[1] * voxel_dims[1], coords[2] * voxel_dims[2]]).T
    centered = points - points.mean(axis=0)
    cov = np.cov(centered.T)
    eig = np.sort(np.linalg.eigvalsh(cov))[::-1]
    elongation = np.sqrt(eig[0] / eig[-1]) if eig[-1] > 0 else 1.0
    return float(elongation), [float(np.sqrt(e) * 2) for e in eig]


def calculate_shape_descriptors(seg_data, tumor_masks, voxel_dims):
    """step4_morphology.py:483-541."""
    wt_mask = tumor_masks["wt"]
    if wt_mask.sum() == 0:
        return {"volume_cm3": 0, "surface_area_mm2": 0, "sphericity": 0, "compactness": 0, "elongation": 1.0,
                "principal_axes_mm": [0, 0, 0]}
    volume_mm3 = wt_mask.sum() * np.prod(voxel_dims)
    volume_cm3 = volume_mm3 / 1000
    surface_area = calculate_surface_area(wt_mask, voxel_dims)
    sphericity = calculate_sphericity(volume_mm3, surface_area)
    compactness = calculate_compactness(volume_mm3, surface_area)
    elongation, principal_axes = calculate_elongation(wt_mask, voxel_dims)
    if sphericity > 0.8:
        shape_class = "Spherical/round"
    elif sphericity > 0.6:
        shape_class = "Ovoid"
    elif sphericity > 0.4:
        shape_class = "Irregular"
    else:
        shape_class = "Highly irregular/complex"
    if elongation > 2.5:
        elongation_class = "Elongated"
    elif elongation > 1.5:
        elongation_class = "Mildly elongated"
    else:
        elongation_class = "Roughly isotropic"
    return {"volume_cm3": float(volume_cm3), "surface_area_mm2": float(surface_area), "sphericity": float(sphericity),
            "compactness": float(compactness), "elongation": float(elongation), "principal_axes_mm": principal_axes,
            "shape_classification": shape_class, "elongation_classification": elongation_class}


def analyze_necrosis_pattern(seg_data, tumor_masks, voxel_dims):
    """step4_morphology.py:400-480."""
    ncr_mask, tc_mask, wt_mask = tumor_masks["ncr"], tumor_masks["tc"], tumor_masks["wt"]
    ncr_volume = ncr_mask.sum() * np.prod(voxel_dims) / 1000
    tc_volume = tc_mask.sum() * np.prod(voxel_dims) / 1000
    wt_volume = wt_mask.sum() * np.prod(voxel_dims) / 1000
    if wt_volume == 0:
        return {"necrosis_present": False, "pattern": "No tumor", "description": "No tumor detected"}
    if ncr_volume == 0:
        return {"necrosis_present": False, "necrosis_volume_cm3": 0, "necrosis_percentage": 0,
                "pattern": "No necrosis", "description": "No central necrosis identified, solid tumor"}
    necrosis_pct = (ncr_volume / wt_volume) * 100
    if ncr_mask.sum() > 0 and tc_mask.sum() > 0:
        ncr_coords = np.where(ncr_mask)
        tc_coords = np.where(tc_mask)
        ncr_centroid = np.array([np.mean(ncr_coords[i]) for i in range(3)])
        tc_centroid = np.array([np.mean(tc_coords[i]) for i in range(3)])
        dist = np.linalg.norm((ncr_centroid - tc_centroid) * voxel_dims)
        tc_radius = (3 * tc_volume * 1000 / (4 * np.pi)) ** (1 / 3)
        if dist < tc_radius * 0.3:
            location, location_description = "Central", "Necrosis centered within tumor"
        elif dist < tc_radius * 0.6:
            location, location_description = "Eccentric", "Necrosis somewhat offset from tumor center"
        else:
            location, location_description = "Peripheral", "Necrosis located eccentrically"
    else:
        location, location_description = "Undetermined", "Could not determine necrosis location"
    if necrosis_pct > 50:
        pattern = "Extensive necrosis"
        description = (f"Large central necrotic component ({necrosis_pct:.0f}% of tumor), "
                       "characteristic of high-grade glioma")
    elif necrosis_pct > 25:
        pattern = "Moderate necrosis"
        description = f"Moderate central necrosis ({necrosis_pct:.0f}% of tumor), suggests high-grade pathology"
    elif necrosis_pct > 10:
        pattern = "Focal necrosis"
        description = f"Focal areas of necrosis ({necrosis_pct:.0f}% of tumor)"
    else:
        pattern = "Minimal necrosis"
        description = f"Small necrotic foci ({necrosis_pct:.0f}% of tumor)"
    return {"necrosis_present": True, "necrosis_volume_cm3": float(ncr_volume),
            "necrosis_percentage": float(necrosis_pct), "pattern": pattern, "location": location,
            "location_description": location_description, "description": description}
