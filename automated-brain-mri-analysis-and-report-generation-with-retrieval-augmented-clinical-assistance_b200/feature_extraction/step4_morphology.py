"""Drop-in for the shape statistics of the reference's feature_extraction/step4_morphology.py
(:33-130 surface / sphericity / elongation / compactness, :400-541 necrosis pattern and shape descriptors).

Counts, coordinate moments and the 6-connected surface-voxel count come from one device pass (bsg_masked_moments);
only closed-form scalar arithmetic remains on the host.
"""
import numpy as np

from .utils import as_mask


def calculate_surface_area(mask, voxel_dims):
    """Surface estimate: #(mask & ~binary_erosion(mask)) x mean voxel face area (reference :33-55)."""
    m = as_mask(mask)
    if m.sum() == 0:
        return 0.0
    avg_face_area = (voxel_dims[0] * voxel_dims[1] + voxel_dims[1] * voxel_dims[2] + voxel_dims[0] * voxel_dims[2]) / 3
    return float(int(m.stats["surface"]) * avg_face_area)


def calculate_sphericity(volume_mm3, surface_area_mm2):
    """Reference :58-75."""
    if surface_area_mm2 == 0 or volume_mm3 == 0:
        return 0.0
    radius = (3 * volume_mm3 / (4 * np.pi)) ** (1 / 3)
    return float(min(1.0, max(0.0, 4 * np.pi * radius ** 2 / surface_area_mm2)))


def calculate_compactness(volume_mm3, surface_area_mm2):
    """Reference :118-130."""
    if surface_area_mm2 == 0:
        return 0.0
    return float(min(1.0, (36 * np.pi * volume_mm3 ** 2) / (surface_area_mm2 ** 3)))


def calculate_elongation(mask, voxel_dims):
    """PCA elongation (reference :78-115).  The 3x3 covariance np.cov(points) (ddof=1) is assembled from exact
    integer coordinate moments: cov_ab = v_a v_b (n S_ab - S_a S_b) / (n (n-1))."""
    m = as_mask(mask)
    n = m.sum()
    if n < 10:
        return 1.0, [1.0, 1.0, 1.0]
    s = m.stats
    S = [int(s["s0"]), int(s["s1"]), int(s["s2"])]
    S2 = {(0, 0): int(s["s00"]), (1, 1): int(s["s11"]), (2, 2): int(s["s22"]), (0, 1): int(s["s01"]),
          (0, 2): int(s["s02"]), (1, 2): int(s["s12"])}
    cov = np.zeros((3, 3))
    for a in range(3):
        for b in range(a, 3):
            num = n * S2[(a, b)] - S[a] * S[b]  # exact Python integers
            cov[a, b] = cov[b, a] = (num / (n * (n - 1))) * voxel_dims[a] * voxel_dims[b]
    eigenvalues = np.sort(np.linalg.eigvalsh(cov))[::-1]
    elongation = np.sqrt(eigenvalues[0] / eigenvalues[-1]) if eigenvalues[-1] > 0 else 1.0
    return float(elongation), [float(np.sqrt(e) * 2) for e in eigenvalues]


def calculate_shape_descriptors(seg_data, tumor_masks, voxel_dims):
    """Whole-tumour shape descriptors (reference :483-541)."""
    wt_mask = tumor_masks["wt"]
    if wt_mask.sum() == 0:
        return {"volume_cm3": 0, "surface_area_mm2": 0, "sphericity": 0, "compactness": 0, "elongation": 1.0,
                "principal_axes_mm": [0, 0, 0]}
    volume_mm3 = wt_mask.sum() * np.prod(voxel_dims)
    surface_area = calculate_surface_area(wt_mask, voxel_dims)
    sphericity = calculate_sphericity(volume_mm3, surface_area)
    compactness = calculate_compactness(volume_mm3, surface_area)
    elongation, principal_axes = calculate_elongation(wt_mask, voxel_dims)
    shape_class = ("Spherical/round" if sphericity > 0.8 else "Ovoid" if sphericity > 0.6 else
                   "Irregular" if sphericity > 0.4 else "Highly irregular/complex")
    elongation_class = ("Elongated" if elongation > 2.5 else "Mildly elongated" if elongation > 1.5 else
                        "Roughly isotropic")
    return {"volume_cm3": float(volume_mm3 / 1000), "surface_area_mm2": float(surface_area),
            "sphericity": float(sphericity), "compactness": float(compactness), "elongation": float(elongation),
            "principal_axes_mm": principal_axes, "shape_classification": shape_class,
            "elongation_classification": elongation_class}


def _centroid(mask):
    s = mask.stats
    n = int(s["count"])
    return np.array([int(s["s0"]) / n, int(s["s1"]) / n, int(s["s2"]) / n])


def analyze_necrosis_pattern(seg_data, tumor_masks, voxel_dims):
    """Necrosis volume fraction and location relative to the tumour core (reference :400-480)."""
    ncr_mask, tc_mask, wt_mask = tumor_masks["ncr"], tumor_masks["tc"], tumor_masks["wt"]
    vox = np.prod(voxel_dims)
    ncr_volume = ncr_mask.sum() * vox / 1000
    tc_volume = tc_mask.sum() * vox / 1000
    wt_volume = wt_mask.sum() * vox / 1000
    if wt_volume == 0:
        return {"necrosis_present": False, "pattern": "No tumor", "description": "No tumor detected"}
    if ncr_volume == 0:
        return {"necrosis_present": False, "necrosis_volume_cm3": 0, "necrosis_percentage": 0,
                "pattern": "No necrosis", "description": "No central necrosis identified, solid tumor"}
    necrosis_pct = (ncr_volume / wt_volume) * 100
    if ncr_mask.sum() > 0 and tc_mask.sum() > 0:
        dist = np.linalg.norm((_centroid(ncr_mask) - _centroid(tc_mask)) * voxel_dims)
        tc_radius = (3 * tc_volume * 1000 / (4 * np.pi)) ** (1 / 3)
        if dist < tc_radius * 0.3:
            location, where = "Central", "Necrosis centered within tumor"
        elif dist < tc_radius * 0.6:
            location, where = "Eccentric", "Necrosis somewhat offset from tumor center"
        else:
            location, where = "Peripheral", "Necrosis located eccentrically"
    else:
        location, where = "Undetermined", "Could not determine necrosis location"
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
            "location_description": where, "description": description}


# ------------------------------------------------------------------------------------------------------------------
# Border regularity, margin definition, cystic / solid (reference :133-397) — device morphology, exact EDT, fp64
# reductions and exact percentiles (csrc/morph.cu); the scalar scoring stays on the host.
# ------------------------------------------------------------------------------------------------------------------

def _grade(score, table, default):
    """first (threshold, classification, description) whose threshold the score exceeds"""
    for thr, name, text in table:
        if score > thr:
            return name, text
    return default


_CONTOUR_GRADES = (
    (0.7, "Smooth contour", "Smooth, regular outer contour (note: does not indicate margin sharpness)"),
    (0.5, "Mildly lobulated", "Some contour irregularity with mild lobulation"),
    (0.3, "Lobulated", "Lobulated/irregular outer contour"),
)
_CONTOUR_WORST = ("Highly irregular", "Highly irregular/spiculated outer contour")

_MARGIN_GRADES = (
    (0.6, "Sharp transition", "Abrupt tumor-brain intensity transition, well-demarcated margin"),
    (0.4, "Moderate transition", "Moderately distinct margin with some gradual transition zones"),
    (0.2, "Gradual transition", "Indistinct margin with gradual intensity blending into brain"),
)
_MARGIN_WORST = ("Infiltrative transition", "No clear intensity demarcation, tumor infiltrates surrounding parenchyma")


def analyze_border_regularity(mask, voxel_dims):
    """Contour smoothness from the variation of |grad(signed distance)| over the surface voxels (reference :133-205)."""
    from .. import voxelops as V
    from .utils import mask_u8

    m = mask_u8(mask)
    if int(m.count_nonzero()) == 0:
        return {"regularity_score": 0, "classification": "No tumor", "description": "No tumor detected"}
    inside = V.distance_transform_edt(m)
    outside = V.distance_transform_edt((m == 0).to(m.dtype))
    count, mean, std = V.surface_gradient_stats(m, inside, outside)
    if count < 10:
        return {"regularity_score": 1.0, "classification": "Too small to assess",
                "description": "Tumor too small for border analysis"}
    regularity = 1.0 / (1.0 + std / mean) if std > 0 else 1.0
    name, text = _grade(regularity, _CONTOUR_GRADES, _CONTOUR_WORST)
    return {"regularity_score": float(regularity), "classification": name, "description": text,
            "surface_voxel_count": int(count), "concept": "contour_smoothness"}


def analyze_margin_definition(t1ce_data, seg_data, tumor_masks, voxel_dims):
    """Sharpness of the tumour-brain intensity transition on T1ce (reference :208-290)."""
    from .. import voxelops as V
    from .utils import mask_u8

    wt = mask_u8(tumor_masks["wt"])
    if int(wt.count_nonzero()) == 0:
        return {"margin_sharpness": 0, "classification": "No tumor", "description": "No tumor detected"}
    t1ce = V.as_intensity(t1ce_data)
    peritumoral = V.mask_andnot(V.binary_dilation(wt, iterations=5), wt)
    _, tumor_mean, _, _, _ = V.intensity_moments(t1ce, wt)
    n_peri, peri_mean, _, _, _ = V.intensity_moments(t1ce, peritumoral)
    if n_peri == 0:
        return {"margin_sharpness": 0.5, "classification": "Could not assess",
                "description": "Insufficient peritumoral tissue for analysis"}
    contrast = abs(tumor_mean - peri_mean) / peri_mean if peri_mean > 0 else 0
    n_in, in_mean, in_std, _, _ = V.intensity_moments(t1ce, V.mask_andnot(wt, V.binary_erosion(wt)))
    n_out, out_mean, out_std, _, _ = V.intensity_moments(t1ce, V.mask_andnot(V.binary_dilation(wt), wt))
    step = abs(in_mean - out_mean) / (in_std + out_std + 1e-6) if n_in > 0 and n_out > 0 else 0
    sharpness = min(1.0, (contrast + step) / 2)
    name, text = _grade(sharpness, _MARGIN_GRADES, _MARGIN_WORST)
    return {"margin_sharpness": float(sharpness), "contrast_ratio": float(contrast), "border_gradient": float(step),
            "classification": name, "description": text, "concept": "intensity_transition"}


def analyze_cystic_vs_solid(t1_data, t2_data, flair_data, seg_data, tumor_masks, voxel_dims):
    """Cystic / solid / mixed classification from CSF-like signal inside the necrotic core (reference :293-397)."""
    from .. import voxelops as V
    from .utils import mask_u8

    ncr_count, wt_count = tumor_masks["ncr"].sum(), tumor_masks["wt"].sum()
    if wt_count == 0:
        return {"classification": "No tumor", "cystic_percentage": 0, "solid_percentage": 0,
                "description": "No tumor detected"}
    voxel_vol = np.prod(voxel_dims) / 1000
    t1, t2, flair = V.as_intensity(t1_data), V.as_intensity(t2_data), V.as_intensity(flair_data)
    # CSF reference levels: bottom 10 % of T1, top 15 % of T2, bottom 20 % of FLAIR (over the non-zero voxels)
    t1_upper = V.MaskedValues(t1).percentiles([10])[0]
    t2_lower = V.MaskedValues(t2).percentiles([85])[0]
    flair_upper = V.MaskedValues(flair).percentiles([20])[0]
    cystic_fraction, t2_cv, flair_t2_ratio = 0, 0, 1
    if ncr_count > 0:
        ncr = mask_u8(tumor_masks["ncr"])
        csf_like = V.masked_threshold_count(ncr, t1, t1_upper * 1.5, t2, t2_lower * 0.8, flair, flair_upper * 2)
        cystic_fraction = csf_like / ncr_count
        _, t2_mean, t2_std, _, _ = V.intensity_moments(t2, ncr)
        _, flair_mean, _, _, _ = V.intensity_moments(flair, ncr)
        if t2_mean > 0:
            t2_cv, flair_t2_ratio = t2_std / t2_mean, flair_mean / t2_mean
    wt_volume = wt_count * voxel_vol
    cystic_volume = ncr_count * voxel_vol * cystic_fraction
    cystic_pct = (cystic_volume / wt_volume * 100) if wt_volume > 0 else 0
    if cystic_pct > 70:
        name, text = "Predominantly cystic", "Large cystic component with thin wall/rim"
    elif cystic_pct > 40:
        name, text = "Cystic with solid component", "Mixed cystic and solid tumor with significant cystic component"
    elif cystic_pct > 15:
        name, text = "Solid with cystic component", "Predominantly solid tumor with cystic/necrotic areas"
    elif ncr_count > 0 and t2_cv > 0.3:
        name, text = "Solid with necrosis", "Solid tumor with central necrotic (non-cystic) component"
    elif ncr_count > 0:
        name, text = "Solid with possible cyst", "Solid tumor with possible small cystic component"
    else:
        name, text = "Solid", "Homogeneous solid tumor without significant cystic component"
    signal = {
        "t2_homogeneity": "Homogeneous" if t2_cv < 0.2 else ("Mildly heterogeneous" if t2_cv < 0.4 else "Heterogeneous"),
        "flair_suppression": "Present (suggests true cyst)" if flair_t2_ratio < 0.7
        else "Absent (suggests necrosis/protein)",
        "csf_like_signal_fraction": float(cystic_fraction),
    }
    return {"classification": name, "cystic_volume_cm3": float(cystic_volume), "cystic_percentage": float(cystic_pct),
            "solid_volume_cm3": float(wt_volume - cystic_volume), "solid_percentage": float(100 - cystic_pct),
            "signal_characteristics": signal, "description": text}


def analyze_morphology(input_folder, segmentation_path, output_path=None):
    """File-level driver (reference :602-687): the four modalities and the label file in, the step-4 result dictionary
    (and JSON file) out.  The narrative `text_summary` of the reference is report generation and is not produced."""
    from . import utils as U

    case_id = U.get_case_id(input_folder)
    paths = U.get_mri_paths(input_folder, case_id)
    t1, _, t1_header = U.load_nifti(paths["t1"])
    volumes = {"t1": t1}
    for mod in ("t1ce", "t2", "flair"):
        volumes[mod] = U.load_nifti(paths[mod])[0]
    seg, _, _ = U.load_nifti(segmentation_path)
    voxel_info = U.get_voxel_dimensions(t1_header)
    dims = [float(v) for v in voxel_info["dimensions_mm"]]
    lv = U.LabelVolume(seg)
    masks = U.get_tumor_masks(lv)
    results = {
        "case_id": case_id,
        "step": "Step 4 - Tumor morphology and margins",
        "voxel_info": voxel_info,
        "shape_descriptors": calculate_shape_descriptors(lv, masks, dims),
        "border_regularity": analyze_border_regularity(masks["wt"], dims),
        "margin_definition": analyze_margin_definition(volumes["t1ce"], lv, masks, dims),
        "necrosis_pattern": analyze_necrosis_pattern(lv, masks, np.array(dims)),
        "cystic_solid_classification": analyze_cystic_vs_solid(volumes["t1"], volumes["t2"], volumes["flair"], lv, masks,
                                                               dims),
    }
    if output_path:
        U.save_results(results, output_path)
    return results


def main(argv=None):
    """`python -m brainseg_b200.feature_extraction.step4_morphology --input DIR --segmentation FILE [--output JSON]`
    (the reference script's command line, :690-703)."""
    import argparse

    cli = argparse.ArgumentParser(description="Step 4: Analyze tumor morphology and margins")
    cli.add_argument("--input", required=True, help="Input folder containing MRI sequences")
    cli.add_argument("--segmentation", required=True, help="Path to segmentation mask (NIfTI)")
    cli.add_argument("--output", default=None, help="Output path for JSON results")
    args = cli.parse_args(argv)
    return analyze_morphology(args.input, args.segmentation, args.output)


if __name__ == "__main__":
    main()
