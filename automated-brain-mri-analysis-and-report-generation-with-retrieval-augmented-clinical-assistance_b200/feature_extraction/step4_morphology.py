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
