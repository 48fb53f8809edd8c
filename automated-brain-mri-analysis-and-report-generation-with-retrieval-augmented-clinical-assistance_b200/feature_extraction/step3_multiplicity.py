"""Drop-in for the connected-component functions of the reference's feature_extraction/step3_multiplicity.py.

`scipy.ndimage.label` (26-connected) plus the O(components x volume) per-component NumPy loop (:63-121) become one
device labelling + statistics call (bsg_ccl26_stats: shared-memory union-find, SciPy raster-order numbering,
warp-aggregated per-component sums).  The dictionaries returned carry the same keys and values as the reference's.
"""
import contextlib
import gc

import numpy as np

from .. import voxelops as V
from .utils import LabelVolume

MIN_LESION_VOLUME_CM3 = 0.1  # step3_multiplicity.py:38


@contextlib.contextmanager
def _gc_paused():
    """Building 1e5 small dicts triggers the cyclic collector hundreds of times for nothing (2.5x slower)."""
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


def _volume(seg_data):
    return seg_data if isinstance(seg_data, LabelVolume) else LabelVolume(seg_data)


def detect_connected_components(seg_data, voxel_dims):
    """Separate tumour components of `seg_data > 0` with their properties (reference :41-152)."""
    lv = _volume(seg_data)
    _, num_components, st = V.ccl26(lv.vol, V.MASK_GT0, want_labels=False)
    if num_components == 0:
        return {"num_components": 0, "components": [], "is_single_lesion": True, "description": "No tumor detected"}
    vox = np.prod(voxel_dims)
    # the reference builds a dict for every component and then splits at 0.1 cm³ (:63-125); only the significant ones
    # are returned, so the split / stable volume sort run on arrays and dicts are built for the survivors only
    counts = st["count"].astype(np.int64)
    volumes = counts * vox / 1000
    keep = np.flatnonzero(volumes >= MIN_LESION_VOLUME_CM3)
    n_noise = num_components - len(keep)
    keep = keep[np.argsort(-volumes[keep], kind="stable")]  # list.sort(reverse=True) is stable too (:128)
    significant = []
    for rank, i in enumerate(keep.tolist()):
        r = st[i]
        count = int(r["count"])
        centroid = {"x": float(int(r["s0"]) / count), "y": float(int(r["s1"]) / count),
                    "z": float(int(r["s2"]) / count)}
        bbox = {"x_min": int(r["mn0"]), "x_max": int(r["mx0"]), "y_min": int(r["mn1"]), "y_max": int(r["mx1"]),
                "z_min": int(r["mn2"]), "z_max": int(r["mx2"])}
        composition = {"ncr": int(r["n1"]), "ed": int(r["n2"]), "et": int(r["n3"])}
        significant.append({
            "id": i + 1,
            "voxel_count": count,
            "volume_cm3": float(volumes[i]),
            "centroid_voxel": centroid,
            "centroid_mm": {k: centroid[k] * voxel_dims[a] for a, k in enumerate("xyz")},
            "bounding_box": bbox,
            "max_diameter_mm": float(max((bbox[f"{k}_max"] - bbox[f"{k}_min"]) * voxel_dims[a]
                                         for a, k in enumerate("xyz"))),
            "composition": composition,
            "has_enhancement": composition["et"] > 0,
            "rank": rank + 1,
            "classification": "Primary lesion" if rank == 0 else f"Secondary lesion #{rank}",
        })
    n_sig = len(significant)
    note = f" ({n_noise} sub-threshold fragments excluded, <{MIN_LESION_VOLUME_CM3} cm³)" if n_noise else ""
    return {"num_components": n_sig, "components": significant, "is_single_lesion": n_sig == 1,
            "description": f"{n_sig} lesion(s) detected{note}", "excluded_fragments": n_noise,
            "minimum_volume_threshold_cm3": MIN_LESION_VOLUME_CM3}


def analyze_enhancing_components(seg_data, voxel_dims):
    """Components of the enhancing tumour `seg_data == 3` (reference :207-263)."""
    lv = _volume(seg_data)
    _, n_et, st = V.ccl26(lv.vol, V.bits_of(3), want_labels=False)
    if n_et == 0:
        return {"num_enhancing_foci": 0, "enhancing_components": [], "pattern": "Non-enhancing",
                "description": "No enhancing tumor components detected"}
    vox = np.prod(voxel_dims)
    counts = st["count"].astype(np.int64)
    volumes = counts * vox / 1000
    # exact: the coordinate sums are < 2**53, so float64(sum) / count == np.mean(coords)
    cents = [st[f"s{a}"].astype(np.float64) / counts * voxel_dims[a] for a in range(3)]
    order = np.argsort(-volumes, kind="stable")  # list.sort(reverse=True) is stable too (:243)
    # arrays are permuted once and walked with zip: random-weight nets can produce 1e5 single-voxel foci
    vols_sorted = volumes[order].tolist()
    with _gc_paused():
        comps = [{"id": i, "volume_cm3": v, "centroid_mm": {"x": x, "y": y, "z": z}}
                 for i, v, x, y, z in zip((order + 1).tolist(), vols_sorted, cents[0][order].tolist(),
                                          cents[1][order].tolist(), cents[2][order].tolist())]
    if n_et == 1:
        pattern = "Single enhancing focus"
    elif n_et <= 3:
        pattern = "Few enhancing foci"
    else:
        pattern = "Multiple/scattered enhancing foci"
    return {"num_enhancing_foci": n_et, "enhancing_components": comps, "pattern": pattern,
            "total_enhancing_volume_cm3": float(sum(vols_sorted)),  # same left-to-right order as the reference
            "description": f"{n_et} separate enhancing focus/foci detected"}


def label_components(seg_data, maskbits=V.MASK_GT0):
    """scipy.ndimage.label(mask, generate_binary_structure(3,3)) on the device: (int32 labels, count)."""
    lv = _volume(seg_data)
    labels, n, _ = V.ccl26(lv.vol, maskbits, stats_cap=0)
    return labels, n


# ------------------------------------------------------------------------------------------------------------------
# Lesion-level bookkeeping on the component list (reference :155-205, :266-375) and the file-level driver (:445-546).
# Host arithmetic over a handful of centroids; kept so that `analyze_multiplicity` returns the reference's sections.
# ------------------------------------------------------------------------------------------------------------------
SATELLITE_DISTANCE_MM = 20  # step3_multiplicity.py:34
SEPARATE_DISTANCE_MM = 40   # step3_multiplicity.py:35


def _centroid_distance(a, b):
    return np.sqrt((a["x"] - b["x"]) ** 2 + (a["y"] - b["y"]) ** 2 + (a["z"] - b["z"]) ** 2)


def classify_distance_relationship(distance_mm):
    if distance_mm < SATELLITE_DISTANCE_MM:
        return "Satellite/adjacent"
    return "Regional spread" if distance_mm < SEPARATE_DISTANCE_MM else "Distant/separate"


def calculate_inter_lesion_distances(components, voxel_dims):
    """Pairwise centroid distances (mm) between the significant lesions."""
    pairs = []
    for i, first in enumerate(components):
        for second in components[i + 1:]:
            d = _centroid_distance(first["centroid_mm"], second["centroid_mm"])
            pairs.append({"component_1": first["id"], "component_2": second["id"], "distance_mm": float(d),
                          "relationship": classify_distance_relationship(d)})
    values = [p["distance_mm"] for p in pairs]
    if not values:
        return {"distances": [], "min_distance_mm": None, "max_distance_mm": None, "mean_distance_mm": None}
    return {"distances": pairs, "min_distance_mm": float(min(values)), "max_distance_mm": float(max(values)),
            "mean_distance_mm": float(np.mean(values))}


def detect_satellite_lesions(components, primary_component, voxel_dims):
    """Secondary lesions whose centroid lies within SATELLITE_DISTANCE_MM of the primary's."""
    if len(components) < 2:
        return {"satellite_count": 0, "satellites": [], "has_satellites": False,
                "description": "Single lesion, no satellites"}
    found = []
    for comp in components[1:]:
        d = _centroid_distance(primary_component["centroid_mm"], comp["centroid_mm"])
        if d < SATELLITE_DISTANCE_MM:
            found.append({"component_id": comp["id"], "volume_cm3": comp["volume_cm3"],
                          "distance_from_primary_mm": float(d), "has_enhancement": comp["has_enhancement"]})
    text = (f"{len(found)} satellite lesion(s) within {SATELLITE_DISTANCE_MM}mm of primary tumor" if found
            else "No satellite lesions detected")
    return {"satellite_count": len(found), "satellites": found, "has_satellites": bool(found),
            "satellite_threshold_mm": SATELLITE_DISTANCE_MM, "description": text}


_PATTERNS = {  # pattern -> (classification, clinical implication, differential considerations)
    "Solitary": ("Single contiguous lesion", "Unifocal disease, typical for primary brain tumor",
                 ["Primary glioma", "Solitary metastasis", "Lymphoma", "Abscess"]),
    "Primary with satellites": ("Main lesion with satellite nodules",
                                "Suggests local tumor spread or infiltrative growth pattern",
                                ["High-grade glioma with infiltration", "Multicentric glioma", "Inflammatory process"]),
    "Regional multifocal": ("Few lesions in regional distribution",
                            "Regional disease, may be contiguous or multicentric",
                            ["Multicentric glioma", "Regional metastases", "Demyelinating disease"]),
    "Distant multifocal": ("Separate lesions in different brain regions",
                           "Multifocal disease, consider metastatic process",
                           ["Metastatic disease", "Multicentric glioma", "CNS lymphoma", "Multifocal infection"]),
    "Diffuse/scattered": ("Multiple lesions throughout brain",
                          "Diffuse disease pattern, high probability of metastatic or systemic process",
                          ["Metastatic carcinoma", "CNS lymphoma", "Miliary tuberculosis", "Septic emboli"]),
}


def classify_distribution_pattern(component_analysis, distance_analysis, satellite_analysis, enhancing_analysis):
    n = component_analysis["num_components"]
    if n == 0:
        return {"pattern": "No tumor", "classification": "No lesion detected", "clinical_implication": "N/A",
                "differential_considerations": []}
    if n == 1:
        pattern = "Solitary"
    elif satellite_analysis["has_satellites"]:
        pattern = "Primary with satellites"
    elif n <= 3:
        far = distance_analysis["max_distance_mm"]
        pattern = "Regional multifocal" if far and far < SEPARATE_DISTANCE_MM else "Distant multifocal"
    else:
        pattern = "Diffuse/scattered"
    foci = enhancing_analysis["num_enhancing_foci"]
    if foci == 0:
        note = "Non-enhancing pattern may suggest low-grade pathology"
    elif foci > n:
        note = "Multiple enhancing foci within lesions suggest heterogeneous enhancement"
    else:
        note = "Enhancement pattern consistent with lesion count"
    classification, implication, differentials = _PATTERNS[pattern]
    return {"pattern": pattern, "classification": classification, "clinical_implication": implication,
            "differential_considerations": list(differentials), "enhancement_note": note, "lesion_count": n,
            "enhancing_foci_count": foci}


def analyze_multiplicity(input_folder, segmentation_path, output_path=None):
    """File-level driver (reference :445-546): NIfTI in, the step-3 result dictionary (and JSON file) out.  The
    narrative `text_summary` of the reference is report generation and is not produced."""
    from . import utils as U

    case_id = U.get_case_id(input_folder)
    _, _, t1_header = U.load_nifti(U.get_mri_paths(input_folder, case_id)["t1"])
    seg, _, _ = U.load_nifti(segmentation_path)
    voxel_info = U.get_voxel_dimensions(t1_header)
    dims = [float(v) for v in voxel_info["dimensions_mm"]]  # float64 from here on (see utils.NiftiHeaderView)
    lv = LabelVolume(seg)  # np.round(seg).astype(int) on the device
    components = detect_connected_components(lv, dims)
    lesions = components["components"]
    distances = calculate_inter_lesion_distances(lesions, dims)
    if lesions:
        satellites = detect_satellite_lesions(lesions, lesions[0], dims)
    else:
        satellites = {"satellite_count": 0, "satellites": [], "has_satellites": False, "description": "No tumor detected"}
    enhancing = analyze_enhancing_components(lv, dims)
    results = {"case_id": case_id, "step": "Step 3 - Lesion multiplicity and distribution", "voxel_info": voxel_info,
               "component_analysis": components, "distance_analysis": distances, "satellite_analysis": satellites,
               "enhancing_analysis": enhancing,
               "distribution_pattern": classify_distribution_pattern(components, distances, satellites, enhancing)}
    if output_path:
        U.save_results(results, output_path)
    return results


def main(argv=None):
    """`python -m brainseg_b200.feature_extraction.step3_multiplicity --input DIR --segmentation FILE [--output JSON]`
    (the reference script's command line, :549-562)."""
    import argparse

    cli = argparse.ArgumentParser(description="Step 3: Analyze lesion multiplicity and distribution")
    cli.add_argument("--input", required=True, help="Input folder containing MRI sequences")
    cli.add_argument("--segmentation", required=True, help="Path to segmentation mask (NIfTI)")
    cli.add_argument("--output", default=None, help="Output path for JSON results")
    args = cli.parse_args(argv)
    return analyze_multiplicity(args.input, args.segmentation, args.output)


if __name__ == "__main__":
    main()
