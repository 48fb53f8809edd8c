"""Small driver for `compute-sanitizer --tool memcheck`: every non-tensor-core kernel of the library once or twice on
odd-shaped volumes (shapes that are not multiples of any tile / vector width), results checked against NumPy / SciPy.

  python scripts/sanitize_voxelops.py                                   # plain run first
  compute-sanitizer --tool memcheck python scripts/sanitize_voxelops.py

Round 1: the plain run passes on a B200; compute-sanitizer is closed on this GPU pool (runs under it left GPUs needing
a reset), so memory safety rests on these odd-shape comparisons and the parity tests.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scipy import ndimage as ndi

from brainseg_b200 import convert_labels_to_brats as CL
from brainseg_b200 import preprocessing as PP
from brainseg_b200 import voxelops as V
from brainseg_b200.feature_extraction import utils as U


def main():
    rng = np.random.default_rng(0)
    for shape in ((7, 9, 13), (33, 17, 70), (5, 64, 3)):
        lab = (ndi.gaussian_filter(rng.standard_normal(shape), 1.0) > 0.2).astype(np.uint8) * rng.integers(1, 4, shape).astype(np.uint8)
        m = lab > 0
        g = V.as_mask(m)
        vol = torch.from_numpy(lab).cuda()
        assert np.array_equal(V.binary_erosion(g, 2).cpu().numpy().astype(bool), ndi.binary_erosion(m, iterations=2))
        assert np.array_equal(V.binary_dilation(g, 3).cpu().numpy().astype(bool), ndi.binary_dilation(m, iterations=3))
        if not m.all():
            assert np.array_equal(V.distance_transform_edt(g).cpu().numpy(), ndi.distance_transform_edt(m))
        for conn, rank in ((6, 1), (18, 2), (26, 3)):
            got, n = V.ccl(vol, V.MASK_GT0, conn)
            want, wn = ndi.label(m, structure=ndi.generate_binary_structure(3, rank))
            assert n == wn and np.array_equal(got.cpu().numpy(), want)
        _, n, stats = V.ccl26(vol)
        assert int(stats["count"].sum()) == int(m.sum())
        data = np.round(rng.standard_normal(shape) * 100).astype(np.float32)
        d = torch.from_numpy(data).cuda()
        cnt, mean, std, lo, hi = V.intensity_moments(d, g)
        assert cnt == int(m.sum()) and lo == data[m].min() and hi == data[m].max()
        sel = V.MaskedValues(d, g)
        assert sel.percentiles([5, 50, 95]) == [float(np.percentile(data[m].astype(np.float64), q)) for q in (5, 50, 95)]
        assert V.MaskedValues(d).median() == float(np.median(data[data > 0].astype(np.float64)))
        assert V.masked_threshold_count(g, d, 10.0, d, -50.0, None, 0.0) == int(((data[m] < 10) & (data[m] > -50)).sum())
        assert np.array_equal(V.label_lut(vol, CL.LUT_BRATS2025).cpu().numpy(), np.asarray(CL.LUT_BRATS2025, np.uint8)[lab])
        other = np.roll(lab, 1, axis=0)
        ens = V.ensemble_round(vol, torch.from_numpy(other).cuda()).cpu().numpy()
        assert np.array_equal(ens, np.round((lab.astype(np.float64) + other) / 2.0).astype(np.uint8))
        h = V.joint_hist(vol, torch.from_numpy(other).cuda())
        assert h.sum() == lab.size and h[1, 2] == int(((lab == 1) & (other == 2)).sum())
        lv = U.LabelVolume(lab)
        masks = U.get_tumor_masks(lv)
        assert masks["wt"].sum() == int(m.sum())
        assert U.get_bounding_box(masks["wt"]) is not None
    head = rng.standard_normal((4, 19, 23, 21)).astype(np.float32)
    head[:, :3] = 0
    head[:, :, :, 18:] = 0
    out, props = PP.preprocess_case(head)
    assert out.shape[0] == 4 and tuple(out.shape[1:]) == tuple(b - a for a, b in props["crop_bbox"])
    torch.cuda.synchronize()
    print("sanitize_voxelops: all kernels ran, results match")


if __name__ == "__main__":
    main()
