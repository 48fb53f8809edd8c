#!/usr/bin/env python
"""BASELINE configs[1] IN FULL through the CPU oracle -> tests/golden/config2_oracle.npz  (TEST INFRASTRUCTURE).

The benchmarked configuration — synthetic case seed 0 (4x155x240x240), model 1 (BatchNorm, 31.2 M) and model 2
(GroupNorm large, 87.4 M), 18 tiles x 8 mirrors each = 288 fp32 forwards, Gaussian weighting, regions decision — is
too slow for a routine test (about 25 minutes on 8 host cores), so it is run ONCE by this script and its result
committed:

    seg1, seg2       oracle label volumes of the two models (uint8, nnU-Net convention), 2 bits per voxel, packed
    decisive1/2      bit masks: every class probability further than 1e-2 from the 0.5 threshold
    probs1/2         class probabilities on the lattice [:, ::4, ::4, ::4] (float32)

tests/test_gpu_config2.py compares the sm_100a path with it (label agreement >= 99.9 %, equality on decisive voxels,
probabilities within 1e-2 on the lattice); bench.py's `result_check` reports the agreement of the benchmarked run's
label volumes with it.  Follows run_brats2021_inference_singlethread.py:97-106 (per-model predict), :144-156
(regions export); the ensemble (:305) and remap are re-derived from seg1/seg2 by the consumers with oracle/postproc.py.

    python oracle/make_config2_golden.py [--threads N] [--out tests/golden/config2_oracle.npz]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

PATCH = (128, 128, 128)
LATTICE = 4
TOL = 1e-2


from synthetic_case import load_config2_oracle as load, pack2, unpack2  # noqa: E402,F401


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "config2_oracle.npz"))
    args = ap.parse_args()
    if args.threads:
        torch.set_num_threads(args.threads)
    from oracle import sliding_window as SW
    from oracle import synthetic as SY
    from tests.helpers import build_dropin_unet, oracle_fns

    vol = SY.case_volume(0, (4, 155, 240, 240))
    models = [build_dropin_unet("bn", base=32, num_pool=5, seed=1),
              build_dropin_unet("gn", base=32, num_pool=5, seed=2, groups=8, encoder_scale=2, max_num_features=512)]
    out = {"shape": np.array(vol.shape[1:]), "lattice": np.array(LATTICE), "tol": np.array(TOL)}
    t0 = time.time()
    for m, net in enumerate(models, 1):
        fwd, _, _ = oracle_fns(net)
        done = [0]

        def hook(*_):
            done[0] += 1
            print(f"[{time.time() - t0:7.0f}s] model {m}: tile {done[0]}/18", flush=True)

        seg, probs = SW.predict_3d_tiled(fwd, torch.sigmoid, vol, 3, PATCH, True, (0, 1, 2), 0.5, True, (1, 2, 3),
                                         tile_hook=hook)
        out[f"seg{m}"] = pack2(seg.astype(np.uint8))
        out[f"decisive{m}"] = np.packbits(np.all(np.abs(probs - 0.5) > TOL, axis=0).reshape(-1))
        out[f"probs{m}"] = np.ascontiguousarray(probs[:, ::LATTICE, ::LATTICE, ::LATTICE])
        assert np.array_equal(unpack2(out[f"seg{m}"], seg.shape), seg.astype(np.uint8))
    np.savez_compressed(args.out, **out)
    print(f"wrote {args.out} ({os.path.getsize(args.out) / 1e6:.1f} MB) in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
