"""Short profiling target: one batch-8 forward of each benchmark model (all conv / norm launches) plus the gather and
head kernels of one tile — small enough to sit under `ncu` (a whole case is ~7000 launches).

  python scripts/profile_forward.py            # plain run (must exit 0 before any ncu run of the same command)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import synthetic_case as SY  # noqa: E402
from brainseg_b200 import sliding  # noqa: E402

PATCH = (128, 128, 128)


def main():
    torch.cuda.set_device(0)
    m1, m2 = SY.build_benchmark_models("large")
    vol = torch.from_numpy(SY.case_volume(0, (4, 128, 128, 136))).cuda()  # 2 tiles along x
    for net in (m1, m2):
        pred = sliding.SlidingWindowPredictor(net.engines_for(PATCH, 8, 1), 0.5, True, sliding.ALL_MIRROR_CODES,
                                              net._nonlin_name())
        for _ in range(2):  # second pass = warm
            acc = pred.accumulate(vol)
        seg, _ = pred.finalize([acc], tuple(vol.shape[1:]), (1, 2, 3), want_probs=False)
        torch.cuda.synchronize()
        print(type(net).__name__, "labels", torch.bincount(seg.flatten().long(), minlength=4).tolist(), flush=True)


if __name__ == "__main__":
    main()
