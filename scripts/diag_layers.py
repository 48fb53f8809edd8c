"""Per-layer device time / TFLOP/s table of both benchmark networks (CUDA events around every engine step, best of 3)
and the back-to-back forward time.  Environment switches of the engine apply (BSG_FUSE_NORM, BSG_OVERFLOW_GUARD,
BSG_ACT_DTYPE).  Run on the GPU box: python scripts/diag_layers.py [per-engine batch, default 4]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import synthetic_case as SY  # noqa: E402
from scripts.diag_case import log, time_steps  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    torch.cuda.set_device(0)
    log(f"BSG_FUSE_NORM={os.environ.get('BSG_FUSE_NORM', '1')} BSG_OVERFLOW_GUARD={os.environ.get('BSG_OVERFLOW_GUARD', '1')} "
        f"BSG_KWPACK={os.environ.get('BSG_KWPACK', '1')} batch {batch}")
    for k, net in enumerate(SY.build_benchmark_models("large")):
        eng = net.engine_for((128, 128, 128), batch)
        log(f"engine {k}: {len(eng.steps)} steps, {eng.flops_algo / 1e9:.1f} GF (algorithmic) per batch of {eng.batch}, "
            f"{eng.fused_norms} norm passes handed to the consumer")
        for _ in range(2):
            eng.run()
        torch.cuda.synchronize()
        ts = time_steps(eng)
        log(f"engine {k}: sum of steps {sum(ts):.2f} ms")
        best = 1e30
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.run()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        log(f"engine {k}: back-to-back run {best:.2f} ms -> {eng.flops_algo / best / 1e9:.1f} TFLOP/s (algorithmic)")
        net.invalidate_engines()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
