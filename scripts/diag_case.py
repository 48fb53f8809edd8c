"""Stage-by-stage run of one benchmark case with a device sync after every launch group: finds hangs and prints a
per-layer time / TFLOP/s table.  Run on the GPU box: python scripts/diag_case.py [standard|large] [batch]
"""
import faulthandler
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.enable()
faulthandler.dump_traceback_later(120, repeat=True)

import numpy as np  # noqa: E402
import torch  # noqa: E402

T0 = time.time()


def log(msg):
    print(f"[{time.time() - T0:7.1f}s] {msg}", flush=True)


def time_steps(eng, reps=3):
    """Per-step device time (ms, best of reps) with a sync around every step."""
    out = []
    for i, step in enumerate(eng.steps):
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.sub_events = []
            e0.record()
            step(None)
            e1.record()
            torch.cuda.synchronize()
            if e0.elapsed_time(e1) < best:
                best = e0.elapsed_time(e1)
                split = f" [conv {e0.elapsed_time(eng.sub_events[0]):.3f} + norm {eng.sub_events[0].elapsed_time(e1):.3f}]" \
                    if eng.sub_events else ""
            eng.sub_events = None
        out.append(best)
        info = eng.step_info[i] if hasattr(eng, "step_info") else {}
        fl = info.get("flops", 0.0)
        log(f"  step {i:3d} {info.get('name', ''):40s} {best:8.3f} ms  {fl / best / 1e9 if best > 0 else 0:8.1f} TFLOP/s  "
            f"{info.get('plan', '')}{split}")
    return out


def main():
    model2 = sys.argv[1] if len(sys.argv) > 1 else "large"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    import bench as B
    from brainseg_b200 import pipeline as PL
    from oracle import synthetic as SY

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    import synthetic_case
    m1, m2 = synthetic_case.build_benchmark_models(model2)
    log("models built")
    pipe = PL.BratsCasePipeline([m1, m2], B.PATCH, 0.5, (0, 1, 2), True, True, (1, 2, 3), "brats2025", batch=batch)
    torch.cuda.synchronize()
    log(f"pipeline ready; mem {torch.cuda.memory_allocated() / 2**30:.2f} GiB")
    for k, pred in enumerate(pipe.predictors):
        eng = pred.engine
        log(f"engine {k}: {len(eng.steps)} steps, {eng.flops / 1e9:.1f} GF per batch of {eng.batch}")
        ts = time_steps(eng)
        tot = sum(ts)
        log(f"engine {k}: sum of steps {tot:.2f} ms -> {eng.flops / tot / 1e9:.1f} TFLOP/s")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run()
        e1.record()
        torch.cuda.synchronize()
        log(f"engine {k}: back-to-back run {e0.elapsed_time(e1):.2f} ms -> {eng.flops / e0.elapsed_time(e1) / 1e9:.1f} TFLOP/s")

    vol = torch.from_numpy(SY.case_volume(0, B.VOL_SHAPE)).to(dev)
    gt = torch.from_numpy(np.ascontiguousarray(SY.label_volume(0, (155, 240, 240)))).to(dev)
    torch.cuda.synchronize()
    log("inputs on device")
    segs = []
    for k, pred in enumerate(pipe.predictors):
        t = time.time()
        acc = pred.accumulate(vol)
        torch.cuda.synchronize()
        log(f"model {k}: accumulate {time.time() - t:.3f} s")
        t = time.time()
        seg, _ = pred.finalize([acc], tuple(vol.shape[1:]), (1, 2, 3), want_probs=False)
        torch.cuda.synchronize()
        log(f"model {k}: finalize {(time.time() - t) * 1e3:.2f} ms; label histogram {torch.bincount(seg.flatten().long(), minlength=4).tolist()}")
        segs.append(seg)
    from brainseg_b200 import voxelops as V
    from brainseg_b200 import convert_labels_to_brats as CL
    from brainseg_b200 import evaluate_segmentation as EV
    from brainseg_b200.feature_extraction import step3_multiplicity as S3
    from brainseg_b200.feature_extraction import step4_morphology as S4
    from brainseg_b200.feature_extraction import utils as FU

    t = time.time()
    brats = V.ensemble_round(segs[0], segs[1], post_lut=CL.LUT_BRATS2025)
    torch.cuda.synchronize()
    log(f"ensemble+remap {(time.time() - t) * 1e3:.2f} ms")
    t = time.time()
    ev = EV.evaluate_arrays(brats, gt)
    log(f"evaluate {(time.time() - t) * 1e3:.2f} ms mean dice {ev['mean_dice']}")
    lv = FU.LabelVolume(brats)
    t = time.time()
    _, n, _ = V.ccl26(brats, want_labels=False)
    log(f"ccl26 raw {(time.time() - t) * 1e3:.2f} ms, {n} components")
    t = time.time()
    comp = S3.detect_connected_components(lv, (1.0, 1.0, 1.0))
    log(f"detect_connected_components {(time.time() - t) * 1e3:.2f} ms: {comp['num_components']} significant, "
        f"{comp['excluded_fragments']} fragments")
    t = time.time()
    enh = S3.analyze_enhancing_components(lv, (1.0, 1.0, 1.0))
    log(f"analyze_enhancing_components {(time.time() - t) * 1e3:.2f} ms: {enh['num_enhancing_foci']} foci")
    t = time.time()
    masks = FU.get_tumor_masks(lv)
    S4.calculate_shape_descriptors(lv, masks, (1.0, 1.0, 1.0))
    S4.analyze_necrosis_pattern(lv, masks, np.array((1.0, 1.0, 1.0)))
    log(f"morphology {(time.time() - t) * 1e3:.2f} ms")
    for i in range(2):
        t = time.time()
        out = pipe.run_case(vol, gt=gt)
        torch.cuda.synchronize()
        log(f"run_case {i}: {time.time() - t:.3f} s")
    for lanes in (1, 2):
        pl = PL.BratsCasePipeline([m1, m2], B.PATCH, 0.5, (0, 1, 2), True, True, (1, 2, 3), "brats2025", batch=batch,
                                  lanes=lanes)
        pl.run_case(vol, gt=gt, features=False)
        torch.cuda.synchronize()
        for i in range(2):
            t = time.time()
            segs = pl.segment(vol)
            torch.cuda.synchronize()
            log(f"lanes={lanes}: segment (both models) {time.time() - t:.3f} s")
        del pl
        m1.invalidate_engines()
        m2.invalidate_engines()
        torch.cuda.empty_cache()
    log("done")


if __name__ == "__main__":
    main()
