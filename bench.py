#!/usr/bin/env python
"""Benchmark of the hot path: cases/sec for a full BraTS case (BASELINE.json metric).

A "step" is one synthetic case through the whole path: sliding-window inference of 4x155x240x240 with the two-model
ensemble and 8-way mirror TTA (2 x 18 tiles x 8 mirrors = 288 forwards), finalize, label-round ensemble + BraTS
remap, Dice vs a synthetic ground truth, 26-connected components + per-component statistics, morphology moments.

  python bench.py --gpus N --steps K --warmup W            our sm_100a path (one process per GPU under torchrun)
        N = 1: configs[1].  N > 1: `value` = cohort throughput (configs[3]: cases sharded over the ranks, no data-path
        collective, distinct seeded cases) and, in the same line, a `latency` record (configs[2]: the (tile, mirror)
        work items of ONE case sharded over the ranks + the accumulator exchange, strong scaling).
  python bench.py --impl reference --gpus N ...            the reference's CPU implementation (oracle port) on the
                                                           host cores, bounded sample per step
Prints ONE JSON line (rank 0).  The "ours" arm never imports `oracle/`.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "cases/sec: 4x240x240x155 BraTS vol, 2-model+8xTTA, 1/2/4/8 B200 vs host CPU"
VOL_SHAPE = (4, 155, 240, 240)
PATCH = (128, 128, 128)
N_TILES = 18
N_MIRRORS = 8
GF_MODEL1, GF_MODEL2_LARGE = 965.5, 3342.2  # algorithmic GFLOP per 128^3 forward (tests/golden/unet_keys.json)

T0 = time.time()


def log(msg):
    sys.stderr.write(f"[bench {time.time() - T0:7.1f}s] {msg}\n")
    sys.stderr.flush()


def usable_cpus():
    """Host threads this process may really use: the cgroup CPU quota if there is one, else the visible cores."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except (OSError, ValueError):
        pass
    return n


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1414.9))),
                "tflops_burst": float(p.get("bf16_tflops", 1624.1)),
                "hbm": float(p.get("hbm_gbs", 6527.1)), "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm": 6650.0,
            "source": "fallback (B200_PROFILING.md: 1.4 PF sustained / 6.65 TB/s)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed regions."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.path = None, f"/tmp/bsg_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 6:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        if sm:
            load = [s for s in sm if s > 0.5 * max(sm)] or sm  # under load = samples above the idle clock
            out = {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


# =====================================================================================================================
# CPU arm: the reference's algorithm (oracle port: fp32 torch eager + NumPy / SciPy) on the host cores
# =====================================================================================================================
class CpuReference:
    """Bounded, MEASURED samples of the reference CPU path (SURVEY §8d "CPU baseline timing plan"): one real 128^3
    forward of EACH model, the post-processing chain in full, and (reference arm only) BASELINE configs[0] in full.
    A whole configs[1] case (288 forwards, ~10 minutes on 16 cores) is extrapolated from those and labelled so."""

    def __init__(self, model2, threads=None):
        from oracle import unet as OU
        from synthetic_case import build_benchmark_models

        torch.set_num_threads(threads or usable_cpus())
        self.threads = torch.get_num_threads()
        self.OU = OU
        self.nets = build_benchmark_models(model2)
        self.sds = [{k: v.detach().float() for k, v in n.state_dict().items()} for n in self.nets]
        self.archs = [OU.arch_from_module(n) for n in self.nets]
        self.gf = [OU.conv_flops(sd, a, PATCH) / 1e9 for sd, a in zip(self.sds, self.archs)]
        self.x = torch.randn(1, 4, *PATCH, generator=torch.Generator().manual_seed(0))
        OU.forward(self.sds[0], self.archs[0], self.x[:, :, :32, :32, :32])  # warm the thread pool

    def forward_pair(self):
        """One forward of model 1 and one of model 2 on a 128^3 tile + sigmoid: 1/144 of a case's forwards."""
        ts = []
        for sd, arch in zip(self.sds, self.archs):
            t0 = time.perf_counter()
            torch.sigmoid(self.OU.forward(sd, arch, self.x))
            ts.append(time.perf_counter() - t0)
        return ts

    def post_chain(self):
        from oracle import postproc as OP
        from synthetic_case import label_pair

        pred, gt = label_pair(0, (240, 240, 155))
        t0 = time.perf_counter()
        ens = OP.ensemble_labels_round(pred, gt)
        brats = OP.convert_labels_to_brats2025(ens.astype(np.float64))
        OP.evaluate_arrays(brats.astype(np.float64), gt.astype(np.float64))
        seg = np.round(brats).astype(np.int32)
        OP.detect_connected_components(seg, (1.0, 1.0, 1.0))
        OP.analyze_enhancing_components(seg, (1.0, 1.0, 1.0))
        masks = OP.get_tumor_masks(brats)
        OP.calculate_shape_descriptors(brats, masks, (1.0, 1.0, 1.0))
        OP.analyze_necrosis_pattern(brats, masks, np.array((1.0, 1.0, 1.0)))
        return time.perf_counter() - t0

    def config0_full(self):
        """BASELINE configs[0] start to finish: the 4x155x240x240 case, model 1, no mirroring, 18 tiles, Gaussian
        weighting, NumPy accumulation, regions decision (the reference's own CPU-runnable configuration)."""
        from oracle import sliding_window as SW
        from synthetic_case import case_volume

        vol = case_volume(0, VOL_SHAPE)
        sd, arch = self.sds[0], self.archs[0]
        t0 = time.perf_counter()
        SW.predict_3d_tiled(lambda x: self.OU.forward(sd, arch, x), torch.sigmoid, vol, 3, PATCH, False, (0, 1, 2), 0.5,
                            True, (1, 2, 3))
        return time.perf_counter() - t0

    @staticmethod
    def case_seconds(t1, t2, t_post):
        return N_TILES * N_MIRRORS * (t1 + t2) + t_post


def cpu_baseline_for_ours(model2):
    """`cpu_baseline` of the N=1 line: ~15-25 s of host work."""
    ref = CpuReference(model2)
    log(f"cpu baseline: {ref.threads} threads")
    t1, t2 = ref.forward_pair()
    log(f"cpu baseline: model-1 forward {t1:.2f} s, model-2 forward {t2:.2f} s")
    t_post = ref.post_chain()
    log(f"cpu baseline: post-processing chain {t_post:.2f} s")
    t_case = ref.case_seconds(t1, t2, t_post)
    return {"value": 1.0 / t_case, "unit": "cases/s", "cores": ref.threads, "kind": "port",
            "sample": (f"measured: one 128^3 forward of model 1 ({t1:.2f} s, {ref.gf[0]:.0f} GF) and one of model 2 "
                       f"({t2:.2f} s, {ref.gf[1]:.0f} GF), fp32 torch eager, + the post-processing chain in full "
                       f"({t_post:.1f} s); a case = 144 x (both forwards) + post-processing, extrapolated to "
                       f"{t_case:.0f} s"),
            "extrapolated": True}


def run_reference(args):
    """`--impl reference`: every step = one measured forward of model 1 + one of model 2 (1/144 of a case's forwards);
    BASELINE configs[0] and the post-processing chain are measured once in full before the steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(args.model2)
    log(f"reference arm: {ref.threads} host threads")
    t_cfg0 = ref.config0_full() if not args.skip_config0 else None
    if t_cfg0 is not None:
        log(f"configs[0] in full (18 forwards + accumulation, model 1, no TTA): {t_cfg0:.1f} s")
    t_post = ref.post_chain()
    log(f"post-processing chain in full: {t_post:.1f} s")
    pairs, budget_s, t_begin = [], 200.0, time.perf_counter()
    for i in range(args.steps + args.warmup):
        ts = ref.forward_pair()
        if i >= args.warmup:
            pairs.append(ts)
        if time.perf_counter() - t_begin > budget_s and len(pairs) >= 1:
            break
    t1 = sum(p[0] for p in pairs) / len(pairs)
    t2 = sum(p[1] for p in pairs) / len(pairs)
    t_case = ref.case_seconds(t1, t2, t_post)
    value = 1.0 / t_case
    sample = (f"per step (measured): 1 forward of model 1 ({t1:.2f} s) + 1 of model 2 ({t2:.2f} s) on a 128^3 tile, fp32 "
              f"torch eager, {ref.threads} threads = 1/144 of a case's 288 forwards; measured once in full: the "
              f"post-processing chain ({t_post:.1f} s)"
              + (f" and BASELINE configs[0] (one model, 18 tiles, no TTA: {t_cfg0:.1f} s)" if t_cfg0 is not None else "")
              + f"; `value` = 1 / (144 x step + post-processing) = 1 / {t_case:.0f} s, an extrapolation; "
                "`ms_per_step` is the measured step")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "cases/s", "n_gpus": args.gpus,
            "steps": len(pairs), "warmup": args.warmup, "ms_per_step": (t1 + t2) * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "extrapolated": True, "ms_per_case_extrapolated": t_case * 1e3, "steps_per_case": N_TILES * N_MIRRORS,
            "config0_full_s": t_cfg0, "post_chain_s": t_post,
            "cpu_baseline": {"value": value, "unit": "cases/s", "cores": ref.threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "cases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args):
    sweep = ""
    if PATCH != (128, 128, 128) or VOL_SHAPE != (4, 155, 240, 240) or args.step_size != 0.5:
        sweep = (f"configs[4] sweep point: volume {'x'.join(map(str, VOL_SHAPE))}, patch {PATCH[0]}^3, step {args.step_size}, "
                 f"{N_TILES} tiles; otherwise as ")
    world = args.gpus
    if args.mode == "latency":
        par = (f"configs[2]: (tile,mirror) work items of ONE case sharded over {world} GPU(s), accumulator exchange per "
               "model (peer-memory reduce+finalize kernel, or one NCCL all-reduce)")
    elif world > 1:
        par = (f"configs[3]: distinct seeded cases sharded over {world} GPUs, no data-path collective (`value`); the line's "
               "`latency` record is configs[2] (one case sharded over the same GPUs)")
    else:
        par = "one GPU"
    return {"workload": sweep + "configs[1]: one synthetic BraTS case 4x155x240x240 (C,z,y,x), patch 128^3, step 0.5, "
                        "Gaussian weighting, 8-way mirror TTA, 2-model ensemble (model 1: Generic_UNet BN 31.2 M; "
                        f"model 2: GroupNorm {'large 87.4 M (encoder_scale 2, max 512)' if args.model2 == 'large' else 'standard 31.2 M'}), "
                        "one fold per model, regions threshold, label-round ensemble, BraTS-2025 remap, Dice vs "
                        "synthetic GT, 26-conn components + stats, morphology moments",
            "forwards_per_case": 2 * N_TILES * N_MIRRORS, "mode": args.mode, "forwards_in_flight": args.batch,
            "post_labels": args.post_labels,
            "case_stream": "steps are timed as a stream of cases: case i+1's inference is submitted before case i's "
                           "post-processing is collected (all submitted cases finish inside the timed region)",
            "l2_policy": "inputs (143 MB fp32 volume, >=1 GB activations per layer) exceed the 126 MB L2",
            "parallelism": par}


# =====================================================================================================================
# The incumbent: the reference's own CUDA path — PyTorch eager / cuDNN, fp16, channels_last_3d — on the same GPU
# =====================================================================================================================
def incumbent_forward_fn(net, dev):
    """Generic_UNet.forward (generic_UNet.py:423-446, do_ds False) + sigmoid as torch.nn.functional calls in fp16 with
    channels_last_3d tensors: what run_brats2021_inference_singlethread.py:209 (`mixed_precision=True`) runs on a GPU,
    with the weights pre-cast (no per-call autocast casts)."""
    import torch.nn.functional as F
    from torch import nn

    cl = torch.channels_last_3d

    def h(t):
        return t.detach().to(dev, torch.float16)

    def block(blk):
        conv, norm = blk.conv, blk.instnorm
        w = h(conv.weight).contiguous(memory_format=cl)
        b = h(conv.bias) if conv.bias is not None else None
        stride, slope = tuple(conv.stride), float(blk.lrelu.negative_slope)
        if isinstance(norm, nn.BatchNorm3d):
            rm, rv, g, be, eps = h(norm.running_mean), h(norm.running_var), h(norm.weight), h(norm.bias), norm.eps

            def f(x):
                return F.leaky_relu(F.batch_norm(F.conv3d(x, w, b, stride, 1), rm, rv, g, be, False, 0.0, eps), slope, True)
        elif isinstance(norm, nn.GroupNorm):
            g, be, eps, groups = h(norm.weight), h(norm.bias), norm.eps, norm.num_groups

            def f(x):
                return F.leaky_relu(F.group_norm(F.conv3d(x, w, b, stride, 1), groups, g, be, eps), slope, True)
        else:
            g = h(norm.weight) if norm.weight is not None else None
            be = h(norm.bias) if norm.bias is not None else None
            eps = norm.eps

            def f(x):
                return F.leaky_relu(F.instance_norm(F.conv3d(x, w, b, stride, 1), None, None, g, be, True, 0.0, eps), slope, True)
        return f

    num_pool = len(net.tu)
    enc = [[block(b) for b in net.conv_blocks_context[d].blocks] for d in range(num_pool)]
    bott = [block(b) for b in list(net.conv_blocks_context[num_pool][0].blocks) + list(net.conv_blocks_context[num_pool][1].blocks)]
    tus = [h(t.weight) for t in net.tu]
    loc = [[block(b) for b in list(l[0].blocks) + list(l[1].blocks)] for l in net.conv_blocks_localization]
    head = net.seg_outputs[num_pool - 1]
    hw, hb = h(head.weight), (h(head.bias) if head.bias is not None else None)

    def forward(x):
        skips = []
        for stage in enc:
            for f in stage:
                x = f(x)
            skips.append(x)
        for f in bott:
            x = f(x)
        for u in range(num_pool):
            x = F.conv_transpose3d(x, tus[u], None, 2)
            x = torch.cat((x, skips[-(u + 1)]), 1)
            for f in loc[u]:
                x = f(x)
        return torch.sigmoid(F.conv3d(x, hw, hb))

    return forward


def time_incumbent(nets, gflops, dev, batch=8, reps=3):
    out = []
    prev = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    try:
        with torch.no_grad():
            for net, gf in zip(nets, gflops):
                fwd = incumbent_forward_fn(net, dev)
                x = torch.randn(batch, 4, *PATCH, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last_3d)
                for _ in range(2):
                    fwd(x)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    y = fwd(x)
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / reps
                out.append({"ms_per_forward_batch": ms, "batch": batch, "tflops": gf * batch / ms, "finite": bool(torch.isfinite(y).all())})
                del fwd, x, y
                torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark = prev
    return out


# =====================================================================================================================
# HBM-bound passes, each timed alone with CUDA events (L2 evicted between repetitions)
# =====================================================================================================================
def hbm_rooflines(dev, peaks, f16):
    from brainseg_b200 import _lib as L
    from brainseg_b200 import convert_labels_to_brats as CL
    from brainseg_b200 import sliding
    from brainseg_b200 import voxelops as V
    from synthetic_case import label_pair

    lib = L.lib()
    flush_buf = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    sink = torch.zeros((), dtype=torch.int64, device=dev)

    def timed(fn, reps=5):
        fn()
        ts = []
        for _ in range(reps):
            sink.copy_(flush_buf.view(torch.int64).sum())  # READ 256 MB: evicts the 126 MB L2, leaves no dirty lines
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    def ptr(t):
        return C.c_void_p(t.data_ptr())

    rows = []

    def row(name, ms, nbytes, note):
        gbs = nbytes / ms / 1e6
        rows.append({"kernel": name, "ms": ms, "algorithmic_bytes": int(nbytes), "achieved": gbs, "unit": "GB/s",
                     "peak": peaks["hbm"], "frac": gbs / peaks["hbm"], "bytes_per_unit": note})

    Z, Y, X = VOL_SHAPE[1:]
    nv = Z * Y * X
    p = PATCH[0]
    pv = p ** 3
    vol = torch.randn(4, Z, Y, X, device=dev)
    codes = (C.c_int * 8)(*range(8))
    dt = torch.float16 if f16 else torch.bfloat16
    xin = torch.empty(8, p, p, p, 16, dtype=dt, device=dev)
    row("gather_patch_kernel (tile crop + 8 mirror copies, 4 -> 16 ch)",
        timed(lambda: L.check(lib.bsg_gather_patch_tta(ptr(vol), 4, Z, Y, X, 27, 56, 56, p, p, p, codes, 8, ptr(xin), 16, f16, 0,
                                                       L.stream_ptr()))),
        4 * pv * 4 + 8 * pv * 32, "per tile voxel: 4 fp32 read + 8 mirrors x 16 ch x 2 B written")
    del xin
    feat = torch.randn(8, p, p, p, 32, device=dev).to(dt)
    acc = torch.zeros(3, Z, Y, X, device=dev)
    gauss = sliding.gaussian_importance_map((p, p, p), dev)
    hw = (C.c_float * 96)(*np.random.default_rng(0).standard_normal(96).astype(np.float32))
    row("head_tta_accumulate_kernel (8 mirrors, 32 ch, 3 classes)",
        timed(lambda: L.check(lib.bsg_head_tta_accumulate(ptr(feat), f16, 32, 32, p, p, p, codes, 8, 0.125, hw, None, 3, 0,
                                                          ptr(gauss), ptr(acc), Z, Y, X, 27, 56, 56, None, 0.0, L.stream_ptr()))),
        8 * pv * 64 + pv * (24 + 4), "per tile voxel: 8 x 32 ch x 2 B features + RMW of 3 fp32 + 4 B Gaussian")
    del feat
    wsum = torch.rand(Z, Y, X, device=dev) + 0.5
    seg = torch.empty(Z, Y, X, dtype=torch.uint8, device=dev)
    ptrs = (C.c_void_p * 1)(acc.data_ptr())
    order = (C.c_int * 3)(1, 2, 3)
    row("finalize_kernel (regions threshold, labels only)",
        timed(lambda: L.check(lib.bsg_finalize(ptrs, 1, ptr(wsum), 3, nv, 1, order, None, ptr(seg), L.stream_ptr()))),
        nv * 17, "per voxel: 3 fp32 accumulators + 1 fp32 weight sum read, 1 label byte written")
    x16 = torch.randn(4, p, p, p, 64, device=dev).to(torch.float16)
    ss = torch.ones(4, 64, 2, device=dev)
    row("norm_apply_kernel (4 x 128^3 x 64 ch, in place)",
        timed(lambda: L.check(lib.bsg_norm_apply_lrelu(ptr(x16), pv, 4, 64, 64, 0, ptr(ss), 0.01, 1, 1, L.stream_ptr()))),
        2 * x16.numel() * 2, "per element: 2 B read + 2 B written")
    del x16, acc, wsum, vol
    pred, gt = label_pair(0, (240, 240, 155))
    a, b = torch.from_numpy(pred).to(dev), torch.from_numpy(gt).to(dev)
    row("pair_round_kernel (label ensemble + remap)", timed(lambda: V.ensemble_round(a, b, post_lut=CL.LUT_BRATS2025)), 3 * nv,
        "per voxel: 2 label bytes read, 1 written")
    if hasattr(V, "ensemble_remap_hist"):
        row("ensemble_hist_kernel (label ensemble + remap + joint histogram vs GT, one pass)",
            timed(lambda: V.ensemble_remap_hist(a, b, b, CL.LUT_BRATS2025)), 4 * nv,
            "per voxel: 3 label bytes read, 1 written")
    buf = torch.empty(257, dtype=torch.int64, device=dev)
    row("joint_hist_kernel (Dice bins)",
        timed(lambda: L.check(lib.bsg_joint_hist_u8(ptr(a), ptr(b), nv, ptr(buf), C.c_void_p(buf.data_ptr() + 2048),
                                                    L.stream_ptr()))), 2 * nv, "per voxel: 2 label bytes read")
    return rows


# =====================================================================================================================
# Ours
# =====================================================================================================================
def sha16(t):
    return hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()[:16]


def result_check(out, golden, is_default_case):
    """Compares the benchmarked run's label volumes with the oracle's for the SAME case (tests/golden/config2_oracle.npz:
    BASELINE configs[1] in full through the fp32 CPU oracle, recorded once by oracle/make_config2_golden.py and pinned
    against this path by tests/test_gpu_config2.py)."""
    rc = {"num_components": out["components"]["num_components"] if "components" in out else None,
          "mean_dice": float(out["evaluation"]["mean_dice"]) if "evaluation" in out else None,
          "label_sha256_16": sha16(out["segmentation"])}
    if golden is None or not is_default_case:
        rc["oracle"] = "not compared (tests/golden/config2_oracle.npz covers the default configs[1] case only)"
        return rc
    segs = [s.cpu().numpy() for s in out["model_segmentations"]]
    agree = [float((s == golden[f"seg{m}"]).mean()) for m, s in enumerate(segs, 1)]
    decisive_bad = [int((s != golden[f"seg{m}"])[golden[f"decisive{m}"]].sum()) for m, s in enumerate(segs, 1)]
    lut = np.zeros(256, np.uint8)
    lut[:4] = (0, 2, 1, 3)  # BraTS-2025 remap (convert_labels_to_brats.py:34-43)
    final_ref = lut[np.round((golden["seg1"].astype(np.float64) + golden["seg2"]) / 2.0).astype(np.uint8)]
    final = out["segmentation"].cpu().numpy()
    rc.update({"oracle": "tests/golden/config2_oracle.npz (fp32 CPU oracle, same case, 288 forwards)",
               "label_agreement_model1": agree[0], "label_agreement_model2": agree[1],
               "label_agreement_final": float((final == final_ref).mean()), "bar": 0.999,
               "mismatches_on_decisive_voxels": decisive_bad,
               "pass": bool(min(agree) >= 0.999 and max(decisive_bad) == 0)})
    return rc


def run_ours(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: brainseg_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from brainseg_b200 import pipeline as PL
    from brainseg_b200 import voxelops as V
    import synthetic_case as SY

    torch.set_num_threads(max(1, usable_cpus() // max(1, world)))
    log(f"rank {rank}/{world}: building models ({torch.get_num_threads()} host threads)")
    m1, m2 = SY.build_benchmark_models(args.model2)
    sharded_only = args.mode == "latency" and world > 1
    pipe = PL.BratsCasePipeline([m1, m2], PATCH, args.step_size, (0, 1, 2), True, True, (1, 2, 3), "brats2025", batch=args.batch,
                                lanes=args.lanes)
    eng1, eng2 = pipe.predictors[0].engine, pipe.predictors[1].engine
    kinds = sorted({"fp16" if e.f16 else "bf16" for e in (eng1, eng2)})
    act_dtype = kinds[0] if len(kinds) == 1 else "+".join(kinds)  # 16-bit operands, fp32 accumulation in TMEM
    log(f"engines ready: {len(pipe.predictors[0].engines)} lane(s) x batch {eng1.batch}; "
        f"{eng1.launches_per_forward} + {eng2.launches_per_forward} launches per forward batch, "
        f"{eng1.flops_algo_per_item / 1e9:.1f} + {eng2.flops_algo_per_item / 1e9:.1f} GF per tile-mirror (algorithmic)")

    # synthetic inputs: pinned host volumes, seeded per rank and step (cohort: distinct cases), + synthetic GT labels
    default_case = VOL_SHAPE == (4, 155, 240, 240) and PATCH == (128, 128, 128) and args.step_size == 0.5 and args.model2 == "large"
    n_distinct = 1 if world == 1 else max(1, min(args.steps, args.distinct))
    seeds = [rank + world * i for i in range(n_distinct)]
    host_vols = [torch.from_numpy(SY.case_volume(s, VOL_SHAPE)).pin_memory() for s in seeds]
    gt_hosts = [torch.from_numpy(np.ascontiguousarray(SY.label_volume(s, VOL_SHAPE[1:]))).pin_memory() for s in seeds]
    blobby = None
    if args.post_labels == "blobby":  # configs[3] as SURVEY §8d writes it: post-processing on blobby label volumes
        blobby = [torch.from_numpy(np.ascontiguousarray(np.roll(g.numpy(), 3, axis=0))).to(dev) for g in gt_hosts]
    dev_vols = [v.to(dev) for v in host_vols]
    dev_gts = [g.to(dev) for g in gt_hosts]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(p, k, vols, gts, owner_of=None, want_host_labels=False, depth=1):
        """k cases as a stream: the inference of the next `depth` cases is submitted before the post-processing of case
        i is collected, so host-side glue overlaps device work.  Every case is submitted AND finished inside the call.
        owner_of(i): sharded mode — the rank that runs case i's post-processing (depth 2 there: the owner's host thread
        is busy with the post-processing for about as long as a sharded case takes, and every other rank would wait
        for its share of the next case at the exchange)."""
        out = seg_host = None
        pend = [p.submit(vols[j % len(vols)], gts[j % len(gts)]) for j in range(min(depth, k))]
        for i in range(k):
            if i + depth < k:
                j = i + depth
                pend.append(p.submit(vols[j % len(vols)], gts[j % len(gts)]))
            mine = owner_of is None or owner_of(i) == rank
            res = p.finish(pend.pop(0), post=mine, labels=blobby[i % len(blobby)] if (blobby is not None and mine) else None)
            if mine:
                out = res
                if want_host_labels:
                    seg_host = res["segmentation"].cpu()  # D2H of the final label volume
        return out, seg_host

    line_extra = {}
    sampler = None
    t_res = t_e2e = t_single = float("nan")
    launches = 0
    out = None
    if not sharded_only:
        log("inputs ready; warm-up")
        t0 = time.time()
        out, _ = run_steps(pipe, max(args.warmup, 1), host_vols, gt_hosts)
        torch.cuda.synchronize()
        log(f"warm-up ({max(args.warmup, 1)} cases): {time.time() - t0:.2f} s, {out['components']['num_components']} significant "
            f"components, {out['components']['excluded_fragments']} fragments, {out['enhancing']['num_enhancing_foci']} ET foci")
        barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        # ---- kernel-side timing: inputs resident in HBM
        l0 = pipe.kernel_launches()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, _ = run_steps(pipe, args.steps, dev_vols, dev_gts)
        e1.record()
        barrier()
        t_res = e0.elapsed_time(e1) / 1e3
        log(f"resident: {t_res / args.steps:.3f} s per case")
        launches = (pipe.kernel_launches() - l0) + args.steps * pipe.extra_launches
        # ---- end to end: host buffers, H2D + D2H inside the timed region
        barrier()
        e0.record()
        out, seg_host = run_steps(pipe, args.steps, host_vols, gt_hosts, want_host_labels=True)
        e1.record()
        barrier()
        t_e2e = e0.elapsed_time(e1) / 1e3
        log(f"e2e: {t_e2e / args.steps:.3f} s per case")
        # single-case latency (no overlap between consecutive cases), host buffers
        barrier()
        t0 = time.perf_counter()
        out1 = pipe.run_case(host_vols[0], gt=gt_hosts[0])
        out1["segmentation"].cpu()
        torch.cuda.synchronize()
        t_single = time.perf_counter() - t0
        log(f"single case, unpipelined: {t_single:.3f} s")
        if world > 1 or args.post_breakdown:
            out = out1  # seed `rank` case (rank 0: the default case) for result_check

    # ---- configs[2]: ONE case sharded over the ranks (strong scaling)
    latency = None
    if world > 1:
        from brainseg_b200 import sharded as SH

        case0 = torch.from_numpy(SY.case_volume(0, VOL_SHAPE)).pin_memory()
        gt0 = torch.from_numpy(np.ascontiguousarray(SY.label_volume(0, VOL_SHAPE[1:]))).pin_memory()
        ref_labels = None
        if not sharded_only:
            # the same case on ONE GPU (every rank computes it redundantly: deterministic), for the equality check
            r1 = pipe.finish(pipe.submit(case0, gt0), post=False)
            ref_labels = r1["segmentation"].clone()
        records = {}
        routes = [args.route] if args.route != "both" else ["peer", "nccl"]
        for route in routes:
            sh = SH.ShardedExchange(rank, world, dev, route=route)
            # forwards in flight sized to the rank's share (an engine always runs its whole batch: 18 items per model on
            # 8 GPUs = two lanes of 9, not 16 + a chunk of 2 that costs 8)
            from brainseg_b200 import sliding as SL
            lat_batch = SL.balanced_batch(-(-N_TILES * N_MIRRORS // world), lanes=2)
            lp = PL.BratsCasePipeline([m1, m2], PATCH, args.step_size, (0, 1, 2), True, True, (1, 2, 3), "brats2025",
                                      batch=lat_batch, rank=rank, world_size=world, shard=sh)
            owner = (lambda i: i % world)
            run_steps(lp, max(2, min(args.warmup, 3)), [case0], [gt0], owner_of=owner, depth=2)
            barrier()
            ll0 = lp.kernel_launches()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lout, _ = run_steps(lp, args.steps, [case0], [gt0], owner_of=owner, want_host_labels=True, depth=2)
            e1.record()
            barrier()
            t_lat = e0.elapsed_time(e1) / 1e3
            lat_launches = lp.kernel_launches() - ll0 + sh.launches
            # single case, unpipelined: submit -> labels on the host (rank 0 also runs the post-processing)
            singles, singles_nopost = [], []
            for rep in range(3):
                barrier()
                t0 = time.perf_counter()
                pend = lp.submit(case0, gt0)
                pend["done"].synchronize()
                t_seg = time.perf_counter() - t0
                res = lp.finish(pend, post=(rank == 0))
                res["segmentation"].cpu()
                torch.cuda.synchronize()
                singles.append(time.perf_counter() - t0)
                singles_nopost.append(t_seg)
            # ... and with the post-processing fed SURVEY §8d's blobby label volume (≈225 components) instead of the
            # random-init nets' noise (one giant component + ~1e5 single-voxel foci, 80 ms of mostly host-side glue)
            blob0 = torch.from_numpy(np.ascontiguousarray(np.roll(gt0.numpy(), 3, axis=0))).to(dev) if rank == 0 else None
            singles_blobby = []
            for rep in range(3):
                barrier()
                t0 = time.perf_counter()
                pend = lp.submit(case0, gt0)
                res = lp.finish(pend, post=(rank == 0), labels=blob0)
                res["segmentation"].cpu()
                torch.cuda.synchronize()
                singles_blobby.append(time.perf_counter() - t0)
            # one more single case with phase events: this rank's forwards of each model and the exchange behind them
            barrier()
            lp.trace = []
            pend = lp.submit(case0, gt0)
            pend["done"].synchronize()
            tr, lp.trace = dict(lp.trace), None
            phases = torch.tensor([tr["start"].elapsed_time(tr["forwards_0"]), tr["forwards_0"].elapsed_time(tr["exchange_0"]),
                                   tr["forwards_0"].elapsed_time(tr["forwards_1"]), tr["forwards_1"].elapsed_time(tr["exchange_1"])],
                                  dtype=torch.float64, device=dev)
            dist.all_reduce(phases, op=dist.ReduceOp.MAX)
            lp.finish(pend, post=False)
            tt = torch.tensor([t_lat, statistics.median(singles), statistics.median(singles_nopost),
                               statistics.median(singles_blobby)], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_lat, t_single_lat, t_single_seg, t_single_blobby = tt.tolist()
            res = lp.finish(lp.submit(case0, gt0), post=False)
            same = None
            if ref_labels is not None:
                same = float((res["segmentation"] == ref_labels).float().mean().item())
            records[route] = {"ms_per_case": t_lat / args.steps * 1e3, "cases_per_s": args.steps / t_lat,
                              "single_case_ms": t_single_lat * 1e3, "single_case_ms_without_post": t_single_seg * 1e3,
                              "single_case_ms_blobby_post": t_single_blobby * 1e3,
                              "gpu_launches": int(lat_launches), "forwards_in_flight": lat_batch,
                              "single_case_phases_ms_max_over_ranks": {
                                  "model1_forwards": phases[0].item(), "model1_exchange_after_forwards": phases[1].item(),
                                  "model2_forwards": phases[2].item(), "model2_exchange_after_forwards": phases[3].item(),
                                  "note": "CUDA events on one unpipelined case; an exchange is timed from the end of the "
                                          "rank's own forwards to its label volume being whole (includes waiting for "
                                          "the slowest rank); model 1's exchange overlaps model 2's forwards"},
                              "exchange_bytes_per_case": int((sh.peer_bytes + sh.nccl_bytes) / max(1, sh.launches) * 2),
                              "labels_equal_to_1gpu": same, "label_sha256_16": sha16(res["segmentation"])}
            log(f"latency mode [{route}]: {t_lat / args.steps * 1e3:.1f} ms per case pipelined, single case "
                f"{t_single_lat * 1e3:.1f} ms ({t_single_seg * 1e3:.1f} ms to the label volumes, {t_single_blobby * 1e3:.1f} ms with "
                f"post-processing on blobby labels); labels equal to 1-GPU: {same}")
            if sharded_only and route == routes[0]:
                out, t_res, t_e2e, t_single, launches = lout, t_lat, t_lat, t_single_lat, lat_launches
                if out is None:  # rank 0 did not own the last case: any owned result serves the line
                    out = lp.finish(lp.submit(case0, gt0), post=True)
            del lp
            try:
                sh.close()
            except Exception as e:  # an in-kernel wait timed out on this rank: keep the line, flag the record
                records[route]["exchange_error"] = str(e)
                log(f"latency mode [{route}]: {e}")
        main_route = routes[0]
        latency = dict(records[main_route])
        latency.update({"config": "configs[2]: the 288 (model, tile, mirror) forwards of ONE case dealt round-robin to the "
                                  f"{world} ranks; exchange = {main_route} route (see brainseg_b200/sharded.py); steps "
                                  "pipelined as a case stream (two cases submitted ahead), post-processing of case i on rank i % N; host buffers "
                                  "(H2D of the volume on every rank and D2H of the labels inside the timed region)",
                        "scaling": "strong", "route": main_route, "steps": args.steps,
                        "label_sha256_16_1gpu": sha16(ref_labels) if ref_labels is not None else None})
        for r, rec in records.items():
            if r != main_route:
                latency[f"route_{r}"] = rec

    # ---- roofline pass: the conv stack of each model alone on one stream (the timed regions above run two stream
    # lanes whose kernels overlap, so per-kernel durations are taken here, live, with CUDA events, same inputs/buffers)
    conv_ms, conv_runs = [], 4
    for e in (eng1, eng2):
        e.run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(conv_runs):
            e.run()
        b.record()
        torch.cuda.synchronize()
        conv_ms.append(a.elapsed_time(b) / conv_runs)
    # dominant launch: the conv layer with the most FLOPs (model 2's 128->64 @128^3 brick-kernel launch), timed alone
    dom_eng = max((eng1, eng2), key=lambda e: max(i["flops"] for i in e.step_info))
    dom_idx = max(range(len(dom_eng.step_info)), key=lambda i: dom_eng.step_info[i]["flops"])
    dom_plan, dom_info = dom_eng.plans[dom_idx], dom_eng.step_info[dom_idx]
    dom_plan.run()
    torch.cuda.synchronize()
    dom_reps = 6
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(dom_reps):
        dom_plan.run()
    b.record()
    torch.cuda.synchronize()
    dom_ms = a.elapsed_time(b) / dom_reps
    clocks = sampler.stop() if sampler is not None else None

    times = torch.tensor([t_res, t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_res, t_e2e = times.tolist()
    cases = args.steps * (1 if sharded_only else world)

    if rank == 0:
        peaks = load_peaks()
        fwd_tflops = [e.flops_algo / (ms / 1e3) / 1e12 for e, ms in zip((eng1, eng2), conv_ms)]
        fwd_per_case = N_TILES * N_MIRRORS
        conv_alone_s = sum(ms / 1e3 / e.batch for ms, e in zip(conv_ms, (eng1, eng2))) * fwd_per_case
        achieved = dom_info["flops"] / (dom_ms / 1e3) / 1e12
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
        if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from one ncu --set full capture
            with open(tpath) as f:
                t = json.load(f)
            traffic = t["bytes_per_launch"]
            traffic_note = f"{t['source']}; algorithmic bytes {t['algorithmic_bytes']}"
        case_tflop = fwd_per_case * (eng1.flops_algo_per_item + eng2.flops_algo_per_item) / 1e12
        roofline = {"bound": "tensor",
                    "kernel": f"conv_brick_kernel (tcgen05 implicit-GEMM conv3d): {dom_info['name']}, "
                              f"{dom_eng.batch} tile-mirrors per launch ({dom_info['plan']})",
                    "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["tflops"], "frac_of_burst_peak": achieved / peaks["tflops_burst"],
                    "traffic": traffic, "traffic_source": traffic_note, "peak_source": peaks["source"],
                    "flops_per_launch": dom_info["flops"], "avg_launch_ms": dom_ms, "launches_timed": dom_reps,
                    "share_of_forward": dom_ms / conv_ms[1 if dom_eng is eng2 else 0],
                    "timed": "dedicated single-stream pass after the timed steps (the steps overlap two stream lanes), "
                             "CUDA events on the launching stream, same buffers",
                    "flops_accounting": "algorithmic (real channel counts: the 4 input channels are not counted as 16)",
                    "model1_forward_ms_per_batch": conv_ms[0], "model2_forward_ms_per_batch": conv_ms[1],
                    "model1_forward_tflops": fwd_tflops[0], "model2_forward_tflops_incl_norm_passes": fwd_tflops[1],
                    "case_tflop_algorithmic": case_tflop,
                    "case_tflops_end_to_end": case_tflop / (t_res / args.steps) if not sharded_only else None,
                    "conv_share_of_step": conv_alone_s / (t_res / args.steps) if not sharded_only else None}
        incumbent = None
        if not args.no_incumbent:
            try:
                gfs = [eng1.flops_algo_per_item / 1e9, eng2.flops_algo_per_item / 1e9]
                inc = time_incumbent((m1, m2), gfs, dev, batch=eng1.batch)
                ours_ms = [ms * inc[i]["batch"] / e.batch for i, (ms, e) in enumerate(zip(conv_ms, (eng1, eng2)))]
                incumbent = {"what": "the reference's own CUDA path on this GPU: torch.nn.functional / cuDNN, fp16 weights "
                                     "and activations, channels_last_3d, cudnn.benchmark, same two networks, same batch "
                                     "(conv stack + norms + sigmoid head of one forward batch)",
                             "model1": inc[0], "model2": inc[1],
                             "ours_ms_per_forward_batch": ours_ms,
                             "speedup_model1": inc[0]["ms_per_forward_batch"] / ours_ms[0],
                             "speedup_model2": inc[1]["ms_per_forward_batch"] / ours_ms[1],
                             "incumbent_conv_ms_per_case": fwd_per_case / inc[0]["batch"] * (inc[0]["ms_per_forward_batch"] + inc[1]["ms_per_forward_batch"])}
                log(f"incumbent (cuDNN fp16): {inc[0]['ms_per_forward_batch']:.1f} + {inc[1]['ms_per_forward_batch']:.1f} ms per batch "
                    f"of {inc[0]['batch']} vs ours {ours_ms[0]:.1f} + {ours_ms[1]:.1f} ms")
            except Exception as e:  # an out-of-memory cuDNN workspace must not take the line down
                incumbent = {"error": f"{type(e).__name__}: {e}"[:300]}
        hbm = None
        if not args.no_hbm:
            hbm = hbm_rooflines(dev, peaks, eng1.f16)
        # post-processing alone (host wall clock, device work included): on the case's own output and on blobby labels
        post = None
        if out is not None and not sharded_only:
            post = {}
            bl = torch.from_numpy(np.ascontiguousarray(np.roll(gt_hosts[0].numpy(), 3, axis=0))).to(dev)

            def post_ms(labels):  # second of two calls: the first re-warms the allocator after the passes above
                for _ in range(2):
                    pend = pipe.submit(dev_vols[0], dev_gts[0])
                    pend["done"].synchronize()
                    t0 = time.perf_counter()
                    r = pipe.finish(pend, labels=labels)
                    dt = (time.perf_counter() - t0) * 1e3
                return dt, r

            post["inference_output_ms"], _ = post_ms(None)
            post["blobby_labels_ms"], r = post_ms(bl)
            post["blobby_components"] = r["components"]["num_components"] + r["components"].get("excluded_fragments", 0)
            post["note"] = ("the random-init nets' own output is one giant component + ~1e5 single-voxel enhancing foci; "
                            "blobby = SURVEY §8d config-4 label volumes (smoothed-noise thresholds)")
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_for_ours(args.model2)
        golden = SY.load_config2_oracle()
        h2d = host_vols[0].numel() * 4 + gt_hosts[0].numel()
        d2h = gt_hosts[0].numel() + 257 * 8 + 4 + 2 * 4096 * 88 + 2 * 8 * 120  # labels + hist + ncomp + stats + moments
        line = {"metric": METRIC, "value": cases / t_res, "unit": "cases/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_res / args.steps * 1e3, "higher_is_better": True,
                "scaling": "strong" if sharded_only else "weak", "vs_baseline": None, "dtype": act_dtype,
                "data": "synthetic", "config": workload_config(args),
                "e2e": {"value": cases / t_e2e, "unit": "cases/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e / args.steps * 1e3,
                        "single_case_latency_ms": t_single * 1e3},
                "gpu_launches": int(launches), "roofline": roofline, "roofline_hbm": hbm, "incumbent": incumbent,
                "cpu_baseline": cpu, "clocks": clocks, "latency": latency, "post_processing": post,
                "fp16_range_guard": {"enabled": bool(eng1.guard or eng2.guard), "overflows": pipe.fp16_overflows},
                "result_check": result_check(out, golden, default_case and not blobby)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="throughput", choices=["throughput", "latency"],
                    help="N > 1: throughput = cases sharded (+ a latency record in the same line); latency = only the "
                         "one-case-sharded configuration, reported as the line's value")
    ap.add_argument("--route", default="both", choices=["peer", "nccl", "both"],
                    help="accumulator exchange of the sharded configuration (both: peer is reported, nccl beside it)")
    ap.add_argument("--model2", default="large", choices=["large", "standard"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-incumbent", action="store_true")
    ap.add_argument("--no-hbm", action="store_true")
    ap.add_argument("--skip-config0", action="store_true", help="reference arm: skip the full configs[0] run")
    ap.add_argument("--post-breakdown", action="store_true")
    ap.add_argument("--post-labels", default="inference", choices=["inference", "blobby"],
                    help="label volume the post-processing consumes: the inference output (default) or SURVEY §8d's "
                         "blobby synthetic labels (configs[3] as written)")
    ap.add_argument("--distinct", type=int, default=8, help="distinct seeded cases per rank (N > 1)")
    ap.add_argument("--batch", type=int, default=16, help="(tile, mirror) forwards in flight per model (2 stream lanes)")
    ap.add_argument("--lanes", type=int, default=2, help="stream lanes the forwards in flight are dealt to")
    ap.add_argument("--patch", type=int, default=128, help="cubic patch size (configs[4] sweep: 128 / 160)")
    ap.add_argument("--step-size", type=float, default=0.5, help="sliding-window step (configs[4] sweep: 0.5 / 0.25)")
    ap.add_argument("--volume", type=int, nargs=3, default=None, metavar=("Z", "Y", "X"),
                    help="volume extents (default 155 240 240; configs[4]: 256 256 256)")
    args = ap.parse_args()
    global PATCH, VOL_SHAPE, N_TILES
    if args.patch != 128 or args.volume is not None or args.step_size != 0.5:  # configs[4]: patch / overlap sweep
        from brainseg_b200 import sliding
        PATCH = (args.patch,) * 3
        if args.volume is not None:
            VOL_SHAPE = (4,) + tuple(args.volume)
        padded = [max(v, p) for v, p in zip(VOL_SHAPE[1:], PATCH)]
        steps = sliding.compute_steps_for_sliding_window(PATCH, padded, args.step_size)
        N_TILES = len(steps[0]) * len(steps[1]) * len(steps[2])
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
