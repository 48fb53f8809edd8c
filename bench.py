#!/usr/bin/env python
"""Benchmark of the hot path: cases/sec for a full BraTS case (BASELINE.json metric).

A "step" is one synthetic case through the whole path: sliding-window inference of 4x155x240x240 with the two-model
ensemble and 8-way mirror TTA (2 x 18 tiles x 8 mirrors = 288 forwards), finalize, label-round ensemble + BraTS
remap, Dice vs a synthetic ground truth, 26-connected components + per-component statistics, morphology moments.

  python bench.py --gpus N --steps K --warmup W            our sm_100a path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N ...            the reference's CPU implementation (oracle port) on the
                                                           host cores, bounded sample per step
Prints ONE JSON line (rank 0).
"""
import argparse
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
                "hbm": float(p.get("hbm_gbs", 6527.1)), "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tflops": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md: 1.4 PF sustained / 6.65 TB/s)"}


def build_models(model2):
    from tests.helpers import build_dropin_unet

    m1 = build_dropin_unet("bn", base=32, num_pool=5, seed=1)
    if model2 == "large":
        m2 = build_dropin_unet("gn", base=32, num_pool=5, seed=2, groups=8, encoder_scale=2, max_num_features=512)
    elif model2 == "standard":
        m2 = build_dropin_unet("gn", base=32, num_pool=5, seed=2, groups=8)
    else:
        raise ValueError(model2)
    return m1, m2


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
            # under load = samples above the idle clock
            load = [s for s in sm if s > 0.5 * max(sm)] or sm
            out = {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def cpu_reference_sample(model2, include_post=True, threads=None):
    """Times the oracle (reference CPU algorithm, fp32 torch eager on the host cores) on a bounded sample:
    one model-1 forward on a 128^3 tile + the full post-processing chain.  Returns seconds and the extrapolation."""
    from oracle import postproc as OP
    from oracle import synthetic as SY
    from oracle import unet as OU
    from tests.helpers import build_dropin_unet

    torch.set_num_threads(threads or usable_cpus())
    log(f"cpu baseline: {torch.get_num_threads()} threads")
    m1 = build_dropin_unet("bn", base=32, num_pool=5, seed=1)
    sd = {k: v.detach().float() for k, v in m1.state_dict().items()}
    arch = OU.arch_from_module(m1)
    x = torch.randn(1, 4, *PATCH, generator=torch.Generator().manual_seed(0))
    OU.forward(sd, arch, x[:, :, :32, :32, :32])  # warm the thread pool
    t0 = time.perf_counter()
    y = OU.forward(sd, arch, x)
    torch.sigmoid(y)
    t_fwd = time.perf_counter() - t0
    log(f"cpu baseline: one 128^3 forward {t_fwd:.2f} s")
    gf1 = OU.conv_flops(sd, arch, PATCH) / 1e9
    gf2 = 3342.2 if model2 == "large" else gf1
    t_post = 0.0
    if include_post:
        pred, gt = SY.label_pair(0, (240, 240, 155))
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
        t_post = time.perf_counter() - t0
        log(f"cpu baseline: post-processing chain {t_post:.2f} s")
    n_fwd = N_TILES * N_MIRRORS
    t_case = t_fwd * n_fwd * (1.0 + gf2 / gf1) + t_post
    return {"t_fwd": t_fwd, "t_post": t_post, "t_case": t_case, "gf1": gf1, "gf2": gf2,
            "threads": torch.get_num_threads()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_reference_sample(args.model2, include_post=True)
    t_post = base["t_post"]
    from oracle import unet as OU
    from tests.helpers import build_dropin_unet

    m1 = build_dropin_unet("bn", base=32, num_pool=5, seed=1)
    sd = {k: v.detach().float() for k, v in m1.state_dict().items()}
    arch = OU.arch_from_module(m1)
    x = torch.randn(1, 4, *PATCH, generator=torch.Generator().manual_seed(0))
    budget_s = 150.0
    steps_total = args.steps + args.warmup
    # one forward per step; fewer timed steps only if even that would blow the few-minutes budget
    times = []
    t_begin = time.perf_counter()
    for i in range(steps_total):
        t0 = time.perf_counter()
        torch.sigmoid(OU.forward(sd, arch, x))
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > budget_s and len(times) >= 1:
            break
    t_fwd = sum(times) / len(times)
    n_fwd = N_TILES * N_MIRRORS
    t_case = t_fwd * n_fwd * (1.0 + base["gf2"] / base["gf1"]) + t_post
    value = 1.0 / t_case
    sample = (f"per step: 1 of the {2 * n_fwd} forwards of a case (model 1, one 128^3 tile, fp32 torch eager, "
              f"{base['threads']} threads, {t_fwd:.2f} s); model-2 forwards scaled by FLOPs ({base['gf2']:.0f}/"
              f"{base['gf1']:.0f} GF); post-processing chain measured once in full ({t_post:.1f} s); "
              f"case time extrapolated = {t_case:.0f} s")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "cases/s", "n_gpus": args.gpus,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": t_case * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": "cases/s", "cores": base["threads"], "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "cases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args):
    sweep = ""
    if PATCH != (128, 128, 128) or VOL_SHAPE != (4, 155, 240, 240) or args.step_size != 0.5:
        sweep = (f"configs[4] sweep point: volume {'x'.join(map(str, VOL_SHAPE))}, patch {PATCH[0]}^3, step {args.step_size}, "
                 f"{N_TILES} tiles; otherwise as ")
    return {"workload": sweep + "configs[1]: one synthetic BraTS case 4x155x240x240 (C,z,y,x), patch 128^3, step 0.5, "
                        "Gaussian weighting, 8-way mirror TTA, 2-model ensemble (model 1: Generic_UNet BN 31.2 M; "
                        f"model 2: GroupNorm {'large 87.4 M (encoder_scale 2, max 512)' if args.model2 == 'large' else 'standard 31.2 M'}), "
                        "one fold per model, regions threshold, label-round ensemble, BraTS-2025 remap, Dice vs "
                        "synthetic GT, 26-conn components + stats, morphology moments",
            "forwards_per_case": 2 * N_TILES * N_MIRRORS, "mode": args.mode, "forwards_in_flight": args.batch,
            "case_stream": "steps are timed as a stream of cases: case i+1's inference is submitted before case i's "
                           "post-processing is collected (all submitted cases finish inside the timed region)",
            "l2_policy": "inputs (143 MB fp32 volume, >=1 GB activations per layer) exceed the 126 MB L2",
            "parallelism": f"cases sharded over {args.gpus} GPU(s), no data-path collective"
            if args.mode == "throughput" else f"(tile,mirror) work items of one case sharded over {args.gpus} GPU(s), "
                                              "one NCCL all-reduce of the fp32 accumulator per model"}


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
    from oracle import synthetic as SY

    torch.set_num_threads(usable_cpus())
    log(f"rank {rank}/{world}: building models ({torch.get_num_threads()} host threads)")
    m1, m2 = build_models(args.model2)
    log("models built")
    reduce_fn = None
    if args.mode == "latency" and world > 1:
        def reduce_fn(acc):
            dist.all_reduce(acc)
            return acc
    pipe = PL.BratsCasePipeline([m1, m2], PATCH, args.step_size, (0, 1, 2), True, True, (1, 2, 3), "brats2025", batch=args.batch,
                                rank=rank if args.mode == "latency" else 0,
                                world_size=world if args.mode == "latency" else 1, reduce_fn=reduce_fn)
    eng1, eng2 = pipe.predictors[0].engine, pipe.predictors[1].engine
    kinds = sorted({"fp16" if e.f16 else "bf16" for e in (eng1, eng2)})
    act_dtype = kinds[0] if len(kinds) == 1 else "+".join(kinds)  # 16-bit operands, fp32 accumulation in TMEM
    log(f"engines ready: {len(pipe.predictors[0].engines)} lane(s) x batch {eng1.batch}; "
        f"{eng1.launches_per_forward} + {eng2.launches_per_forward} launches per forward batch, "
        f"{eng1.flops_per_item / 1e9:.1f} + {eng2.flops_per_item / 1e9:.1f} GF per tile-mirror")

    # synthetic inputs: pinned host volume (seeded per rank) + synthetic ground truth labels
    seed = 0 if args.mode == "latency" else rank
    host_vol = torch.from_numpy(SY.case_volume(seed, VOL_SHAPE)).pin_memory()
    gt_host = torch.from_numpy(np.ascontiguousarray(SY.label_volume(seed, VOL_SHAPE[1:]))).pin_memory()
    dev_vol = host_vol.to(dev)
    dev_gt = gt_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(k, vol, gt):
        """k cases as a stream (throughput mode of the cohort configs): the inference of case i+1 is submitted before
        the post-processing of case i is collected, so the host-side glue overlaps device work.  Every case is
        submitted AND finished inside the call."""
        out = seg_host = None
        pend = pipe.submit(vol, gt)
        for i in range(k):
            nxt = pipe.submit(vol, gt) if i + 1 < k else None
            out = pipe.finish(pend)
            if vol is host_vol:
                seg_host = out["segmentation"].cpu()  # D2H of the final label volume
            pend = nxt
        return out, seg_host

    log("inputs ready; warm-up")
    t0 = time.time()
    out, _ = run_steps(max(args.warmup, 1), host_vol, gt_host)
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
    out, _ = run_steps(args.steps, dev_vol, dev_gt)
    e1.record()
    barrier()
    t_res = e0.elapsed_time(e1) / 1e3
    log(f"resident: {t_res / args.steps:.3f} s per case")
    launches = (pipe.kernel_launches() - l0) + args.steps * pipe.extra_launches
    # ---- end to end: host buffers, H2D + D2H inside the timed region
    barrier()
    e0.record()
    out, seg_host = run_steps(args.steps, host_vol, gt_host)
    e1.record()
    barrier()
    t_e2e = e0.elapsed_time(e1) / 1e3
    log(f"e2e: {t_e2e / args.steps:.3f} s per case")
    # single-case latency (no overlap between consecutive cases), host buffers
    barrier()
    t0 = time.perf_counter()
    out1 = pipe.run_case(host_vol, gt=gt_host)
    out1["segmentation"].cpu()
    torch.cuda.synchronize()
    t_single = time.perf_counter() - t0
    log(f"single case, unpipelined: {t_single:.3f} s")
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
        conv_ms.append(a.elapsed_time(b))
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
    cases = args.steps * (world if args.mode == "throughput" else 1)

    if rank == 0:
        peaks = load_peaks()
        # forward aggregates (all conv launches of each model) and the dominant launch
        fwd_tflops = [e.flops * conv_runs / (ms / 1e3) / 1e12 for e, ms in zip((eng1, eng2), conv_ms)]
        fwd_per_case = N_TILES * N_MIRRORS
        conv_alone_s = sum(ms / 1e3 / conv_runs / e.batch for ms, e in zip(conv_ms, (eng1, eng2))) * fwd_per_case
        achieved = dom_info["flops"] / (dom_ms / 1e3) / 1e12
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
        if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from one ncu --set full capture
            with open(tpath) as f:
                t = json.load(f)
            traffic = t["bytes_per_launch"]
            traffic_note = f"{t['source']}; algorithmic bytes {t['algorithmic_bytes']}"
        roofline = {"bound": "tensor",
                    "kernel": f"conv_brick_kernel (tcgen05 implicit-GEMM conv3d): {dom_info['name']}, "
                              f"{dom_eng.batch} tile-mirrors per launch ({dom_info['plan']})",
                    "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["tflops"], "traffic": traffic, "traffic_source": traffic_note,
                    "peak_source": peaks["source"],
                    "flops_per_launch": dom_info["flops"], "avg_launch_ms": dom_ms, "launches_timed": dom_reps,
                    "share_of_forward": dom_ms / (conv_ms[1 if dom_eng is eng2 else 0] / conv_runs),
                    "timed": "dedicated single-stream pass after the timed steps (the steps overlap two stream lanes), "
                             "CUDA events on the launching stream, same buffers",
                    "model1_forward_tflops": fwd_tflops[0], "model2_forward_tflops_incl_norm_passes": fwd_tflops[1],
                    "conv_share_of_step": conv_alone_s / (t_res / args.steps)}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            c = cpu_reference_sample(args.model2, include_post=True)
            cpu = {"value": 1.0 / c["t_case"], "unit": "cases/s", "cores": c["threads"], "kind": "port",
                   "sample": (f"1 model-1 forward on one 128^3 tile ({c['t_fwd']:.2f} s, fp32 torch eager) of the "
                              f"{2 * N_TILES * N_MIRRORS} forwards per case, model-2 forwards scaled by FLOPs "
                              f"({c['gf2']:.0f}/{c['gf1']:.0f} GF), + the full post-processing chain measured in "
                              f"full ({c['t_post']:.1f} s); extrapolated case time {c['t_case']:.0f} s")}
        h2d = host_vol.numel() * 4 + gt_host.numel()
        d2h = seg_host.numel() + 257 * 8 + 4 + 2 * 4096 * 88 + 2 * 8 * 120  # labels + hist + ncomp + stats + moments
        line = {"metric": METRIC, "value": cases / t_res, "unit": "cases/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_res / args.steps * 1e3, "higher_is_better": True,
                "scaling": "weak" if args.mode == "throughput" else "strong", "vs_baseline": None, "dtype": act_dtype,
                "data": "synthetic", "config": workload_config(args),
                "e2e": {"value": cases / t_e2e, "unit": "cases/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e / args.steps * 1e3,
                        "single_case_latency_ms": t_single * 1e3},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
                "result_check": {"num_components": out["components"]["num_components"],
                                 "mean_dice": float(out["evaluation"]["mean_dice"])}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="throughput", choices=["throughput", "latency"])
    ap.add_argument("--model2", default="large", choices=["large", "standard"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=16, help="(tile, mirror) forwards in flight per model (2 stream lanes)")
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
