"""GPU bring-up of the tcgen05 conv kernel against torch (cuDNN/ATen fp32) on bf16-rounded operands.

Run on the GPU box: python scripts/bringup_conv.py [case substrings...]
"""
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from brainseg_b200 import _lib as L
from brainseg_b200 import packing as P

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def run_case(name, kind, N, D, H, W, cin, cout, stride=1, khshift=-1, act=0, stats=False, in_extra=0, out_extra=0,
             out_coff=0, seed=0, bias=True, time_it=False, algo=-1, f16=False, pair=-1):
    g = torch.Generator(device="cpu").manual_seed(seed)
    dt = torch.float16 if f16 else torch.bfloat16
    cin_pad = P.round_up(cin, 16)
    x = torch.randn(N, cin, D, H, W, generator=g)
    in_ctot = cin_pad + in_extra
    xb = torch.zeros(N, D, H, W, in_ctot, dtype=dt)
    xb[..., :cin] = x.permute(0, 2, 3, 4, 1).to(dt)
    if in_extra:
        xb[..., cin_pad:] = 7.0  # poison: must never be read
    xb = xb.to(dev)
    xr = xb[..., :cin].permute(0, 4, 1, 2, 3).float()
    if kind == L.BSG_CONV_K3:
        w = torch.randn(cout, cin, 3, 3, 3, generator=g) / (27 * cin) ** 0.5
        wp = P.pack_conv3_weight(w.to(dev), cin_pad, dt)
        wr = w.to(dt).float().to(dev)
        Do, Ho, Wo = D // stride, H // stride, W // stride
    elif kind == L.BSG_CONVT_K2S2:
        w = torch.randn(cin, cout, 2, 2, 2, generator=g) / cin ** 0.5
        wp = P.pack_convT2_weight(w.to(dev), cin_pad, dt)
        wr = w.to(dt).float().to(dev)
        Do, Ho, Wo = D * 2, H * 2, W * 2
    else:
        w = torch.randn(cout, cin, 1, 1, 1, generator=g) / cin ** 0.5
        wp = P.pack_conv1_weight(w.to(dev), cin_pad, dt)
        wr = w.to(dt).float().to(dev)
        Do, Ho, Wo = D, H, W
    b = torch.randn(cout, generator=g).to(dev) if bias and kind != L.BSG_CONVT_K2S2 else None
    bp = P.pad_bias(b, cout).to(dev) if b is not None else None
    out_ctot = P.round_up(cout, 8) + out_extra
    out = torch.full((N, Do, Ho, Wo, out_ctot), 5.0, dtype=dt, device=dev)
    st = torch.zeros(N, cout, 2, dtype=torch.float64, device=dev) if stats else None
    plan = L.ConvPlan(kind=kind, stride=stride, N=N, D=D, H=H, W=W, cin=cin_pad, in_ptr=xb.data_ptr(),
                      in_ctot=in_ctot, cout=cout, out_ptr=out.data_ptr(), out_ctot=out_ctot, out_coff=out_coff,
                      weights=wp.data_ptr(), bias=bp.data_ptr() if bp is not None else None, act=act, slope=0.01,
                      stats=st.data_ptr() if st is not None else None, use_khshift=khshift, max_ctas=0, algo=algo, pair=pair,
                      in_f16=1 if f16 else 0, out_f16=1 if f16 else 0)
    inf = plan.info()
    plan.run()
    torch.cuda.synchronize()
    if kind == L.BSG_CONV_K3:
        ref = F.conv3d(xr, wr, b, stride=stride, padding=1)
    elif kind == L.BSG_CONVT_K2S2:
        ref = F.conv_transpose3d(xr, wr, None, stride=2)
    else:
        ref = F.conv3d(xr, wr, b)
    pre = ref
    if act == 1:
        ref = F.leaky_relu(ref, 0.01)
    got = out[..., out_coff:out_coff + cout].permute(0, 4, 1, 2, 3).float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    ok = err <= (3e-3 if f16 else 2e-2) * max(scale, 1.0)
    msg = (f"{name}: box {inf.bw}x{inf.bh}x{inf.bd}x{inf.bn} ntile {inf.ntile}x{inf.n_ntiles} cc {inf.cc} "
           f"stages {inf.nstages} khs {inf.khshift} grid {inf.grid} smem {inf.smem_bytes} | max err {err:.4g} "
           f"(ref max {scale:.3g})")
    # untouched channels of the output buffer must keep the fill value
    if out_ctot > cout:
        mask = torch.ones(out_ctot, dtype=torch.bool, device=dev)
        mask[out_coff:out_coff + cout] = False
        untouched = bool((out[..., mask] == 5.0).all().item())
        ok = ok and untouched
        msg += f" untouched={untouched}"
    if stats:
        s_ref = torch.stack([pre.sum(dim=(2, 3, 4)), (pre * pre).sum(dim=(2, 3, 4))], dim=-1)
        serr = ((st.float() - s_ref).abs() / (s_ref.abs() + 1.0)).max().item()
        ok = ok and serr < 1e-3
        msg += f" stats relerr {serr:.3g}"
    if time_it:
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 10
        for _ in range(reps):
            plan.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        msg += f" | {ms:.3f} ms {inf.flops / ms / 1e9:.1f} TFLOP/s"
    print(("PASS " if ok else "FAIL ") + msg, flush=True)
    return ok


CASES = [
    # name, kwargs
    ("k3_c32_32_sw64", dict(kind=0, N=1, D=4, H=16, W=16, cin=32, cout=32, khshift=0)),
    ("k3_c32_32_sw64_khs", dict(kind=0, N=1, D=4, H=16, W=16, cin=32, cout=32, khshift=1)),
    ("k3_c64_64_sw128", dict(kind=0, N=1, D=4, H=16, W=16, cin=64, cout=64, khshift=0)),
    ("k3_c64_64_sw128_khs", dict(kind=0, N=1, D=4, H=16, W=16, cin=64, cout=64, khshift=1)),
    ("k3_c4_32_sw32", dict(kind=0, N=1, D=4, H=16, W=16, cin=4, cout=32, khshift=0)),
    ("k3_c4_32_sw32_khs", dict(kind=0, N=1, D=4, H=16, W=16, cin=4, cout=32, khshift=1)),
    ("k3_c128_128_multi", dict(kind=0, N=2, D=8, H=32, W=24, cin=128, cout=128, khshift=-1, act=1)),
    ("k3_stats", dict(kind=0, N=2, D=8, H=16, W=16, cin=32, cout=64, khshift=-1, stats=True)),
    ("k3_s2", dict(kind=0, N=1, D=8, H=32, W=16, cin=32, cout=64, stride=2)),
    ("k3_s2_small", dict(kind=0, N=2, D=16, H=16, W=16, cin=64, cout=96, stride=2, act=1)),
    ("k3_8cube", dict(kind=0, N=2, D=8, H=8, W=8, cin=64, cout=320, act=1)),
    ("k3_4cube", dict(kind=0, N=3, D=4, H=4, W=4, cin=320, cout=320, act=1, stats=True)),
    ("k3_s2_to4", dict(kind=0, N=1, D=8, H=8, W=8, cin=320, cout=320, stride=2)),
    ("k3_ragged", dict(kind=0, N=1, D=5, H=20, W=20, cin=32, cout=32, khshift=-1)),
    ("k3_ragged_nokhs", dict(kind=0, N=1, D=5, H=10, W=10, cin=32, cout=32, khshift=0)),
    ("k3_slices", dict(kind=0, N=1, D=4, H=16, W=16, cin=32, cout=32, in_extra=32, out_extra=32, out_coff=32)),
    ("convT", dict(kind=1, N=1, D=4, H=16, W=8, cin=64, cout=32)),
    ("convT_big", dict(kind=1, N=2, D=4, H=4, W=4, cin=320, cout=320, out_extra=320)),
    ("k1", dict(kind=2, N=1, D=4, H=16, W=16, cin=32, cout=32)),
    ("brick_c32_32", dict(kind=0, N=1, D=8, H=16, W=8, cin=32, cout=32)),
    ("brick_c32_32_multi", dict(kind=0, N=2, D=16, H=32, W=16, cin=32, cout=32, act=1)),
    ("brick_c64_32", dict(kind=0, N=1, D=8, H=16, W=16, cin=64, cout=32)),
    ("brick_c64_64_stream", dict(kind=0, N=2, D=8, H=32, W=16, cin=64, cout=64, act=1)),
    ("brick_c128_64_stream", dict(kind=0, N=1, D=8, H=16, W=16, cin=128, cout=64)),
    ("brick_c4_32", dict(kind=0, N=1, D=8, H=16, W=16, cin=4, cout=32)),
    ("brick_c4_64", dict(kind=0, N=1, D=8, H=16, W=16, cin=4, cout=64, act=1)),
    ("brick_c32_24", dict(kind=0, N=1, D=8, H=16, W=8, cin=32, cout=24)),
    ("brick_stats", dict(kind=0, N=2, D=8, H=16, W=16, cin=32, cout=64, stats=True)),
    ("brick_slices", dict(kind=0, N=1, D=8, H=16, W=16, cin=32, cout=32, in_extra=32, out_extra=32, out_coff=32)),
    ("brick_many_units", dict(kind=0, N=3, D=32, H=48, W=40, cin=32, cout=32, act=1)),
    ("brick_many_units_stream", dict(kind=0, N=3, D=16, H=48, W=40, cin=64, cout=64, act=1)),
    ("f16_tile_c128", dict(kind=0, N=2, D=8, H=32, W=24, cin=128, cout=128, act=1, f16=True)),
    ("f16_tile_s2", dict(kind=0, N=1, D=8, H=32, W=16, cin=32, cout=64, stride=2, f16=True)),
    ("f16_convT", dict(kind=1, N=1, D=4, H=16, W=8, cin=64, cout=32, f16=True)),
    ("f16_brick_c32_32", dict(kind=0, N=2, D=16, H=32, W=16, cin=32, cout=32, act=1, f16=True)),
    ("f16_brick_stats", dict(kind=0, N=2, D=8, H=16, W=16, cin=32, cout=64, stats=True, f16=True)),
    ("perfb_4_32_128", dict(kind=0, N=2, D=128, H=128, W=128, cin=4, cout=32, act=1, time_it=True)),
    ("perfb_32_32_128", dict(kind=0, N=2, D=128, H=128, W=128, cin=32, cout=32, act=1, time_it=True)),
    ("perfb_64_32_128", dict(kind=0, N=2, D=128, H=128, W=128, cin=64, cout=32, act=1, time_it=True)),
    ("perfb_64_64_128", dict(kind=0, N=2, D=128, H=128, W=128, cin=64, cout=64, act=1, time_it=True)),
    ("perfb_128_64_128", dict(kind=0, N=1, D=128, H=128, W=128, cin=128, cout=64, act=1, time_it=True)),
    ("perfb_64_64_64", dict(kind=0, N=8, D=64, H=64, W=64, cin=64, cout=64, act=1, time_it=True)),
    ("perfb_128_64_64", dict(kind=0, N=8, D=64, H=64, W=64, cin=128, cout=64, act=1, time_it=True)),
    ("perfh_128_64_64_f16", dict(kind=0, N=8, D=64, H=64, W=64, cin=128, cout=64, act=1, time_it=True, f16=True)),
    ("perfh_128_64_64_bf16", dict(kind=0, N=8, D=64, H=64, W=64, cin=128, cout=64, act=1, time_it=True)),
    ("perfh_128_64_64_bf16_stats", dict(kind=0, N=8, D=64, H=64, W=64, cin=128, cout=64, time_it=True, stats=True)),
    ("perfh_128_64_64_f16_stats", dict(kind=0, N=8, D=64, H=64, W=64, cin=128, cout=64, time_it=True, stats=True, f16=True)),
    ("perfh_64_32_128_f16", dict(kind=0, N=2, D=128, H=128, W=128, cin=64, cout=32, act=1, time_it=True, f16=True)),
    ("perfh_64_32_128_bf16", dict(kind=0, N=2, D=128, H=128, W=128, cin=64, cout=32, act=1, time_it=True)),
    ("perfh_64_32_128_f16_stats", dict(kind=0, N=2, D=128, H=128, W=128, cin=64, cout=32, time_it=True, stats=True, f16=True)),
    ("perfh_256_256_32_f16", dict(kind=0, N=8, D=32, H=32, W=32, cin=256, cout=256, act=1, time_it=True, f16=True)),
    ("perfh_256_256_32_bf16", dict(kind=0, N=8, D=32, H=32, W=32, cin=256, cout=256, act=1, time_it=True)),
    ("perfd_128_64_128_b8_f16_stats", dict(kind=0, N=8, D=128, H=128, W=128, cin=128, cout=64, time_it=True, stats=True, f16=True)),
    ("perfT_64_64_64_b8", dict(kind=1, N=8, D=64, H=64, W=64, cin=64, cout=64, out_extra=64, time_it=True, f16=True)),
    ("perfT_64_32_64_b8", dict(kind=1, N=8, D=64, H=64, W=64, cin=64, cout=32, out_extra=32, time_it=True)),
    ("perfs2_32_64_128_b8", dict(kind=0, N=8, D=128, H=128, W=128, cin=32, cout=64, stride=2, act=1, time_it=True)),
    ("pair_c128_128", dict(kind=0, N=2, D=8, H=32, W=16, cin=128, cout=128, act=1, pair=1)),
    ("pair_c64_256_stats", dict(kind=0, N=2, D=8, H=16, W=16, cin=64, cout=256, stats=True, pair=1, f16=True)),
    ("pair_s2", dict(kind=0, N=2, D=16, H=32, W=16, cin=32, cout=64, stride=2, pair=1)),
    ("pair_c320", dict(kind=0, N=4, D=8, H=8, W=8, cin=64, cout=320, act=1, pair=1)),
    ("pair_many", dict(kind=0, N=6, D=16, H=32, W=24, cin=128, cout=128, act=1, pair=1)),
    ("perfp_128_128_64_pair", dict(kind=0, N=8, D=64, H=64, W=64, cin=128, cout=128, act=1, time_it=True, pair=1)),
    ("perfp_128_128_64_solo", dict(kind=0, N=8, D=64, H=64, W=64, cin=128, cout=128, act=1, time_it=True, pair=0)),
    ("perfp_256_128_64_pair", dict(kind=0, N=8, D=64, H=64, W=64, cin=256, cout=128, act=1, time_it=True, pair=1)),
    ("perfp_256_128_64_solo", dict(kind=0, N=8, D=64, H=64, W=64, cin=256, cout=128, act=1, time_it=True, pair=0)),
    ("perfp_s2_64_128_128_pair", dict(kind=0, N=8, D=128, H=128, W=128, cin=64, cout=128, stride=2, act=1, time_it=True, pair=1)),
    ("perfp_s2_64_128_128_solo", dict(kind=0, N=8, D=128, H=128, W=128, cin=64, cout=128, stride=2, act=1, time_it=True, pair=0)),
    ("perfp_s2_32_64_128_pair", dict(kind=0, N=8, D=128, H=128, W=128, cin=32, cout=64, stride=2, act=1, time_it=True, pair=1)),
    ("perfp_512_256_32_pair", dict(kind=0, N=8, D=32, H=32, W=32, cin=512, cout=256, act=1, time_it=True, pair=1)),
    ("perfp_512_256_32_solo", dict(kind=0, N=8, D=32, H=32, W=32, cin=512, cout=256, act=1, time_it=True, pair=0)),
    ("perft_32_32_128", dict(kind=0, N=2, D=128, H=128, W=128, cin=32, cout=32, act=1, time_it=True, algo=0)),
    ("perft_64_64_128", dict(kind=0, N=2, D=128, H=128, W=128, cin=64, cout=64, act=1, time_it=True, algo=0)),
    ("perft_128_64_128", dict(kind=0, N=1, D=128, H=128, W=128, cin=128, cout=64, act=1, time_it=True, algo=0)),
    ("perf_32_32_128", dict(kind=0, N=1, D=128, H=128, W=128, cin=32, cout=32, act=1, time_it=True, algo=0)),
    ("perf_32_32_128_nokhs", dict(kind=0, N=1, D=128, H=128, W=128, cin=32, cout=32, act=1, khshift=0, time_it=True, algo=0)),
    ("perf_64_32_128", dict(kind=0, N=1, D=128, H=128, W=128, cin=64, cout=32, act=1, time_it=True, algo=0)),
    ("perf_64_64_64", dict(kind=0, N=1, D=64, H=64, W=64, cin=64, cout=64, act=1, time_it=True, algo=0)),
    ("perf_128_128_32", dict(kind=0, N=1, D=32, H=32, W=32, cin=128, cout=128, act=1, time_it=True)),
    ("perf_256_256_16", dict(kind=0, N=1, D=16, H=16, W=16, cin=256, cout=256, act=1, time_it=True)),
    ("perf_320_320_8_b8", dict(kind=0, N=8, D=8, H=8, W=8, cin=320, cout=320, act=1, time_it=True)),
]

def _watchdog(limit_s):
    """A hung kernel cannot be cancelled from inside the process: bail out hard so the GPU box is not held."""
    import threading

    state = {"t": time.time(), "name": "startup"}

    def loop():
        while True:
            time.sleep(1.0)
            if time.time() - state["t"] > limit_s:
                print(f"WATCHDOG: case {state['name']} exceeded {limit_s}s, aborting", flush=True)
                os._exit(3)

    threading.Thread(target=loop, daemon=True).start()
    return state


if __name__ == "__main__":
    wd = _watchdog(45)
    L.check(L.lib().bsg_check_device())
    print("SMs", L.lib().bsg_sm_count(), torch.cuda.get_device_name(0), flush=True)
    sel = sys.argv[1:]
    nfail = 0
    for name, kw in CASES:
        if sel and not any(s in name for s in sel):
            continue
        wd["t"], wd["name"] = time.time(), name
        try:
            if not run_case(name, **kw):
                nfail += 1
        except Exception as e:  # keep going: one bad shape must not hide the others
            nfail += 1
            print(f"ERROR {name}: {type(e).__name__}: {e}", flush=True)
            if "CUDA" in str(e) or "cuda" in str(e):
                break
    print("failures:", nfail, flush=True)
    sys.exit(1 if nfail else 0)
