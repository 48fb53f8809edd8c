"""Conv-kernel level parity (through the C ABI, bsg_conv_plan_*) against a plain PyTorch fp32 reference of the same op
on the same 16-bit-rounded operands: the in-consumer norm transform of the brick kernel (K chunk 32 / 64 channels x N tile
32 / 64 x with / without output statistics; resident weight slabs only), the 3x3x1 kernel of the kw-packed first layer with
its gather layout, and the fp16 range flag.  Reference op: ConvDropoutNormNonlin.forward,
model_architecture/generic_UNet.py:68-72 (norm + LeakyReLU of the producing block, then the consuming Conv3d)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _setup():
    from brainseg_b200 import _lib as L
    from brainseg_b200 import packing as P
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return L, P, torch.device("cuda", torch.cuda.current_device())


def _conv_case(cin, cout, N, D, H, W, stats, f16=True, in_norm=False, seed=0, scale_in=1.0, scale_w=1.0, want_flag=False):
    L, P, dev = _setup()
    g = torch.Generator(device="cpu").manual_seed(seed)
    dt = torch.float16 if f16 else torch.bfloat16
    x = (torch.randn(N, cin, D, H, W, generator=g) * scale_in).to(dt)
    xb = x.permute(0, 2, 3, 4, 1).contiguous().to(dev)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g) / (27 * cin) ** 0.5 * scale_w
    wp = P.pack_conv3_weight(w.to(dev), cin, dt)
    b = torch.randn(cout, generator=g).to(dev)
    bp = P.pad_bias(b, cout).to(dev)
    out = torch.zeros(N, D, H, W, cout, dtype=dt, device=dev)
    st = torch.zeros(N, cout, 2, dtype=torch.float64, device=dev) if stats else None
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    kw = dict(kind=L.BSG_CONV_K3, stride=1, N=N, D=D, H=H, W=W, cin=cin, in_ptr=xb.data_ptr(), in_ctot=cin, cout=cout,
              out_ptr=out.data_ptr(), out_ctot=cout, out_coff=0, weights=wp.data_ptr(), bias=bp.data_ptr(), act=0 if stats else 1,
              slope=0.01, stats=st.data_ptr() if stats else None, use_khshift=-1, max_ctas=0, in_f16=int(f16), out_f16=int(f16),
              overflow=flag.data_ptr() if want_flag else None)
    xr = x.float().to(dev)
    if in_norm:
        table = torch.zeros(N, cin, 4, device=dev)
        table[..., 0] = 0.5 + torch.rand(N, cin, generator=g).to(dev)       # scale
        table[..., 1] = 0.5 * torch.randn(N, cin, generator=g).to(dev)      # shift
        table[..., 2] = 0.01                                                # slope
        table[:, : cin // 4, 2] = 1.0                                       # identity channels (slope 1) as a concat half has
        kw.update(in_norm=table.data_ptr(), in_norm_c=cin)
        y = xr * table[..., 0].view(N, cin, 1, 1, 1) + table[..., 1].view(N, cin, 1, 1, 1)
        y = torch.where(y > 0, y, y * table[..., 2].view(N, cin, 1, 1, 1))
        xr = y.to(dt).float()  # the transform writes the stage back in the activation dtype
    plan = L.ConvPlan(**kw)
    info = plan.info()
    plan.run()
    plan.run()  # idempotent on the input (the transform works on the shared-memory copy only)
    torch.cuda.synchronize()
    ref = F.conv3d(xr, w.to(dt).float().to(dev), b, padding=1)
    pre = ref
    if not stats:
        ref = F.leaky_relu(ref, 0.01)
    got = out.permute(0, 4, 1, 2, 3).float()
    return L, info, got, ref, pre, st, flag


@pytest.mark.parametrize("cin,cout,D,H,W,stats", [
    (32, 32, 16, 16, 16, False),   # CC 32, NT 32, resident slabs (kw-fused boxes)
    (32, 32, 8, 32, 24, True),
    (64, 32, 16, 16, 16, False),   # CC 64, NT 32, resident
    (64, 32, 8, 16, 32, True),
    (32, 64, 8, 16, 16, False),    # CC 32, NT 64, resident
    (32, 64, 4, 32, 16, True),
])
def test_brick_conv_with_in_consumer_norm(cin, cout, D, H, W, stats):
    """y = conv3d(lrelu(x * scale[n, c] + shift[n, c])) with the affine + LeakyReLU applied to the activation boxes in
    shared memory; zero padding is padding of the NORMALISED tensor (NaN out-of-bounds fill -> 0), batch items carry
    different tables, and a channel range with slope 1 / identity passes through untouched."""
    L, info, got, ref, pre, st, _ = _conv_case(cin, cout, 3, D, H, W, stats, in_norm=True, seed=cin + cout + D)
    assert info.khshift >= 1000, "planner did not pick the in-consumer transform"
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"{cin}->{cout} @{D}x{H}x{W} stats={stats}: plan marker {info.khshift}, cc {info.cc}, max err {err:.4g} (ref max {scale:.3g})")
    assert torch.isfinite(got).all()
    assert err <= 3e-3 * max(scale, 1.0)
    if stats:
        s_ref = torch.stack([pre.sum(dim=(2, 3, 4)), (pre * pre).sum(dim=(2, 3, 4))], dim=-1)
        assert ((st.float() / 2 - s_ref).abs() / (s_ref.abs() + 1.0)).max().item() < 1e-3  # two runs accumulated


@pytest.mark.parametrize("cin,cout", [(64, 64), (128, 64), (128, 32)])
def test_in_consumer_norm_is_refused_for_streamed_slabs(cin, cout):
    """Layers whose weight slabs do not stay resident would transform every element three times (kw-shifted boxes) on a
    shared-memory port the MMA operand reads already fill — measured slower than the pass it replaces, so the planner
    refuses and the caller keeps bsg_norm_apply_lrelu."""
    L, _, _ = _setup()
    with pytest.raises(L.BsgError):
        _conv_case(cin, cout, 1, 8, 16, 16, False, in_norm=True)


def test_in_consumer_norm_is_refused_outside_the_brick_kernel():
    L, P, dev = _setup()
    x = torch.zeros(1, 8, 8, 8, 128, dtype=torch.float16, device=dev)
    out = torch.zeros(1, 8, 8, 8, 128, dtype=torch.float16, device=dev)
    wp = torch.zeros(27, 128, 128, dtype=torch.float16, device=dev)
    table = torch.zeros(1, 128, 4, device=dev)
    with pytest.raises(L.BsgError):  # Cout 128 -> tile kernel: no transform there
        L.ConvPlan(kind=L.BSG_CONV_K3, stride=1, N=1, D=8, H=8, W=8, cin=128, in_ptr=x.data_ptr(), in_ctot=128, cout=128,
                   out_ptr=out.data_ptr(), out_ctot=128, out_coff=0, weights=wp.data_ptr(), bias=None, act=1, slope=0.01,
                   stats=None, use_khshift=-1, max_ctas=0, in_f16=1, out_f16=1, in_norm=table.data_ptr(), in_norm_c=128)


@pytest.mark.parametrize("cin,cout,shape", [(32, 32, (16, 16, 16)), (64, 128, (8, 16, 16)), (64, 64, (4, 16, 16))])
def test_fp16_range_flag(cin, cout, shape):
    """The epilogue raises the overflow flag iff a stored fp16 value left the fp16 range (brick and tile kernels)."""
    D, H, W = shape
    _, _, got, _, _, _, flag = _conv_case(cin, cout, 1, D, H, W, False, seed=5, want_flag=True)
    assert int(flag.item()) == 0 and torch.isfinite(got).all()
    _, _, got, _, _, _, flag = _conv_case(cin, cout, 1, D, H, W, False, seed=5, scale_in=300.0, scale_w=300.0, want_flag=True)
    assert int(flag.item()) == 1 and not torch.isfinite(got).all()


@pytest.mark.parametrize("cout,stats,shape", [(32, False, (16, 16, 16)), (64, True, (8, 32, 24)), (64, False, (4, 16, 16))])
def test_brick_conv_3x3x1_kernel(cout, stats, shape):
    """kw_taps = 1: a (3, 3, 1) kernel over 16 input channels — the kw-packed first layer of the network — against
    torch's conv3d with the same kernel."""
    L, P, dev = _setup()
    D, H, W = shape
    N, cin = 2, 16
    g = torch.Generator(device="cpu").manual_seed(cout + D)
    x = torch.randn(N, cin, D, H, W, generator=g).to(torch.float16)
    xb = x.permute(0, 2, 3, 4, 1).contiguous().to(dev)
    w = torch.randn(cout, cin, 3, 3, 1, generator=g) / (9 * cin) ** 0.5
    wp = torch.zeros(9, P.round_up(cout, 32), 16, dtype=torch.float16)
    wp[:, :cout] = w[..., 0].permute(2, 3, 0, 1).reshape(9, cout, cin).to(torch.float16)  # taps (kd, kh)
    wp = wp.to(dev)
    b = torch.randn(cout, generator=g).to(dev)
    bp = P.pad_bias(b, cout).to(dev)
    out = torch.zeros(N, D, H, W, cout, dtype=torch.float16, device=dev)
    st = torch.zeros(N, cout, 2, dtype=torch.float64, device=dev) if stats else None
    plan = L.ConvPlan(kind=L.BSG_CONV_K3, stride=1, N=N, D=D, H=H, W=W, cin=cin, in_ptr=xb.data_ptr(), in_ctot=cin, cout=cout,
                      out_ptr=out.data_ptr(), out_ctot=cout, out_coff=0, weights=wp.data_ptr(), bias=bp.data_ptr(),
                      act=0 if stats else 1, slope=0.01, stats=st.data_ptr() if stats else None, use_khshift=-1, max_ctas=0,
                      in_f16=1, out_f16=1, kw_taps=1)
    plan.run()
    torch.cuda.synchronize()
    ref = F.conv3d(x.float().to(dev), w.to(torch.float16).float().to(dev), b, padding=(1, 1, 0))
    pre = ref
    if not stats:
        ref = F.leaky_relu(ref, 0.01)
    got = out.permute(0, 4, 1, 2, 3).float()
    err = (got - ref).abs().max().item()
    assert err <= 3e-3 * max(ref.abs().max().item(), 1.0)
    if stats:
        s_ref = torch.stack([pre.sum(dim=(2, 3, 4)), (pre * pre).sum(dim=(2, 3, 4))], dim=-1)
        assert ((st.float() - s_ref).abs() / (s_ref.abs() + 1.0)).max().item() < 1e-3


def test_gather_kwpack_matches_torch_layout():
    """bsg_gather_patch_tta(kwpack = 1): every mirrored copy of the tile holds, per voxel, its three w neighbours' channels
    in the COPY's orientation (zeros outside the tile) — the same tensor packing.kwpack_input builds from the flipped tile."""
    import ctypes as C
    L, P, dev = _setup()
    vol = torch.randn(4, 20, 24, 40, generator=torch.Generator().manual_seed(3)).to(dev)
    p0, p1, p2 = 16, 16, 32
    z0, y0, x0 = 3, 5, 7
    codes = (C.c_int * 8)(*range(8))
    out = torch.zeros(8, p0, p1, p2, 16, dtype=torch.float16, device=dev)
    L.check(L.lib().bsg_gather_patch_tta(C.c_void_p(vol.data_ptr()), 4, 20, 24, 40, z0, y0, x0, p0, p1, p2, codes, 8,
                                         C.c_void_p(out.data_ptr()), 16, 1, 1, L.stream_ptr()))
    tile = vol[:, z0:z0 + p0, y0:y0 + p1, x0:x0 + p2]
    for m in range(8):
        dims = [a for a, bit in ((3, 1), (2, 2), (1, 4)) if m & bit]
        flipped = torch.flip(tile, dims) if dims else tile
        ref = P.kwpack_input(flipped[None], 16)[0].to(torch.float16)
        assert torch.equal(out[m], ref), f"mirror code {m}"


# ---------------------------------------------------------------------------------------------- TMA tensor-store epilogue
@pytest.mark.parametrize("cin,cout,N,shape,tma", [
    (64, 64, 2, (8, 16, 16), 1),      # canonical 8x16x1x1 box, N tiles of 256 columns = 4 parities x 64
    (64, 32, 1, (8, 16, 16), 1),      # N tile = 8 parities x 32
    (128, 128, 3, (4, 4, 4), 1),      # 4x4x4x2 box, odd batch (partial tile along n), N tiles of 256 = 2 parities
    (320, 320, 2, (4, 4, 4), 1),      # cout_pad 320: N tiles straddle parity boundaries
    (64, 48, 1, (6, 10, 12), 1),      # cout 48 (channels 48..63 of the second chunk clipped), partial tiles in h and w
    (64, 64, 2, (8, 16, 16), 0),      # planner's choice (staged: 64-channel rows are whole lines)
    (64, 64, 2, (8, 16, 16), 2),      # the direct (per-thread row) epilogue
    (64, 32, 1, (8, 16, 16), 0),      # planner's choice (direct: 32-channel rows are half lines)
])
def test_transposed_conv_tma_store(cin, cout, N, shape, tma):
    """ConvTranspose3d k2 s2 (generic_UNet.py:363-364) through the tile kernel with tma_store = 1: the epilogue stages 32
    voxels x 32 channels per warp in shared memory and writes them with cp.async.bulk.tensor stores into the output's
    parity views; the output lands inside a wider buffer (the skip half of the concat buffer must stay untouched)."""
    L, P, dev = _setup()
    D, H, W = shape
    g = torch.Generator(device="cpu").manual_seed(cin + cout + D)
    x = torch.randn(N, cin, D, H, W, generator=g).to(torch.float16)
    xb = x.permute(0, 2, 3, 4, 1).contiguous().to(dev)
    w = torch.randn(cin, cout, 2, 2, 2, generator=g) / cin ** 0.5
    wp = P.pack_convT2_weight(w.to(dev), cin, torch.float16)
    ctot, coff = 2 * cout + 8, 8
    out = torch.full((N, 2 * D, 2 * H, 2 * W, ctot), 7.0, dtype=torch.float16, device=dev)
    plan = L.ConvPlan(kind=L.BSG_CONVT_K2S2, stride=1, N=N, D=D, H=H, W=W, cin=cin, in_ptr=xb.data_ptr(), in_ctot=cin,
                      cout=cout, out_ptr=out.data_ptr(), out_ctot=ctot, out_coff=coff, weights=wp.data_ptr(), bias=None,
                      act=0, slope=0.0, stats=None, use_khshift=0, max_ctas=0, in_f16=1, out_f16=1, tma_store=tma)
    plan.run()
    torch.cuda.synchronize()
    ref = F.conv_transpose3d(x.float().to(dev), w.to(torch.float16).float().to(dev), stride=2)
    got = out[..., coff:coff + cout].permute(0, 4, 1, 2, 3).float()
    err = (got - ref).abs().max().item()
    print(f"convT {cin}->{cout} @{D}x{H}x{W} x{N} tma_store={tma}: max err {err:.4g} (ref max {ref.abs().max().item():.3g})")
    assert err <= 2e-3 * max(ref.abs().max().item(), 1.0)
    assert (out[..., :coff] == 7.0).all() and (out[..., coff + cout:] == 7.0).all()  # neighbours in the buffer untouched


@pytest.mark.parametrize("cin,cout,stride,shape,stats", [
    (64, 128, 1, (8, 16, 16), False),   # KHS mode
    (64, 128, 2, (16, 16, 32), True),   # stride 2 + statistics
    (128, 256, 1, (4, 8, 8), False),    # 8x8x2 box
    (64, 80, 1, (5, 9, 11), True),      # cout 80 -> cout_pad 96, partial tiles everywhere
])
def test_tile_conv_with_forced_tma_store(cin, cout, stride, shape, stats):
    """tma_store = 1 on 3x3x3 convs of the tile kernel: same results as the direct epilogue, statistics and the fp16 flag
    are computed from the registers either way."""
    L, P, dev = _setup()
    D, H, W = shape
    N = 2
    g = torch.Generator(device="cpu").manual_seed(cin + cout + stride)
    x = torch.randn(N, cin, D, H, W, generator=g).to(torch.float16)
    xb = x.permute(0, 2, 3, 4, 1).contiguous().to(dev)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g) / (27 * cin) ** 0.5
    wp = P.pack_conv3_weight(w.to(dev), cin, torch.float16)
    b = torch.randn(cout, generator=g).to(dev)
    bp = P.pad_bias(b, cout).to(dev)
    Do, Ho, Wo = D // stride, H // stride, W // stride
    outs, sts = [], []
    for tma in (1, 2):
        out = torch.zeros(N, Do, Ho, Wo, cout, dtype=torch.float16, device=dev)
        st = torch.zeros(N, cout, 2, dtype=torch.float64, device=dev) if stats else None
        plan = L.ConvPlan(kind=L.BSG_CONV_K3, stride=stride, N=N, D=D, H=H, W=W, cin=cin, in_ptr=xb.data_ptr(), in_ctot=cin,
                          cout=cout, out_ptr=out.data_ptr(), out_ctot=cout, out_coff=0, weights=wp.data_ptr(),
                          bias=bp.data_ptr(), act=0 if stats else 1, slope=0.01, stats=st.data_ptr() if stats else None,
                          use_khshift=-1, max_ctas=0, in_f16=1, out_f16=1, algo=0, tma_store=tma)
        plan.run()
        torch.cuda.synchronize()
        outs.append(out)
        sts.append(st)
    ref = F.conv3d(x.float().to(dev), w.to(torch.float16).float().to(dev), b, stride=stride, padding=1)
    if not stats:
        ref = F.leaky_relu(ref, 0.01)
    got = outs[0].permute(0, 4, 1, 2, 3).float()
    assert (got - ref).abs().max().item() <= 3e-3 * max(ref.abs().max().item(), 1.0)
    # the staging buffers take shared memory from the pipeline, so the planner may pick another stage layout (tap order):
    # same values up to the order of the fp32 additions
    assert (outs[0].float() - outs[1].float()).abs().max().item() <= 2e-3 * max(ref.abs().max().item(), 1.0)
    if stats:
        assert torch.allclose(sts[0], sts[1], rtol=1e-4, atol=1e-3)


# ---------------------------------------------------------------------------------------------- M blocking
@pytest.mark.parametrize("cin,cout,stride,shape,stats", [
    (64, 128, 2, (16, 32, 32), True),    # stride 2, one-tap stages (the planner's own choice at the benchmark sizes)
    (32, 64, 2, (16, 32, 16), False),    # stride 2, three kh taps per stage, 32-channel parity views
    (128, 128, 1, (8, 16, 16), True),    # stride 1, haloed kh box with two planes per stage
    (64, 96, 1, (6, 16, 24), False),     # N tile 96, planes 4 and 5 form the last pair
    (64, 128, 2, (12, 40, 24), True),    # partial tiles in h / w next to the plane pairs (Do = 6, Ho = 20, Wo = 12)
])
def test_tile_conv_m_blocking(cin, cout, stride, shape, stats):
    """mblock = 1: a work item is two M tiles (adjacent output planes) sharing every weight stage, two accumulators per
    TMEM buffer.  Against torch's conv3d and against the unblocked plan (same tap order: equal up to nothing — the MMA
    sequence per tile is the same, so the results are bit-identical)."""
    L, P, dev = _setup()
    D, H, W = shape
    N = 2
    g = torch.Generator(device="cpu").manual_seed(cin + cout + stride + D)
    x = torch.randn(N, cin, D, H, W, generator=g).to(torch.float16)
    xb = x.permute(0, 2, 3, 4, 1).contiguous().to(dev)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g) / (27 * cin) ** 0.5
    wp = P.pack_conv3_weight(w.to(dev), cin, torch.float16)
    b = torch.randn(cout, generator=g).to(dev)
    bp = P.pad_bias(b, cout).to(dev)
    Do, Ho, Wo = D // stride, H // stride, W // stride
    outs, sts, marks = [], [], []
    for mblock in (1, 2):
        out = torch.zeros(N, Do, Ho, Wo, cout, dtype=torch.float16, device=dev)
        st = torch.zeros(N, cout, 2, dtype=torch.float64, device=dev) if stats else None
        plan = L.ConvPlan(kind=L.BSG_CONV_K3, stride=stride, N=N, D=D, H=H, W=W, cin=cin, in_ptr=xb.data_ptr(), in_ctot=cin,
                          cout=cout, out_ptr=out.data_ptr(), out_ctot=cout, out_coff=0, weights=wp.data_ptr(),
                          bias=bp.data_ptr(), act=0 if stats else 1, slope=0.01, stats=st.data_ptr() if stats else None,
                          use_khshift=-1, max_ctas=0, in_f16=1, out_f16=1, algo=0, pair=0, mblock=mblock)
        marks.append(plan.info().khshift)
        plan.run()
        torch.cuda.synchronize()
        outs.append(out)
        sts.append(st)
    assert marks[0] % 100 >= 10 and marks[1] % 100 < 10, f"plan markers {marks}: M blocking not taken / not switched off"
    ref = F.conv3d(x.float().to(dev), w.to(torch.float16).float().to(dev), b, stride=stride, padding=1)
    pre = ref
    if not stats:
        ref = F.leaky_relu(ref, 0.01)
    got = outs[0].permute(0, 4, 1, 2, 3).float()
    err = (got - ref).abs().max().item()
    print(f"M blocking {cin}->{cout} s{stride} @{D}x{H}x{W}: plan markers {marks}, max err {err:.4g}")
    assert err <= 3e-3 * max(ref.abs().max().item(), 1.0)
    assert (outs[0].float() - outs[1].float()).abs().max().item() <= 2e-3 * max(ref.abs().max().item(), 1.0)
    if stats:
        s_ref = torch.stack([pre.sum(dim=(2, 3, 4)), (pre * pre).sum(dim=(2, 3, 4))], dim=-1)
        assert ((sts[0].float() - s_ref).abs() / (s_ref.abs() + 1.0)).max().item() < 1e-3
