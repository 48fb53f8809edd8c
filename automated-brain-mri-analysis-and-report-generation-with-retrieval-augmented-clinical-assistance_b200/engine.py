"""B200 execution engine for Generic_UNet: turns the module tree into a fixed sequence of sm_100a kernel launches.

Data layout in HBM: every activation is a channels-last (N, D, H, W, C) 16-bit tensor, IEEE fp16 by default — the
type the reference's own CUDA path computes in under torch.cuda.amp.autocast; same tensor-pipe rate and bytes as bf16
with 3 more mantissa bits, which is what it takes to keep label agreement with the fp32 path above 99.9 % on
random-init weights (bf16: 99.67 %, tests/test_gpu_unet.py::test_config1_full_case_brats_architecture).
BSG_ACT_DTYPE=bf16 trades that for bf16's range (un-normalised activations beyond 65504).  The skip connection of level d
and the transposed-conv output that is concatenated with it (generic_UNet.py:435-438) share one buffer of 2*C
channels — the encoder conv writes channels [C, 2C), the transposed conv writes [0, C) — so `torch.cat` never runs and
the first decoder conv reads one tensor.  Eval-mode BatchNorm is folded into the conv weights; InstanceNorm / GroupNorm
use per-(n, c) sums produced by the conv epilogue and one in-place normalise + LeakyReLU pass.
"""
import ctypes as C
import os

import torch
from torch import nn

from . import _lib as L
from . import packing as P


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class _Act:
    """A channel slice [coff, coff+c) of a channels-last 16-bit buffer (n, d, h, w, ctot).  `parts` (engine dtype "fp32"
    only): the slice holds fp16x3 split tensors — [(offset within the slice, logical channels)], each occupying the
    three blocks [hi | hi | lo] of that many channels."""

    def __init__(self, buf, coff, c, parts=None):
        self.buf, self.coff, self.c, self.parts = buf, coff, c, parts

    @property
    def ctot(self):
        return self.buf.shape[-1]

    @property
    def spatial(self):
        return tuple(self.buf.shape[1:4])

    def ptr(self):
        return self.buf.data_ptr() + 2 * self.coff

    def view(self):
        return self.buf[..., self.coff:self.coff + self.c]


_overflow_flags = {}


def overflow_flag(device):
    """The device's fp16-overflow flag (one int32): set by any guarded conv epilogue, cleared by whoever handles it."""
    device = torch.device(device)
    if device not in _overflow_flags:
        _overflow_flags[device] = torch.zeros(1, dtype=torch.int32, device=device)
    return _overflow_flags[device]


class UNetEngine:
    def __init__(self, net, patch_size, batch, device=None, act_dtype=None):
        if not torch.cuda.is_available():
            raise L.BsgError("brainseg_b200 needs a CUDA device (sm_100); there is no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(self.device):
            L.check(L.lib().bsg_check_device())
        self.net, self.batch, self.patch = net, int(batch), tuple(int(p) for p in patch_size)
        div = [int(v) for v in net.input_shape_must_be_divisible_by]
        if any(p % d for p, d in zip(self.patch, div)):
            raise ValueError(f"patch size {self.patch} must be divisible by {div}")
        # activation dtype: the network's own setting (SegmentationNetwork.engine_dtype — the fp16 range guard flips it
        # to "bf16" after an overflow) wins over the BSG_ACT_DTYPE environment default
        mode = (act_dtype or getattr(net, "engine_dtype", None) or os.environ.get("BSG_ACT_DTYPE", "auto")).lower()
        if mode not in ("auto", "bf16", "fp16", "fp32"):
            raise ValueError(f"activation dtype {mode!r}: expected auto, bf16, fp16 or fp32")
        # auto = fp16 for every stack.  (The InstanceNorm / GroupNorm stacks need it outright: bf16 activations miss the
        # 1e-2 probability bar there.)
        self.f16 = int(mode != "bf16")
        # "fp32": the reference's CPU arithmetic (generic_UNet.py:68-72 outside autocast) on the 16-bit tensor pipe — every
        # activation y is kept as the fp16 pair hi = fp16(y), lo = fp16(y - hi) in three channel blocks [hi | hi | lo],
        # every weight as [w_hi | w_lo | w_hi] along K, so one fp32 multiply-add becomes three fp16 MMAs with fp32
        # accumulation (hi*w_hi + hi*w_lo + lo*w_hi; the dropped lo*w_lo term is O(2^-22)).  3x the MMA work and bytes
        # of the fp16 mode, tile kernel only; norms and the head work on hi + lo in fp32.
        self.split = mode == "fp32"
        # fp16 range guard: one device flag (shared by all engines of the device), raised by a conv epilogue when a value
        # it stored left the fp16 range (bsg_conv_desc.overflow); the pipeline reads it back once per case
        self.guard = bool(self.f16) and os.environ.get("BSG_OVERFLOW_GUARD", "1") != "0"
        self._overflow = overflow_flag(self.device) if self.guard else None
        self._weight_slots = []  # (kind, module, packed tensors): reload_weights() re-packs into them in place
        # InstanceNorm / GroupNorm blocks whose only consumer is a brick-kernel conv hand their normalise + LeakyReLU to
        # that conv (applied in shared memory on the way to the tensor core) instead of a separate HBM pass
        self.fuse_norm = os.environ.get("BSG_FUSE_NORM", "1") != "0" and not self.split
        self.fused_norms = 0
        self.try_kwpack = os.environ.get("BSG_KWPACK", "1") != "0" and not self.split
        # tile-kernel epilogue through shared memory + TMA tensor stores (measurement switch): 0 = planner's choice, 1 = every
        # tile-kernel layer, 2 = none, 3 = the transposed convs only
        self.tma_store = int(os.environ.get("BSG_TMA_STORE", "0"))
        self.mblock = int(os.environ.get("BSG_MBLOCK", "0"))  # tile-kernel M blocking: 0 planner's choice, 1 wherever possible, 2 off
        self.kwpack = False  # the first conv reads the kw-packed input layout (set by _add_block when its plan took it)
        self.flops_algo = 0.0    # algorithmic FLOPs on the real channel counts (the 4 input channels are padded to 16)
        self.act_dtype = torch.float16 if self.f16 else torch.bfloat16
        self.steps = []       # callables, in launch order
        self.step_info = []   # per step: name, algorithmic flops, plan geometry (diagnostics / bench breakdown)
        self.plans = []       # per step: the conv plan (bench.py times the dominant layer's launch alone)
        self.keep = []        # tensors the plans point at
        # norm statistics of every IN / GN layer live in one arena, zeroed by ONE fill at the start of a forward
        # (run()); a step called on its own (diagnostics) zeroes its slice itself
        self._stats_arena = None
        self._stats_used = 0
        self._in_run = False
        self.launches_per_forward = 0
        self.flops = 0.0
        self.final_norm = None  # (scale_shift [batch][C][2], slope) when the last block's norm is left to the head
        self.sub_events = None  # scripts/diag_case.py: events recorded between a conv and its norm passes
        self.event_log = None  # bench.py: list collecting (start, end) CUDA events around each run()
        self._build()

    # ------------------------------------------------------------------ construction
    def _alloc(self, spatial, c):
        t = torch.zeros((self.batch,) + tuple(spatial) + (c,), dtype=self.act_dtype, device=self.device)
        self.keep.append(t)
        return t

    def _pack_block(self, blk, cin_pad, kwpack=False, parts=None):
        """Packed 16-bit weights + fp32 bias of one conv block (eval BatchNorm folded in), norm affine parameters."""
        conv, norm = blk.conv, blk.instnorm
        w = conv.weight.detach().to(self.device, torch.float32)
        b = conv.bias.detach().to(self.device, torch.float32) if conv.bias is not None else torch.zeros(
            w.shape[0], device=self.device)
        gamma = beta = None
        if isinstance(norm, nn.BatchNorm3d):
            # eval BatchNorm == per-channel affine: fold into the conv (generic_UNet.py:72 with network.eval())
            scale = norm.weight.detach().to(self.device).float() / torch.sqrt(
                norm.running_var.detach().to(self.device).float() + norm.eps)
            w = w * scale.view(-1, 1, 1, 1, 1)
            b = (b - norm.running_mean.detach().to(self.device).float()) * scale + norm.bias.detach().to(
                self.device).float()
        elif isinstance(norm, (nn.InstanceNorm3d, nn.GroupNorm)):
            gamma = norm.weight.detach().to(self.device).float().contiguous() if norm.weight is not None else None
            beta = norm.bias.detach().to(self.device).float().contiguous() if norm.bias is not None else None
        else:
            raise NotImplementedError(f"norm {type(norm).__name__}")
        if parts is not None:
            wp = P.pack_conv3_weight(P.split_k_weight(w, parts, cin_pad, 1), cin_pad, self.act_dtype)
        elif kwpack:
            wp = P.pack_conv3_weight_kwpacked(w, self.act_dtype)
        else:
            wp = P.pack_conv3_weight(w, cin_pad, self.act_dtype)
        return (wp, P.pad_bias(b, w.shape[0]).to(self.device), gamma, beta)

    def _overflow_slot(self):
        return self._overflow.data_ptr() if self._overflow is not None else None

    def _add_block(self, blk, src, dst, spatial_in, defer_apply=False, claim=None, first=False):
        """One ConvDropoutNormNonlin block.  `claim`: the norm state of the block that produced `src`, when this conv is
        its only consumer — if the planner can run this conv with the in-consumer transform (brick kernel), the
        producer's normalise + LeakyReLU pass is dropped and its statistics go to this conv's input table instead.
        Returns the block's own norm state (None for BatchNorm-folded blocks)."""
        conv, norm = blk.conv, blk.instnorm
        stride = int(conv.stride[0])
        slope = float(blk.lrelu.negative_slope)
        cout = conv.out_channels
        stats = None
        if isinstance(norm, nn.BatchNorm3d):
            act = L.BSG_ACT_LRELU
        else:
            act = L.BSG_ACT_NONE
        cin_pad = src.c
        d, h, wd = spatial_in
        # the network's first conv: with 3 * C <= 16 input channels the gather kernel packs the three w neighbours of a
        # voxel into its 16 channels and the conv runs as a 3x3x1 kernel — 9 taps of K = 16 instead of 27
        kwpack = False
        if first and self.try_kwpack and 3 * conv.in_channels <= 16 and cin_pad == 16 and stride == 1:
            kwpack = True
        wp, bp, gamma, beta = self._pack_block(blk, cin_pad, kwpack, src.parts)
        if act == L.BSG_ACT_NONE:
            stats = self._carve_stats(cout)
        desc = dict(out_split_stride=cout if self.split else 0, tma_store={1: 1, 2: 2}.get(self.tma_store, 0),
                    mblock=self.mblock,
                    kind=L.BSG_CONV_K3, stride=stride, N=self.batch, D=d, H=h, W=wd, cin=cin_pad,
                    in_ptr=src.ptr(), in_ctot=src.ctot, cout=cout, out_ptr=dst.buf.data_ptr(),
                    out_ctot=dst.ctot, out_coff=dst.coff, weights=wp.data_ptr(), bias=bp.data_ptr(), act=act,
                    slope=slope, stats=stats.data_ptr() if stats is not None else None,
                    out_f16=self.f16, in_f16=self.f16, use_khshift=-1,
                    max_ctas=0, overflow=self._overflow_slot())
        plan = None
        if kwpack:
            try:
                plan = L.ConvPlan(kw_taps=1, **desc)
                self.kwpack = True
            except L.BsgError:  # shape does not suit the brick kernel: plain 27-tap first layer
                kwpack = False
                wp, bp, gamma, beta = self._pack_block(blk, cin_pad, False)
                desc.update(weights=wp.data_ptr(), bias=bp.data_ptr())
        self.keep += [wp, bp]
        self._weight_slots.append(("block_kw" if kwpack else "block", blk, (cin_pad, src.parts), (wp, bp, gamma, beta)))
        if claim is not None and self.fuse_norm and src.coff == 0 and src.c == src.ctot:
            table = torch.zeros(self.batch, cin_pad, 4, dtype=torch.float32, device=self.device)
            try:
                plan = L.ConvPlan(in_norm=table.data_ptr(), in_norm_c=cin_pad,
                                  in_norm_cc=int(os.environ.get("BSG_XF_CC", "0")), **desc)
            except L.BsgError:
                plan = None  # layer does not suit the brick kernel: the producer keeps its separate pass
            if plan is not None:
                claim["table"] = table
                claim["apply"] = False
                self.keep.append(table)
                self.launches_per_forward -= 1
                self.fused_norms += 1
        if plan is None:
            plan = L.ConvPlan(**desc)
        self.flops += plan.info().flops
        so_ = tuple(s // stride for s in spatial_in)
        self.flops_algo += 2.0 * 27 * conv.in_channels * cout * so_[0] * so_[1] * so_[2] * self.batch
        self._note(f"conv3 s{stride} {cin_pad}->{cout} @{'x'.join(map(str, spatial_in))}"
                   f"{' +norm' if stats is not None else ''}{' (input normalised in shared memory)' if plan.desc.in_norm else ''}",
                   plan)
        lib = L.lib()
        if stats is None:
            self.steps.append(plan.run)
            self.launches_per_forward += 1
            return None
        groups = norm.num_groups if isinstance(norm, nn.GroupNorm) else 0
        ss = torch.empty(self.batch, cout, 2, dtype=torch.float32, device=self.device)
        self.keep += [stats, ss, gamma, beta]
        so = tuple(s // stride for s in spatial_in)
        vox = so[0] * so[1] * so[2]
        eps = float(norm.eps)
        gp = _ptr(gamma) if gamma is not None else None
        bp2 = _ptr(beta) if beta is not None else None

        # apply: this block runs its own normalise + LeakyReLU pass; table: the consuming conv applies it on the fly
        # (set by the consumer's _add_block when it claims this block), the statistics then go to its input table
        state = {"apply": not defer_apply, "table": None}

        def run(stream=None, plan=plan, stats=stats, ss=ss, state=state):
            sp = L.stream_ptr(stream)
            if not self._in_run:
                stats.zero_()
            plan.run(stream)
            if self.sub_events is not None:  # diagnostics: split the step into conv | norm passes
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                self.sub_events.append(ev)
            if state["table"] is not None:
                L.check(lib.bsg_norm_finalize_table(_ptr(stats), self.batch, cout, groups, float(vox), eps, gp, bp2, slope,
                                                    _ptr(state["table"]), cout, 0, sp))
                return
            L.check(lib.bsg_norm_finalize(_ptr(stats), self.batch, cout, groups, float(vox), eps, gp, bp2, _ptr(ss), sp))
            if state["apply"] and self.split:
                L.check(lib.bsg_norm_apply_lrelu_split(_ptr(dst.buf), vox, self.batch, cout, dst.ctot, dst.coff, _ptr(ss),
                                                       slope, sp))
            elif state["apply"]:
                L.check(lib.bsg_norm_apply_lrelu(_ptr(dst.buf), vox, self.batch, cout, dst.ctot, dst.coff, _ptr(ss), slope,
                                                 self.f16, self.f16, sp))

        self.steps.append(run)
        self.launches_per_forward += 2 if defer_apply else 3  # conv, norm_finalize(, norm_apply): the library's own kernels
        if defer_apply:  # the consumer (head kernel / forward_logits) normalises on the fly
            self.final_norm = (ss, slope)
            return None
        return state

    def _add_tu(self, tu, src, dst, spatial_in):
        wp = self._pack_tu(tu, src.c, src.parts)
        w = tu.weight
        self.keep.append(wp)
        self._weight_slots.append(("tu", tu, (src.c, src.parts), (wp,)))
        d, h, wd = spatial_in
        plan = L.ConvPlan(kind=L.BSG_CONVT_K2S2, stride=1, N=self.batch, D=d, H=h, W=wd, cin=src.c, in_ptr=src.ptr(),
                          in_ctot=src.ctot, cout=w.shape[1], out_ptr=dst.buf.data_ptr(), out_ctot=dst.ctot,
                          out_coff=dst.coff, weights=wp.data_ptr(), bias=None, act=L.BSG_ACT_NONE, slope=0.0,
                          stats=None, out_f16=self.f16, in_f16=self.f16, use_khshift=0, max_ctas=0,
                          overflow=self._overflow_slot(), out_split_stride=w.shape[1] if self.split else 0,
                          tma_store={1: 1, 2: 2, 3: 1}.get(self.tma_store, 0))
        self.flops += plan.info().flops
        self.flops_algo += 2.0 * 8 * w.shape[0] * w.shape[1] * d * h * wd * self.batch
        self._note(f"convT2 {src.c}->{w.shape[1]} @{'x'.join(map(str, spatial_in))}", plan)
        self.steps.append(plan.run)
        self.launches_per_forward += 1

    def _pack_tu(self, tu, cin_pad, parts):
        w = tu.weight.detach().to(self.device, torch.float32)
        if parts is not None:
            w = P.split_k_weight(w, parts, cin_pad, 0)
        return P.pack_convT2_weight(w, cin_pad, self.act_dtype)

    def _note(self, name, plan):
        i = plan.info()
        self.plans.append(plan)
        self.step_info.append({"name": name, "flops": i.flops,
                               "plan": f"box {i.bw}x{i.bh}x{i.bd}x{i.bn} ntile {i.ntile}x{i.n_ntiles} cc {i.cc} "
                                       f"stages {i.nstages} khs {i.khshift} grid {i.grid}"})

    def _build(self):
        net = self.net
        num_pool = len(net.tu)
        in_ch = net.conv_blocks_context[0].blocks[0].conv.in_channels
        self.in_channels = in_ch
        m = 3 if self.split else 1  # physical channels per logical channel
        self.cin_pad = P.round_up(m * in_ch, 16)
        self.x = _Act(self._alloc(self.patch, self.cin_pad), 0, self.cin_pad, [(0, in_ch)] if self.split else None)
        cur, spatial = self.x, self.patch
        cats = []
        for d in range(num_pool + 1):
            stage = net.conv_blocks_context[d]
            blocks = list(stage.blocks) if d < num_pool else list(stage[0].blocks) + list(stage[1].blocks)
            prev = None  # norm state of the previous block of this stage: its output has exactly one consumer
            for i, blk in enumerate(blocks):
                cout, stride = blk.conv.out_channels, int(blk.conv.stride[0])
                if cout % 16:
                    raise NotImplementedError(f"channel width {cout} is not a multiple of 16")
                out_spatial = tuple(s // stride for s in spatial)
                one = [(0, cout)] if self.split else None
                if d < num_pool and i == len(blocks) - 1:
                    cat = self._alloc(out_spatial, 2 * m * cout)  # [0,C): transposed-conv output, [C,2C): this skip
                    cats.append(cat)
                    dst = _Act(cat, m * cout, m * cout, one)
                    is_skip = True
                else:
                    dst = _Act(self._alloc(out_spatial, m * cout), 0, m * cout, one)
                    is_skip = False
                prev = self._add_block(blk, cur, dst, spatial, claim=prev if i > 0 else None, first=(d == 0 and i == 0))
                if is_skip:
                    prev = None  # a skip tensor also feeds the decoder: it must be materialised
                cur, spatial = dst, out_spatial
        for u in range(num_pool):
            cat = cats[-(u + 1)]
            cskip = cat.shape[-1] // (2 * m)
            tu = net.tu[u]
            if tu.out_channels != cskip:
                raise NotImplementedError("transposed conv width differs from the skip width")
            self._add_tu(tu, cur, _Act(cat, 0, m * cskip, [(0, cskip)] if self.split else None), spatial)
            spatial = tuple(2 * s for s in spatial)
            cur = _Act(cat, 0, 2 * m * cskip, [(0, cskip), (3 * cskip, cskip)] if self.split else None)
            loc = net.conv_blocks_localization[u]
            blks = list(loc[0].blocks) + list(loc[1].blocks)
            prev = None
            for j, blk in enumerate(blks):
                co = blk.conv.out_channels
                dst = _Act(self._alloc(spatial, m * co), 0, m * co, [(0, co)] if self.split else None)
                # the very last block's norm + LeakyReLU is applied by its only consumer, the head
                last = (u == num_pool - 1) and (j == len(blks) - 1) and co <= 64 and not self.split
                prev = self._add_block(blk, cur, dst, spatial, defer_apply=last, claim=prev if j > 0 else None)
                cur = dst
        self.features = cur
        self._head = net.seg_outputs[num_pool - 1]
        self._load_head()
        self.num_classes = self._head.out_channels
        head_flops = 2.0 * self.num_classes * self.head_w.shape[1] * (self.patch[0] * self.patch[1] * self.patch[2])
        self.flops_per_item = self.flops / self.batch + head_flops            # as launched (input channels padded to 16)
        self.flops_algo_per_item = self.flops_algo / self.batch + head_flops  # algorithmic (965.5 / 3342.2 GF)

    def _load_head(self):
        head = self._head
        self.head_w = head.weight.detach().float().reshape(head.out_channels, -1).cpu().contiguous()
        self.head_b = head.bias.detach().float().cpu().contiguous() if head.bias is not None else None

    def reload_weights(self):
        """Re-packs the module tree's current parameters into the engine's EXISTING device tensors (plans and tensor
        maps point at fixed buffers, so they stay valid): what load_state_dict / load_checkpoint_ram need per fold,
        instead of rebuilding activation buffers, plans and tensor maps."""
        for kind, mod, (cin_pad, parts), tensors in self._weight_slots:
            if kind in ("block", "block_kw"):
                for dst, src in zip(tensors, self._pack_block(mod, cin_pad, kind == "block_kw", parts)):
                    if dst is not None:
                        dst.copy_(src)
            else:
                tensors[0].copy_(self._pack_tu(mod, cin_pad, parts))
        self._load_head()

    def close(self):
        """Drops plans, buffers and the step closures (which reference the engine): frees the device memory now
        instead of at the next cyclic garbage collection."""
        self.steps, self.plans, self.keep, self._weight_slots = [], [], [], []
        self.x = self.features = self._stats_arena = self._overflow = None

    # ------------------------------------------------------------------ execution
    def _carve_stats(self, cout):
        """(batch, cout, 2) fp64 slice of the statistics arena"""
        if self._stats_arena is None:
            self._stats_arena = torch.zeros(1 << 20, dtype=torch.float64, device=self.device)  # 8 MB: > 60 layers of 512 ch
        n = self.batch * cout * 2
        if self._stats_used + n > self._stats_arena.numel():
            raise L.BsgError("norm statistics arena exhausted")
        view = self._stats_arena[self._stats_used:self._stats_used + n].view(self.batch, cout, 2)
        self._stats_used += (n + 31) // 32 * 32  # 256-byte aligned slices
        return view

    def _run_steps(self, stream):
        self._in_run = True
        try:
            if self._stats_used:
                self._stats_arena[:self._stats_used].zero_()
            for step in self.steps:
                step(stream)
        finally:
            self._in_run = False

    def run(self, stream=None):
        """All conv / norm launches of one forward over the resident input batch `self.x`."""
        if self.event_log is None:
            self._run_steps(stream)
            return
        s = stream if stream is not None else torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        self._run_steps(stream)
        e1.record(s)
        self.event_log.append((e0, e1))

    def forward_logits(self, x):
        """(n <= batch, C, D, H, W) float tensor -> (n, num_classes, D, H, W) fp32 logits (head in torch: this
        convenience path backs `Generic_UNet.forward`; predict_3D uses the fused head kernel instead)."""
        n = x.shape[0]
        if n > self.batch or tuple(x.shape[2:]) != self.patch:
            raise ValueError("input does not match the engine geometry")
        self.x.buf.zero_()
        if self.split:
            xl = x.to(self.device, torch.float32).permute(0, 2, 3, 4, 1)
            hi = xl.to(torch.float16)
            c = self.in_channels
            self.x.buf[:n, ..., 0:c] = hi
            self.x.buf[:n, ..., c:2 * c] = hi
            self.x.buf[:n, ..., 2 * c:3 * c] = (xl - hi.float()).to(torch.float16)
        elif self.kwpack:
            self.x.buf[:n] = P.kwpack_input(x.to(self.device, torch.float32), self.cin_pad).to(self.act_dtype)
        else:
            self.x.buf[:n, ..., :self.in_channels] = x.to(self.device).permute(0, 2, 3, 4, 1).to(self.act_dtype)
        self.run()
        f = self.features.view()[:n].float()  # (n, d, h, w, c)
        if self.split:
            c = self.head_w.shape[1]
            f = f[..., 0:c] + f[..., 2 * c:3 * c]
        if self.final_norm is not None:
            ss, slope = self.final_norm
            f = torch.nn.functional.leaky_relu(f * ss[:n, :, 0].view(n, 1, 1, 1, -1) + ss[:n, :, 1].view(n, 1, 1, 1, -1), slope)
        logits = torch.einsum("ndhwc,kc->nkdhw", f, self.head_w.to(self.device))
        if self.head_b is not None:
            logits = logits + self.head_b.to(self.device).view(1, -1, 1, 1, 1)
        return logits
