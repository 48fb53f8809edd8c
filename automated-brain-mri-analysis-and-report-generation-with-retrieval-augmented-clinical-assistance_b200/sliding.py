"""Sliding-window driver: nnU-Net v1 `_internal_predict_3D_3Dconv_tiled` on the device (SURVEY.md Appendix A.3-A.6;
reference call site run_brats2021_inference_singlethread.py:97-106).

Host logic only — step grid, Gaussian importance map, work-item packing, accumulator management.  Every voxel
operation is a kernel of libbrainseg_b200.so.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L

# upstream order of the 8 mirror passes, as bit codes (bit0 = flip x / dim 4, bit1 = flip y / dim 3, bit2 = flip z / dim 2)
ALL_MIRROR_CODES = (0, 1, 2, 3, 4, 5, 6, 7)


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def compute_steps_for_sliding_window(patch_size, image_size, step_size):
    """nnU-Net v1 `_compute_steps_for_sliding_window` (App. A.3)."""
    assert all(i >= j for i, j in zip(image_size, patch_size)), "image size must be as large or larger than patch_size"
    assert 0 < step_size <= 1, "step_size must be larger than 0 and smaller or equal to 1"
    steps = []
    for img, patch in zip(image_size, patch_size):
        n = int(np.ceil((img - patch) / (patch * step_size))) + 1
        span = img - patch
        stride = span / (n - 1) if n > 1 else 99999999999
        steps.append([int(np.round(stride * k)) for k in range(n)])
    return steps


def mirror_codes_for(mirror_axes, do_mirroring=True):
    """Codes of the mirror passes upstream executes for `mirror_axes` (App. A.6): pass m runs iff all its axes are allowed."""
    if not do_mirroring:
        return [0]
    allowed = sum(1 << (2 - a) for a in set(mirror_axes))  # axis 0 (z) -> bit2, axis 1 (y) -> bit1, axis 2 (x) -> bit0
    return [m for m in ALL_MIRROR_CODES if (m & ~allowed) == 0]


def shard_work_items(num_tiles, mirror_codes, rank=0, world_size=1):
    """(tile index, mirror code) pairs owned by `rank`: round-robin over the items sorted by tile, so the shares
    differ by at most one forward and a rank's items of one tile stay adjacent (one gather / head launch per tile)."""
    items = [(t, m) for t in range(num_tiles) for m in mirror_codes]
    return items[rank::world_size]


def balanced_batch(items_per_rank, lanes=2, lo=4, hi=12):
    """Forwards in flight (lanes x per-lane batch) for a rank that owns `items_per_rank` work items per model: an engine
    always runs its whole batch, so a last chunk that is only partly filled is wasted work (18 items on two lanes of
    8: 16 + a chunk of 2 run as 8 = 25 % more forwards than needed).  Picks the per-lane batch in [lo, hi] that wastes the
    fewest forwards over the ceil(items / (lanes * b)) rounds, the larger batch on ties (the <= 8^3 levels amortise)."""
    best = None
    for b in range(lo, hi + 1):
        rounds = -(-items_per_rank // (lanes * b))
        # the last round fills lane after lane: only its last non-empty lane can be partly filled
        rem = items_per_rank - (rounds - 1) * lanes * b
        waste = -(-rem // b) * b - rem
        key = (waste, rounds, -b)
        if best is None or key < best[0]:
            best = (key, b)
    return lanes * best[1]


_gauss_cache = {}


def gaussian_importance_map(patch_size, device, sigma_scale=1.0 / 8):
    """nnU-Net v1 `_get_gaussian` (App. A.4) evaluated on the device in float64.

    scipy.ndimage.gaussian_filter of a centred delta is the outer product of three truncated (4 sigma), normalised 1-D
    Gaussians, multiplied axis by axis in the order the separable filter runs; the same float64 products are formed
    here, then / max, cast to float32, zeros replaced by the smallest non-zero value."""
    key = (tuple(patch_size), str(device), sigma_scale)
    if key not in _gauss_cache:
        ws = []
        for p in patch_size:
            # 1-D kernel exactly as scipy's _gaussian_kernel1d builds it (numpy float64 on the host: <= 2*4*sigma+1 taps)
            sigma = p * sigma_scale
            radius = int(4.0 * sigma + 0.5)
            x = np.arange(-radius, radius + 1)
            phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
            phi = phi / phi.sum()
            w = np.zeros(p, dtype=np.float64)
            for i in range(p):
                off = i - p // 2  # offset of voxel i from the delta
                if abs(off) <= radius:
                    w[i] = phi[off + radius]
            ws.append(torch.from_numpy(w).to(device))
        g = (ws[0][:, None, None] * ws[1][None, :, None]) * ws[2][None, None, :]
        g = (g / g.max() * 1).to(torch.float32)
        nz = g[g != 0]
        g[g == 0] = nz.min()
        _gauss_cache[key] = g.contiguous()
    return _gauss_cache[key]


class SlidingWindowPredictor:
    """Tiled prediction of one (C, Z, Y, X) volume.

    `engines`: one UNetEngine or a list of them ("lanes") sharing geometry and weights.  Every lane owns a CUDA stream:
    a chunk of work items is dealt over the lanes, whose gather / conv / norm launches then run concurrently — the
    HBM-bound passes of one lane (gather, norm apply, head) fill the gaps of the tensor-bound convs of the other.
    The accumulator updates (head kernels) are chained with events in a fixed order, so the fp32 sums are the same
    from run to run."""

    def __init__(self, engines, step_size=0.5, use_gaussian=True, mirror_codes=ALL_MIRROR_CODES, nonlin="sigmoid",
                 rank=0, world_size=1):
        self.engines = list(engines) if isinstance(engines, (list, tuple)) else [engines]
        self.engine = self.engines[0]
        engine = self.engine
        assert all(e.patch == engine.patch and e.device == engine.device for e in self.engines)
        self.patch = engine.patch
        self.step_size, self.use_gaussian = step_size, use_gaussian
        self.mirror_codes = list(mirror_codes)
        self.nonlin = {"sigmoid": 0, "softmax": 1, "identity": 2}[nonlin]
        self.rank, self.world_size = rank, world_size
        self.device = engine.device
        self._geom = {}
        self._streams = None
        self.kernel_launches = 0

    def _lane_streams(self):
        if self._streams is None:
            # high-priority streams: when the post-processing of the previous case (second stream of the pipeline) and the
            # forwards of this one both have blocks to place, the forwards go first
            prio = int(os.environ.get("BSG_LANE_PRIORITY", "-1"))
            self._streams = [torch.cuda.Stream(self.device, priority=prio) for _ in self.engines] \
                if len(self.engines) > 1 else [None]
        return self._streams

    def geometry(self, shape):
        """Step grid, Gaussian map and the (geometry-only) weight-sum volume for a padded volume shape."""
        shape = tuple(shape)
        if shape not in self._geom:
            steps = compute_steps_for_sliding_window(self.patch, shape, self.step_size)
            tiles = [(z, y, x) for z in steps[0] for y in steps[1] for x in steps[2]]
            gauss = None
            if self.use_gaussian and len(tiles) > 1:
                gauss = gaussian_importance_map(self.patch, self.device)
            add = gauss if gauss is not None else torch.ones(self.patch, dtype=torch.float32, device=self.device)
            wsum = torch.zeros(shape, dtype=torch.float32, device=self.device)
            p = self.patch
            for (z, y, x) in tiles:  # aggregated_nb_of_predictions[:, tile] += gaussian — same order, same fp32 sums
                wsum[z:z + p[0], y:y + p[1], x:x + p[2]] += add
            self._geom[shape] = (tiles, gauss, wsum)
        return self._geom[shape]

    def work_items(self, tiles):
        """(tile index, mirror code) pairs owned by this rank: round-robin over items sorted by tile."""
        return shard_work_items(len(tiles), self.mirror_codes, self.rank, self.world_size)

    def accumulate(self, vol, acc=None, stream=None):
        """vol: fp32 cuda tensor (C, Z, Y, X) with every extent >= patch.  Adds this rank's share of
        sum_tiles gauss * mean_mirrors(nonlin(net(flip(tile)))) into `acc` (fp32 [num_classes, Z, Y, X])."""
        eng0, lib = self.engine, L.lib()
        Cn, Z, Y, X = vol.shape
        tiles, gauss, _ = self.geometry((Z, Y, X))
        if acc is None:
            acc = torch.zeros((eng0.num_classes, Z, Y, X), dtype=torch.float32, device=self.device)
        p0, p1, p2 = self.patch
        pv = p0 * p1 * p2
        items = self.work_items(tiles)
        weight = 1.0 / len(self.mirror_codes)
        hw = eng0.head_w.numpy().ctypes.data_as(C.POINTER(C.c_float))
        hb = eng0.head_b.numpy().ctypes.data_as(C.POINTER(C.c_float)) if eng0.head_b is not None else None
        cfeat = eng0.head_w.shape[1]
        gptr = _ptr(gauss) if gauss is not None else None

        def nss(eng, first_item):  # deferred last-block norm of `eng`: scale/shift rows of the group's batch items
            if eng.final_norm is None:
                return None
            ss = eng.final_norm[0]
            return C.c_void_p(ss.data_ptr() + first_item * ss.shape[1] * 2 * 4)

        def nslope(eng):
            return float(eng.final_norm[1]) if eng.final_norm is not None else 0.0

        def groups_of(chunk):
            # consecutive items of the same tile: one gather / one head launch per group
            out, start = [], 0
            for i in range(1, len(chunk) + 1):
                if i == len(chunk) or chunk[i][0] != chunk[start][0]:
                    out.append((start, i))
                    start = i
            return out

        main = stream if stream is not None else torch.cuda.current_stream(self.device)
        lanes = list(zip(self.engines, self._lane_streams()))
        multi = len(lanes) > 1
        if multi:
            ready = torch.cuda.Event()
            ready.record(main)
            for _, s in lanes:
                s.wait_event(ready)  # vol / acc / geometry were produced on the caller's stream
        head_done = None  # event after the latest accumulator update: the head kernels form one ordered chain
        pos = 0
        while pos < len(items):
            for eng, s in lanes:
                chunk = items[pos:pos + eng.batch]
                pos += len(chunk)
                if not chunk:
                    break
                groups = groups_of(chunk)
                feat = eng.features
                with torch.cuda.stream(s if multi else main):
                    sp = L.stream_ptr(None)
                    for (a, b) in groups:
                        z, y, x = tiles[chunk[a][0]]
                        codes = (C.c_int * (b - a))(*[m for _, m in chunk[a:b]])
                        out = eng.x.buf.data_ptr() + 2 * a * pv * eng.x.ctot
                        L.check(lib.bsg_gather_patch_tta(_ptr(vol), Cn, Z, Y, X, z, y, x, p0, p1, p2, codes, b - a,
                                                         C.c_void_p(out), eng.x.ctot, eng.f16,
                                                         2 if eng.split else int(eng.kwpack), sp))
                    eng.run(None)
                    if multi and head_done is not None:
                        s.wait_event(head_done)
                    for (a, b) in groups:
                        z, y, x = tiles[chunk[a][0]]
                        codes = (C.c_int * (b - a))(*[m for _, m in chunk[a:b]])
                        fptr = feat.buf.data_ptr() + 2 * (a * pv * feat.ctot + feat.coff)
                        if eng.split:  # fp16x3 split features: the head reads hi + lo
                            L.check(lib.bsg_head_tta_accumulate_split(C.c_void_p(fptr), cfeat, feat.ctot, p0, p1, p2, codes,
                                                                      b - a, weight, hw, hb, eng.num_classes, self.nonlin,
                                                                      gptr, _ptr(acc), Z, Y, X, z, y, x, sp))
                            continue
                        L.check(lib.bsg_head_tta_accumulate(C.c_void_p(fptr), eng.f16, cfeat, feat.ctot, p0, p1, p2, codes, b - a,
                                                            weight, hw, hb, eng.num_classes, self.nonlin, gptr,
                                                            _ptr(acc), Z, Y, X, z, y, x, nss(eng, a), nslope(eng), sp))
                    if multi:
                        head_done = torch.cuda.Event()
                        head_done.record(s)
                self.kernel_launches += 2 * len(groups) + eng.launches_per_forward
        if multi:
            for _, s in lanes:  # join: everything the lanes did is ordered before what the caller does next
                done = torch.cuda.Event()
                done.record(s)
                main.wait_event(done)
        return acc

    def finalize(self, accs, shape, regions_class_order=None, want_probs=True, stream=None):
        """class_probabilities = acc / weight-sum (mean over `accs`), then argmax or ordered threshold -> uint8."""
        _, _, wsum = self.geometry(tuple(shape))
        ncls = accs[0].shape[0]
        nvox = wsum.numel()
        probs = torch.empty((ncls,) + tuple(shape), dtype=torch.float32, device=self.device) if want_probs else None
        seg = torch.empty(tuple(shape), dtype=torch.uint8, device=self.device)
        ptrs = (C.c_void_p * len(accs))(*[a.data_ptr() for a in accs])
        order = None
        mode = 0
        if regions_class_order is not None:
            mode = 1
            order = (C.c_int * ncls)(*[int(c) for c in regions_class_order])
        L.check(L.lib().bsg_finalize(ptrs, len(accs), _ptr(wsum), ncls, nvox, mode, order,
                                     _ptr(probs) if probs is not None else None, _ptr(seg), L.stream_ptr(stream)))
        self.kernel_launches += 1
        return seg, probs
