// tcgen05 / TMEM / TMA implicit-GEMM kernel for the 3-D conv stacks of Generic_UNet
// (reference: model_architecture/generic_UNet.py:56,69 Conv3d k3; :285-288 stride-2 conv pooling; :363-364
// ConvTranspose3d k2 s2).  Persistent, warp-specialised:
//   warps 0..3  epilogue       (tcgen05.ld -> bias / LeakyReLU / norm statistics -> 16-bit channels-last stores;
//   warps 6..9                  TMEM lane quadrant = warp id % 4; the two warps of a quadrant take alternate 32-column
//                               chunks of the N tile — the wide-N, short-K launches (transposed convs, the <= 8^3 levels)
//                               are bound by the epilogue's convert + store rate, not by the MMAs)
//   warp 4      TMA producer   (activation halo boxes + weight slabs -> swizzled smem ring)
//   warp 5      MMA issuer     (one elected lane, tcgen05.mma kind::f16, fp32 accumulators in TMEM, 2 buffers)
#pragma once
#include <cuda_fp16.h>
#include "bsg_common.cuh"
#include "bsg_ptx.cuh"
#include "conv_epilogue.cuh"
#include "conv_tc.cuh"

namespace bsg {

namespace {

constexpr int kThreads = 320;  // warps 0-3 + 6-9 epilogue, 4 producer, 5 MMA issuer
constexpr int kMaxStages = 12;

struct TileCoord {
    int nt;              // N tile
    int w0, h0, d0, n0;  // tile origin in the tile coordinate space
};

__device__ __forceinline__ TileCoord decode_tile(const ConvArgs& a, int tile) {
    TileCoord t;
    t.nt = tile % a.n_ntiles;
    int s = tile / a.n_ntiles;
    int iw = s % a.tw;
    s /= a.tw;
    int ih = s % a.th;
    s /= a.th;
    int id = s % a.td;
    s /= a.td;
    t.w0 = iw * a.bw;
    t.h0 = ih * a.bh;
    t.d0 = id * a.bd * a.mb;  // mb = 2: a work item is two M tiles, adjacent planes d0 and d0 + 1 (bd == 1 then)
    t.n0 = s * a.bn;
    return t;
}

// CC: channels per K chunk (16/32/64 <-> swizzle 32/64/128 B).
// MODE: kModeGeneric (one tap per stage, everything decided at run time), kModeKhs (kh halo reuse: 3 kh taps per
// stage), kModeS1 / kModeS2 (27 taps at stride 1 / 2, one tap per stage, no pair mode: the producer's tap loop is
// unrolled so that the parity view and the coordinate offsets of every tap are compile-time constants — ncu showed the
// generic producer thread spending ~105 dependent instructions = ~470 cycles per one-tap stage, 2..4x the stage's
// MMA time).
// kModeS1x3 / kModeS2x3: as kModeS1 / kModeS2 with the three kh taps of a (kd, kw) in ONE stage (three activation boxes,
// one 3-tap weight box): the layers whose one-tap stages are bound by the per-stage round trip rather than by MMA time
// (stride-2 32->64, the <= 8^3 levels) run a third of the stages.
constexpr int kModeGeneric = 0, kModeKhs = 1, kModeS2 = 2, kModeS1 = 3, kModeS2x3 = 4, kModeS1x3 = 5;
template <int CC, int MODE, int EPI, int MB>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvArgs a) {
    constexpr bool KHS = MODE == kModeKhs;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-B aligned carve-up (swizzle-128B atoms need it)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t stage_bytes = a.a_stage_bytes + a.b_stage_bytes;
    uint8_t* bar_area = smem + static_cast<size_t>(a.nstages) * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_area);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tfull_bar = empty_bar + kMaxStages;  // [2]
    uint64_t* tempty_bar = tfull_bar + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* sbias = reinterpret_cast<float*>(bar_area + 1024);  // [cout_pad] (<= 512), zeros when there is no bias
    // tma_out: two 4 KB staging buffers (32 voxels x 64 or 128 bytes) per epilogue warp, 1024-byte aligned (swizzle atoms)
    uint8_t* ostage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bar_area) + 3072 + 1023) & ~uintptr_t(1023));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&a.mapW);
        tma_prefetch_desc(&a.mapA[0]);
        if constexpr (EPI == kEpiStage) tma_prefetch_desc(&a.mapO[0]);
        for (int s = 0; s < a.nstages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], a.pair ? 2 : 1);  // pair mode: both CTAs' MMAs must have drained the stage
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 8);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 5) {
        tmem_alloc(tmem_slot, a.tmem_cols);
        tmem_relinquish();
    }
    if (warp < 4) {
        for (int i = threadIdx.x; i < a.cout_pad; i += 128) sbias[i] = (a.bias != nullptr) ? a.bias[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t crank = a.pair ? cluster_ctarank() : 0u;
    if (a.pair) cluster_sync_all();  // the peer's barriers are initialised before any multicast / remote arrive

    const int total_tiles = a.tn * a.td * a.th * a.tw * a.n_ntiles;
    // work items: plain mode = tiles, dealt round-robin to the CTAs; pair mode = (pair of neighbouring M tiles, N tile),
    // dealt to the clusters — both CTAs of a cluster walk the same item sequence, hence the same K-step sequence
    const int item0 = a.pair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int item_step = a.pair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int nitems = a.pair ? total_tiles / 2 : total_tiles;
    auto tile_of = [&](int item) {
        return a.pair ? (2 * (item / a.n_ntiles) + static_cast<int>(crank)) * a.n_ntiles + item % a.n_ntiles : item;
    };
    constexpr bool X3 = MODE == kModeS1x3 || MODE == kModeS2x3;
    const int ntg = (KHS || X3) ? 9 : a.ntaps;  // pipeline steps per chunk: (kd,kw) pairs or single taps
    const int ksteps = ntg * a.nchunks;
    constexpr int NKH = (KHS || X3) ? 3 : 1;
    constexpr uint32_t kRowBytes = CC * 2u;
    constexpr uint32_t kSbo = 8u * kRowBytes;
    const int nstages = a.nstages;

    if (warp == 4) {
        // =========================================================== TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            if constexpr (X3) {
                const uint32_t a_bytes = a.a_stage_bytes, a_tap_bytes = a.a_stage_bytes / 3, tx_bytes = a.stage_tx_bytes;
                const int nchunks = a.nchunks;
                for (int item = item0; item < nitems; item += item_step) {
                    const TileCoord t = decode_tile(a, tile_of(item));
                    const int nrow0 = t.nt * a.ntile;
#pragma unroll
                    for (int tg = 0; tg < 9; ++tg) {  // (kd, kw); the stage's taps are kh = 0, 1, 2
                        const int kd = tg / 3, kw = tg % 3;
                        constexpr bool S2 = MODE == kModeS2x3;
                        const int cw = S2 ? t.w0 - (kw == 0) : t.w0 + kw - 1;
                        const int cd = S2 ? t.d0 - (kd == 0) : t.d0 + kd - 1;
                        for (int c = 0; c < nchunks; ++c) {
                            mbar_wait(&empty_bar[stage], phase ^ 1u);
                            uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
                            mbar_expect_tx(&full_bar[stage], tx_bytes);
#pragma unroll
                            for (int kh = 0; kh < 3; ++kh) {
                                const int mi = S2 ? ((kw + 1) & 1) | (((kh + 1) & 1) << 1) | (((kd + 1) & 1) << 2) : 0;
                                const int ch = S2 ? t.h0 - (kh == 0) : t.h0 + kh - 1;
                                tma_load_5d(sa + kh * a_tap_bytes, &a.mapA[mi], &full_bar[stage], c * CC, cw, ch, cd, t.n0);
                            }
                            tma_load_3d(sa + a_bytes, &a.mapW, &full_bar[stage], c * CC, nrow0, tg * 3);
                            if (++stage == nstages) {
                                stage = 0;
                                phase ^= 1u;
                            }
                        }
                    }
                }
            } else if constexpr (MODE == kModeS2 || MODE == kModeS1) {
                const uint32_t a_bytes = a.a_stage_bytes, tx_bytes = a.stage_tx_bytes;
                const int nchunks = a.nchunks;
                for (int item = item0; item < nitems; item += item_step) {
                    const TileCoord t = decode_tile(a, tile_of(item));
                    const int nrow0 = t.nt * a.ntile;
#pragma unroll
                    for (int tap = 0; tap < 27; ++tap) {  // tap order (kd, kw, kh), kh fastest
                        const int kd = tap / 9, kw = (tap / 3) % 3, kh = tap % 3;
                        // stride 2: input index 2*o + k - 1: k=0 -> odd parity view, o-1; k=1 -> even, o; k=2 -> odd, o
                        constexpr bool S2 = MODE == kModeS2;
                        const int mi = S2 ? ((kw + 1) & 1) | (((kh + 1) & 1) << 1) | (((kd + 1) & 1) << 2) : 0;
                        const int cw = S2 ? t.w0 - (kw == 0) : t.w0 + kw - 1;
                        const int ch = S2 ? t.h0 - (kh == 0) : t.h0 + kh - 1;
                        const int cd = S2 ? t.d0 - (kd == 0) : t.d0 + kd - 1;
                        for (int c = 0; c < nchunks; ++c) {
                            mbar_wait(&empty_bar[stage], phase ^ 1u);
                            uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
                            mbar_expect_tx(&full_bar[stage], tx_bytes);
                            tma_load_5d(sa, &a.mapA[mi], &full_bar[stage], c * CC, cw, ch, cd, t.n0);
                            if (a.pair) {  // my half of the weight rows, to both CTAs of the pair
                                const int half = a.ntile >> 1, r0 = static_cast<int>(crank) * half;
                                tma_load_3d_mc(sa + a_bytes + r0 * kRowBytes, &a.mapWh, &full_bar[stage], c * CC, nrow0 + r0,
                                               tap, static_cast<uint16_t>(3));
                            } else {
                                tma_load_3d(sa + a_bytes, &a.mapW, &full_bar[stage], c * CC, nrow0, tap);
                            }
                            if (++stage == nstages) {
                                stage = 0;
                                phase ^= 1u;
                            }
                        }
                    }
                }
            } else
            for (int item = item0; item < nitems; item += item_step) {
                const TileCoord t = decode_tile(a, tile_of(item));
                const int nrow0 = t.nt * a.ntile;
                int kd = 0, kw = 0, kh = 0;  // tap order (kd, kw, kh): kh fastest, absent when KHS
                for (int tg = 0; tg < ntg; ++tg) {
                    int cw, ch, cd, mi = 0, tap;
                    if (a.ntaps == 1) {
                        cw = t.w0;
                        ch = t.h0;
                        cd = t.d0;
                        tap = 0;
                    } else if (a.stride == 1) {
                        cw = t.w0 + kw - 1;
                        ch = t.h0 + kh - 1;
                        cd = t.d0 + kd - 1;
                        tap = KHS ? tg * 3 : tg;
                    } else {
                        // input index 2*o + k - 1: k=0 -> odd parity, o-1; k=1 -> even parity, o; k=2 -> odd parity, o
                        const int pw = (kw + 1) & 1, ph = (kh + 1) & 1, pd = (kd + 1) & 1;
                        mi = pw | (ph << 1) | (pd << 2);
                        cw = t.w0 - (kw == 0);
                        ch = t.h0 - (kh == 0);
                        cd = t.d0 - (kd == 0);
                        tap = tg;
                    }
                    const CUtensorMap* mapA = &a.mapA[mi];
                    for (int c = 0; c < a.nchunks; ++c) {
                        mbar_wait(&empty_bar[stage], phase ^ 1u);
                        uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
                        uint8_t* sb = sa + a.a_stage_bytes;
                        mbar_expect_tx(&full_bar[stage], a.stage_tx_bytes);
                        tma_load_5d(sa, mapA, &full_bar[stage], c * CC, cw, ch, cd, t.n0);
                        if (a.pair) {
                            // my half of the N rows of every tap of the stage, to both CTAs (same smem offset)
                            const int half = a.ntile >> 1;
                            for (int j = 0; j < NKH; ++j)
                                tma_load_3d_mc(sb + (j * a.ntile + static_cast<int>(crank) * half) * kRowBytes, &a.mapWh,
                                               &full_bar[stage], c * CC, nrow0 + static_cast<int>(crank) * half, tap + j,
                                               static_cast<uint16_t>(3));
                        } else {
                            tma_load_3d(sb, &a.mapW, &full_bar[stage], c * CC, nrow0, tap);
                        }
                        if (++stage == nstages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    if (KHS) {
                        if (++kw == 3) {
                            kw = 0;
                            ++kd;
                        }
                    } else if (++kh == 3) {
                        kh = 0;
                        if (++kw == 3) {
                            kw = 0;
                            ++kd;
                        }
                    }
                }
            }
        }
    } else if (warp == 5) {
        // =========================================================== MMA issuer (one elected thread runs the whole
        // role: inside elect.sync the compiler emits straight UTCHMMA sequences without per-stage reconvergence waits)
        if (elect_one()) {
            const uint32_t idesc = make_idesc_16(128, static_cast<uint32_t>(a.ntile), a.in_f16);
            constexpr uint32_t kLayout = (CC == 64) ? kLayoutSW128 : (CC == 32 ? kLayoutSW64 : kLayoutSW32);
            // descriptor = constant high word | (start address >> 4): only the low word moves between MMAs
            const uint64_t desc_base = make_smem_desc(0, kSbo, kLayout);
            const uint32_t desc_hi = static_cast<uint32_t>(desc_base >> 32);
            const uint32_t b_tap16 = (static_cast<uint32_t>(a.ntile) * kRowBytes) >> 4;
            const uint32_t smem0_16 = static_cast<uint32_t>(desc_base) + (smem_u32(smem) >> 4);  // LBO field + address
            const uint32_t stage16 = stage_bytes >> 4, a16 = a.a_stage_bytes >> 4;
            // distance between the A operands of consecutive kh taps inside a stage: one row group of the haloed box
            // (KHS) or one whole 128-row box (three boxes per stage)
            const uint32_t a_kh16 = X3 ? (a.a_stage_bytes / 3) >> 4 : kSbo >> 4;
            // M blocking: plane d0 + 1 follows plane d0 inside every activation box ((bh [+ 2]) x bw rows further)
            const uint32_t a_m16 = (static_cast<uint32_t>((a.bh + (KHS ? 2 : 0)) * a.bw) * kRowBytes) >> 4;
            int stage = 0;
            uint32_t phase = 0;
            uint32_t tcount = 0;
            bool ready = false;  // next stage's full barrier already seen complete (probed early, see below)
            for (int item = item0; item < nitems; item += item_step, ++tcount) {
                const uint32_t acc = tcount & 1u;
                const uint32_t acc_phase = (tcount >> 1) & 1u;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * static_cast<uint32_t>(a.ntile * MB);
                for (int ks = 0; ks < ksteps; ++ks) {
                    if (!ready) mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa16 = smem0_16 + static_cast<uint32_t>(stage) * stage16;
                    const uint32_t sb16 = sa16 + a16;
                    const int nstage = (stage + 1 == nstages) ? 0 : stage + 1;
                    const uint32_t nphase = (stage + 1 == nstages) ? (phase ^ 1u) : phase;
#pragma unroll
                    for (int kh = 0; kh < NKH; ++kh) {
#pragma unroll
                        for (int k = 0; k < CC / 16; ++k) {
                            const uint32_t ad = sa16 + kh * a_kh16 + ((k * 32) >> 4);
                            const uint32_t bd = sb16 + kh * b_tap16 + ((k * 32) >> 4);
                            umma_bf16_lo(d_tmem, ad, bd, desc_hi, idesc, (kh | k) != 0 ? 1u : (ks != 0 ? 1u : 0u));
                            // M blocking: the second M tile (the next plane of the same activation box) against the
                            // SAME weight operand — half the weight bytes per MMA cycle through shared memory and L2
                            if constexpr (MB == 2)
                                umma_bf16_lo(d_tmem + static_cast<uint32_t>(a.ntile), ad + a_m16, bd, desc_hi, idesc,
                                             (kh | k) != 0 ? 1u : (ks != 0 ? 1u : 0u));
                            // probe the next stage's barrier behind the first MMA: its latency overlaps queued work
                            if (kh == 0 && k == 0) ready = mbar_try_wait(&full_bar[nstage], nphase);
                        }
                    }
                    // frees the smem slot once these MMAs have read it (pair mode: in both CTAs — either producer
                    // writes weight rows into both)
                    if (a.pair)
                        umma_commit_mc(&empty_bar[stage], static_cast<uint16_t>(3));
                    else
                        umma_commit(&empty_bar[stage]);
                    if (ks == ksteps - 1) umma_commit(&tfull_bar[acc]);
                    stage = nstage;
                    phase = nphase;
                }
            }
        }
        __syncwarp();
    } else {
        // =========================================================== epilogue (4 warps, one TMEM lane quadrant each)
        const int q = warp & 3;
        const int half = warp >= 6 ? 1 : 0;  // which of the quadrant's two epilogue warps: odd / even chunks
        const int row = q * 32 + lane;
        int r = row;
        const int iw = r % a.bw;
        r /= a.bw;
        const int ih = r % a.bh;
        r /= a.bh;
        const int id = r % a.bd;
        r /= a.bd;
        const int in = r;
        EpiParams epi;
        epi.sbias = sbias;
        epi.has_bias = a.bias != nullptr;
        epi.stats = a.stats;
        epi.cout = a.cout;
        epi.No = a.No;
        epi.act = a.act;
        epi.slope = a.slope;
        epi.out_f16 = a.out_f16;
        epi.guard = (a.overflow != nullptr && a.out_f16) ? 1 : 0;
        epi.split_stride = a.split_stride;
        epi.stage = nullptr;
        epi.stage_wide = epi.stage_sub = 0;
        // tma_out: this warp's 32 rows are the sub-box (bw, 32/bw .. ) of the tile at these offsets; its stores rotate
        // through two staging buffers, lane 0 issues and tracks them
        const int ewarp = q + 4 * half;
        uint8_t* const my_stage = ostage + ewarp * 8192;
        const int sub_h = ((q * 32) / a.bw) % a.bh, sub_d = ((q * 32) / (a.bw * a.bh)) % a.bd,
                  sub_n = (q * 32) / (a.bw * a.bh * a.bd);
        uint32_t nstore = 0;
        EpiGuard guard;
        guard.init();
        // running norm statistics of the (up to 8) 32-column chunks of the current N tile, flushed when the batch item
        // of this warp's rows or the N tile changes.  The rows of one warp share the batch index as long as a tile holds
        // >= 32 voxels per item (bn <= 4); smaller boxes take the grouped path (stats_chunk_grouped).
        const int vox_per_item = a.bw * a.bh * a.bd;
        const bool grouped = a.stats != nullptr && vox_per_item < 32;
        if (grouped) epi.stats = nullptr;  // the chunk body then skips the statistics; the grouped path gets a.stats
        StatAcc sacc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) sacc[j].s1 = sacc[j].s2 = 0.f;
        int stat_n = -1, stat_nt = 0;
        float unused1[32], unused2[32];  // per-thread statistic sums: brick kernel only
        uint32_t tcount = 0;
        for (int item = item0; item < nitems; item += item_step, ++tcount) {
            const TileCoord t = decode_tile(a, tile_of(item));
            const uint32_t acc = tcount & 1u;
            const uint32_t acc_phase = (tcount >> 1) & 1u;
            const int w = t.w0 + iw, h = t.h0 + ih, n = t.n0 + in;
            if (a.stats != nullptr && !grouped) {
                const int n_warp = __shfl_sync(0xffffffffu, n, 0);
                if (n_warp != stat_n || t.nt != stat_nt) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) flush_stats(epi, sacc[j], stat_nt * a.ntile + j * 32, lane, stat_n);
                    stat_n = n_warp;
                    stat_nt = t.nt;
                }
            }
            const int q0 = t.nt * a.ntile;  // first GEMM column of this tile
            __nv_bfloat16* obase = a.out + n * a.os_n + a.out_c_off;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int mt = 0; mt < MB; ++mt) {  // the item's M tiles (M blocking: planes d0 and d0 + 1)
            const int d = t.d0 + id + mt;
            const bool valid = (w < a.Wo) && (h < a.Ho) && (d < a.Do) && (n < a.No);
            __nv_bfloat16* orow = obase + static_cast<long long>(d) * a.os_d + static_cast<long long>(h) * a.os_h +
                                  static_cast<long long>(w) * a.os_w;
            const uint32_t t_addr = tmem_base + (acc * static_cast<uint32_t>(MB) + static_cast<uint32_t>(mt)) * static_cast<uint32_t>(a.ntile) +
                                    (static_cast<uint32_t>(q * 32) << 16);
            // This warp's chunks: half, half + 2, ...  ONE copy of the chunk body in the instruction stream (a runtime
            // loop): unrolled over the 8 chunk positions it was ~110 KB of SASS per warp flavour and the epilogue
            // warps spent a quarter of their time waiting for instruction fetches (ncu: stall_no_inst, transposed conv).
            // The running statistics stay in registers: the chunk's result is added to sacc[j] by a predicated,
            // unrolled select instead of a dynamic index.  (Issuing the next chunk's tcgen05.ld ahead of this chunk's
            // conversion + stores was tried on top: slower, 0.613 -> 0.682 ms on the 64->64 transposed conv.)
            const int nchunk = a.ntile >> 5;
            int par = 0, cpar = 0;
            if (a.out_mul == 2) {  // (parity, channel) of this warp's first chunk
                par = q0 / a.cout_pad;
                cpar = q0 - par * a.cout_pad;
                if (half && (cpar += 32) >= a.cout_pad) {
                    cpar = 0;
                    ++par;
                }
            }
            if constexpr (EPI == kEpiStage) {
                // Staged route: the warp's unit is one store row block — 32 voxels x `store_cols` (32 or 64) channels of ONE
                // output parity = rows of 64 or 128 bytes (whole lines when the layer has >= 64 output channels), written
                // by one TMA tensor store.  The two warps of a quadrant take alternate units.
                const int wide = a.store_cols >> 5;  // 32-column chunks per unit
                const int nunits = nchunk / wide;
#pragma unroll 1
                for (int u = half; u < nunits; u += 2) {
                    const int col0 = q0 + u * a.store_cols;  // first GEMM column of the unit
                    int upar = 0, co0 = col0;
                    if (a.out_mul == 2) {
                        upar = col0 / a.cout_pad;
                        co0 = col0 - upar * a.cout_pad;
                    }
                    // the buffer about to be overwritten was handed to the store before last: at most one (the other
                    // buffer's) may still be reading
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
                    epi.stage = my_stage + (nstore & 1u) * 4096u;
                    epi.stage_wide = wide - 1;
#pragma unroll 1
                    for (int sub = 0; sub < wide; ++sub) {
                        const int j = u * wide + sub;
                        uint32_t v[32];
                        tmem_ld_32x32(t_addr + j * 32, v);
                        tmem_ld_wait();
                        StatAcc chunk_stats;
                        chunk_stats.s1 = chunk_stats.s2 = 0.f;
                        epi.stage_sub = sub;
                        epilogue_32cols<false, EPI>(v, epi, co0 + sub * 32, valid, lane, chunk_stats, orow, unused1, unused2, guard);
                        if (grouped) {
                            stats_chunk_grouped(t_addr + j * 32, sbias, a.bias != nullptr, a.stats, a.cout, a.No, co0 + sub * 32,
                                                valid, lane, vox_per_item, t.n0 + (q * 32) / vox_per_item);
                        } else if (a.stats != nullptr) {
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                if (k == j) {
                                    sacc[k].s1 += chunk_stats.s1;
                                    sacc[k].s2 += chunk_stats.s2;
                                }
                        }
                    }
                    fence_proxy_async();  // every lane's staging writes -> visible to the async proxy
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_5d(&a.mapO[upar], epi.stage, co0, t.w0, t.h0 + sub_h, t.d0 + sub_d + mt, t.n0 + sub_n);
                        tma_store_commit();
                    }
                    ++nstore;
                }
            } else
#pragma unroll 1
            for (int j = half; j < nchunk; j += 2) {
                const int cb = j * 32;
                uint32_t v[32];
                tmem_ld_32x32(t_addr + cb, v);
                tmem_ld_wait();
                int co = q0 + cb;
                if (a.out_mul == 2) {
                    // transposed conv: GEMM columns enumerate (parity (pd, ph, pw), channel); an N tile may span
                    // several parities, so the output voxel is re-derived per 32-column chunk (no division: the
                    // parity / channel pair is advanced chunk by chunk)
                    co = cpar;
                    orow = obase + static_cast<long long>(2 * d + ((par >> 2) & 1)) * a.os_d +
                           static_cast<long long>(2 * h + ((par >> 1) & 1)) * a.os_h +
                           static_cast<long long>(2 * w + (par & 1)) * a.os_w;
#pragma unroll
                    for (int step = 0; step < 2; ++step)  // on to this warp's next chunk: two positions further
                        if ((cpar += 32) >= a.cout_pad) {
                            cpar = 0;
                            ++par;
                        }
                }
                StatAcc chunk_stats;
                chunk_stats.s1 = chunk_stats.s2 = 0.f;
                epilogue_32cols<false, EPI>(v, epi, co, valid, lane, chunk_stats, orow, unused1, unused2, guard);
                if (grouped) {
                    stats_chunk_grouped(t_addr + cb, sbias, a.bias != nullptr, a.stats, a.cout, a.No, co, valid, lane, vox_per_item,
                                        t.n0 + (q * 32) / vox_per_item);
                } else if (a.stats != nullptr) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k == j) {
                            sacc[k].s1 += chunk_stats.s1;
                            sacc[k].s2 += chunk_stats.s2;
                        }
                }
            }
            }  // mt
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        if (epi.stats != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) flush_stats(epi, sacc[j], stat_nt * a.ntile + j * 32, lane, stat_n);
        }
        if (epi.guard) guard.flush(a.overflow);
        if (EPI == kEpiStage && lane == 0) tma_store_wait_all();  // the staging buffers must outlive the stores reading them
    }

    tc_fence_before();
    __syncthreads();
    if (a.pair) cluster_sync_all();  // no CTA leaves while its peer can still multicast into it / arrive on its barriers
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

template <int CC, int MODE, int EPI, int MB>
static cudaError_t launch_variant(const ConvArgs& a, int grid, size_t smem_bytes, cudaStream_t stream) {
    static unsigned long long attr_done = 0;  // per device
    if (cudaError_t e = ensure_max_smem(conv_tc_kernel<CC, MODE, EPI, MB>, &attr_done, 232448); e != cudaSuccess) return e;
    if (a.pair) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(static_cast<unsigned>(grid));
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, conv_tc_kernel<CC, MODE, EPI, MB>, a);
    }
    conv_tc_kernel<CC, MODE, EPI, MB><<<grid, kThreads, smem_bytes, stream>>>(a);
    return cudaGetLastError();
}

// MB = 2 (M blocking) exists for the tap-table modes only: the planner never pairs it with the generic mode.
template <int EPI, int MB = 1>
static cudaError_t launch_modes(const ConvArgs& a, int grid, size_t smem_bytes, cudaStream_t stream) {
    if (a.khshift) {
        if (a.cc == 64) return launch_variant<64, kModeKhs, EPI, MB>(a, grid, smem_bytes, stream);
        if (a.cc == 32) return launch_variant<32, kModeKhs, EPI, MB>(a, grid, smem_bytes, stream);
        return launch_variant<16, kModeKhs, EPI, MB>(a, grid, smem_bytes, stream);
    }
    if (a.taps3 && a.ntaps == 27 && !a.pair) {
        if (a.stride == 2) {
            if (a.cc == 64) return launch_variant<64, kModeS2x3, EPI, MB>(a, grid, smem_bytes, stream);
            if (a.cc == 32) return launch_variant<32, kModeS2x3, EPI, MB>(a, grid, smem_bytes, stream);
            return launch_variant<16, kModeS2x3, EPI, MB>(a, grid, smem_bytes, stream);
        }
        if (a.cc == 64) return launch_variant<64, kModeS1x3, EPI, MB>(a, grid, smem_bytes, stream);
        if (a.cc == 32) return launch_variant<32, kModeS1x3, EPI, MB>(a, grid, smem_bytes, stream);
        return launch_variant<16, kModeS1x3, EPI, MB>(a, grid, smem_bytes, stream);
    }
    if (a.stride == 2 && a.ntaps == 27) {
        if (a.cc == 64) return launch_variant<64, kModeS2, EPI, MB>(a, grid, smem_bytes, stream);
        if (a.cc == 32) return launch_variant<32, kModeS2, EPI, MB>(a, grid, smem_bytes, stream);
        return launch_variant<16, kModeS2, EPI, MB>(a, grid, smem_bytes, stream);
    }
    if (a.stride == 1 && a.ntaps == 27) {
        if (a.cc == 64) return launch_variant<64, kModeS1, EPI, MB>(a, grid, smem_bytes, stream);
        if (a.cc == 32) return launch_variant<32, kModeS1, EPI, MB>(a, grid, smem_bytes, stream);
        return launch_variant<16, kModeS1, EPI, MB>(a, grid, smem_bytes, stream);
    }
    if constexpr (MB == 2) {
        return cudaErrorInvalidConfiguration;
    } else {
        if (a.cc == 64) return launch_variant<64, kModeGeneric, EPI, MB>(a, grid, smem_bytes, stream);
        if (a.cc == 32) return launch_variant<32, kModeGeneric, EPI, MB>(a, grid, smem_bytes, stream);
        return launch_variant<16, kModeGeneric, EPI, MB>(a, grid, smem_bytes, stream);
    }
}

}  // namespace
}  // namespace bsg
