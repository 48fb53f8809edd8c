// "Brick" tcgen05 / TMEM / TMA implicit-GEMM kernel for the stride-1 3x3x3 convs with Cout <= 64
// (reference: model_architecture/generic_UNet.py:56,69 — the full- and half-resolution conv blocks, where the
// activations are large and the channel counts small).
//
// Measured on B200 (profiles/r01_probe_umma_issue_rates.log): a tcgen05.mma with M = 128 costs max(N/2, ~45) cycles,
// whether A comes from shared memory or TMEM — at N = Cout = 32 the tensor pipe can never be more than 36 % busy, and
// the plain tile kernel (conv_tc.cu) additionally re-reads every activation box 9x and all 27 weight taps per 128
// output voxels.  Here a CTA owns a brick of P output planes (8 w x 16 h voxels each) whose fp32 accumulators sit side
// by side in TMEM, plane q at column (P-1-q)*NT, so that
//   * the three kd taps of one (kh, kw, channel step) become ONE MMA of N = 3*NT: input plane p feeds output planes
//     p, p-1, p-2 through weights [W(kd=0) | W(kd=1) | W(kd=2)] — N = 96 / 192 instead of 32 / 64;
//   * one activation box (8 w x 18 h haloed rows of ONE input plane, one kw shift, one channel chunk) is loaded once
//     and serves the nine (kd, kh) taps it takes part in (kh through row-shifted descriptors);
//   * weights are staged as per-phase slabs (phase = (channel chunk, kw): 9 taps x NT x CC, laid out [kh][kd][NT][CC])
//     that stay resident in shared memory for the whole launch when all phases fit, else stream through two buffers.
// kw-fused variant (KWF, used whenever the slabs are resident): the box is 10 w x 18 h and every (kh, kw) tap reads it
// through a descriptor that starts at row kh*10 + kw with a 10-row group pitch — start addresses that are not
// aligned to the 8-row swizzle atom are fine because the tensor core applies the swizzle to absolute shared-memory
// address bits (profiles/r01_probe_umma_shift.log).  One TMA box then feeds 27 taps instead of 9, which matters
// because the box fill and the MMA operand reads share the same 128 B/clk shared-memory port.
// Roles: warps 0..3 epilogue (TMEM lane quadrant = warp id), warp 4 activation TMA producer, warp 5 weight-slab TMA
// producer, warp 6 MMA issuer (one thread); CC = 16 instantiations add a second epilogue group on warps 7..10 (the
// groups take alternate planes).  With one group the issuer is the highest warp id of its scheduler partition: the
// warp arbiter favours the highest id, so the latency-critical tcgen05.mma stream is never queued behind the
// instruction-heavy epilogue warp it shares the partition with.
#pragma once
#include <cuda_fp16.h>
#include "bsg_common.cuh"
#include "bsg_ptx.cuh"
#include "conv_brick.cuh"
#include "conv_epilogue.cuh"

namespace bsg {

namespace {

constexpr int kMaxStages = 12;
constexpr int kMaxSlabs = 6;
constexpr int kMaxAcc = 16;

struct Unit {
    int w0, h0, d0, n;
};

__device__ __forceinline__ Unit decode_unit(const BrickArgs& a, int u, int P) {
    Unit t;
    const int b = u % a.tb;
    int s = u / a.tb;
    const int wt = s % a.tw;
    s /= a.tw;
    const int ht = s % a.th;
    t.n = s / a.th;
    t.w0 = wt * 8;
    t.h0 = ht * 16;
    t.d0 = b * P;
    return t;
}

// STATS: the epilogue also accumulates the norm statistics (a.stats != null) — a separate instantiation because the
// per-thread sums cost 2 * NT registers.
//
// CC == 16 (the 4-channel network input, one K chunk): a plane takes only 9 MMAs, so one epilogue warp group
// (TMEM -> bias/stats/activation -> global) is the bottleneck; those instantiations run a second group on warps
// 7..10 and the groups take alternate planes.  With 352 threads the per-thread register budget is 186, so the
// NT = 64 statistics there are reduced per tile (shuffles) instead of kept as 128 per-thread sums.
//
// XF (a.in_norm != null; CC >= 32): the INPUT tensor is the raw output of an InstanceNorm / GroupNorm conv block whose
// normalise + LeakyReLU pass (bsg_norm_apply_lrelu: a full extra read + write of the activation in HBM) was skipped.
// Warps 7..10 apply y = lrelu(x * scale[n][c] + shift[n][c]) to every activation box in place in shared memory, between
// the TMA landing (full barrier) and the MMAs (xfull barrier).  The pass is address-based: a thread keeps one 16-byte
// column of the 2 KB it strides by, and because TMA's swizzle is a function of shared-memory address bits alone, that
// column always holds the same 8 channels of the K chunk — its 8 (scale, shift, slope) triples live in registers.
// Conv padding: the activation map of an XF plan is encoded with NaN out-of-bounds fill, so padding arrives as NaN
// and is written back as 0 (zeros in y-space, as the reference pads AFTER the norm); no box geometry needed.
template <int CC, int NT, bool STATS, bool KWF, bool XF>
__global__ void __launch_bounds__(brick_threads(CC, NT, STATS, XF), 1) conv_brick_kernel(const __grid_constant__ BrickArgs a) {
    static_assert(!(XF && CC == 16), "the in-consumer norm transform needs CC >= 32");
    constexpr int P = 256 / NT;
    constexpr int kEpiGroups = brick_epi_groups(CC, NT, STATS);
    // 352-thread instantiations (second epilogue group, or the XF transform warps) have 186 registers per thread: the
    // NT = 64 statistics are then reduced per tile (shuffles) instead of kept as 128 per-thread sums
    // column split: the two epilogue groups share every plane (32 columns each) instead of alternating planes
    constexpr bool kColSplit = kEpiGroups == 2 && NT == 64 && STATS;
    constexpr int kChunks = kColSplit ? 1 : NT / 32;  // 32-column chunks one epilogue warp handles per plane
    constexpr bool kThreadAcc = !(XF && NT == 64);
    constexpr uint32_t kRowBytes = CC * 2u;
    constexpr uint32_t kAtom = 8u * kRowBytes;              // 8 rows: one swizzle atom, one h step of the 8-wide box
    constexpr uint32_t kTapBytes = NT * kRowBytes;          // one tap of a weight slab
    constexpr uint32_t kLayout = (CC == 64) ? kLayoutSW128 : (CC == 32 ? kLayoutSW64 : kLayoutSW32);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* slabs = smem;
    uint8_t* stages = slabs + static_cast<size_t>(a.nslabbuf) * a.slab_bytes;
    uint8_t* bar_area = stages + static_cast<size_t>(a.nstages) * a.a_stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_area);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* wfull_bar = empty_bar + kMaxStages;
    uint64_t* wempty_bar = wfull_bar + kMaxSlabs;
    uint64_t* tfull_bar = wempty_bar + kMaxSlabs;
    uint64_t* tempty_bar = tfull_bar + kMaxAcc;
    uint64_t* xfull_bar = tempty_bar + kMaxAcc;  // [kMaxStages] XF: stage transformed, MMAs may read it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfull_bar + kMaxStages);
    float* sbias = reinterpret_cast<float*>(bar_area + 1024);  // [NT]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&a.mapW);
        tma_prefetch_desc(&a.mapA);
        for (int s = 0; s < kMaxStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&xfull_bar[s], 4);  // one arrive per transform warp
        }
        for (int s = 0; s < kMaxSlabs; ++s) {
            mbar_init(&wfull_bar[s], 1);
            mbar_init(&wempty_bar[s], 1);
        }
        for (int i = 0; i < kMaxAcc; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], kColSplit ? 8 : 4);  // one arrive per epilogue warp that reads the plane
        }
        fence_barrier_init();
    }
    if (warp == 6) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    if (warp < 4) {
        for (int i = threadIdx.x; i < NT; i += 128) sbias[i] = (a.bias != nullptr) ? a.bias[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int units = a.tn * a.th * a.tw * a.tb;
    const int nphases = a.nphases;
    const bool resident = a.nslabbuf >= nphases;

    if (warp == 4) {
        // =========================================================== activation producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const Unit t = decode_unit(a, u, P);
                const int nloads = KWF ? a.nchunks : nphases;  // KWF: one haloed box per (chunk, plane) for all kw
                for (int ph = 0; ph < nloads; ++ph) {
                    const int c = KWF ? ph : ph / a.kwn, kw = (KWF || a.kwn == 1) ? 1 : ph - c * 3;
                    for (int p = 0; p < P + 2; ++p) {
                        const int d = t.d0 + p - 1;  // d = -1 / D: the box is all out of bounds -> zeros (conv padding)
                        mbar_wait(&empty_bar[stage], phase ^ 1u);
                        mbar_expect_tx(&full_bar[stage], a.a_tx_bytes);
                        tma_load_5d(stages + static_cast<size_t>(stage) * a.a_stage_bytes, &a.mapA, &full_bar[stage],
                                    c * CC, t.w0 + kw - 1 - (KWF ? 1 : 0), t.h0 - 1, d, t.n);
                        if (++stage == a.nstages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 5) {
        // =========================================================== weight-slab producer
        if (elect_one()) {
            uint32_t su = 0;  // slab uses so far
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                if (resident && su >= static_cast<uint32_t>(nphases)) break;
                for (int ph = 0; ph < nphases; ++ph, ++su) {
                    uint32_t buf;
                    if (resident) {
                        buf = static_cast<uint32_t>(ph);
                    } else {
                        buf = su % static_cast<uint32_t>(a.nslabbuf);
                        mbar_wait(&wempty_bar[buf], ((su / static_cast<uint32_t>(a.nslabbuf)) & 1u) ^ 1u);
                    }
                    const int c = ph / a.kwn, kw = ph - c * a.kwn;
                    uint8_t* dst = slabs + static_cast<size_t>(buf) * a.slab_bytes;
                    mbar_expect_tx(&wfull_bar[buf], a.slab_bytes);
                    for (int kh = 0; kh < 3; ++kh)  // box (CC, NT, 1, 1, 3 kd) -> [kd][NT][CC] behind each kh
                        tma_load_5d(dst + kh * 3 * kTapBytes, &a.mapW, &wfull_bar[buf], c * CC, 0, kh, kw, 0);
                }
            }
        }
    } else if (warp == 6) {
        // =========================================================== MMA issuer (one elected thread runs the whole
        // role: inside elect.sync the compiler knows the code is warp-uniform and emits straight UTCHMMA sequences)
        if (elect_one()) {
            const uint32_t idesc1 = make_idesc_16(128, NT, a.in_f16), idesc2 = make_idesc_16(128, 2 * NT, a.in_f16),
                           idesc3 = make_idesc_16(128, 3 * NT, a.in_f16);
            const uint64_t desc_base = make_smem_desc(0, kAtom, kLayout);
            const uint32_t desc_hi = static_cast<uint32_t>(desc_base >> 32);
            const uint32_t desc_lo0 = static_cast<uint32_t>(desc_base);  // LBO field; the start address is added to it
            const uint32_t stages16 = desc_lo0 + (smem_u32(stages) >> 4), slabs16 = desc_lo0 + (smem_u32(slabs) >> 4);
            const uint32_t stage16 = a.a_stage_bytes >> 4, slab16 = a.slab_bytes >> 4;
            constexpr uint32_t kTap16 = kTapBytes >> 4, kAtom16 = kAtom >> 4;
            uint64_t* const ready_bar = XF ? xfull_bar : full_bar;  // what tells the issuer a stage may be read
            int stage = 0;
            uint32_t phase = 0, su = 0, tcount = 0;
            // `ready`: the full barrier of the current stage was already seen complete.  It is probed (one non-blocking
            // try_wait) in the middle of the previous stage's MMAs, so its ~100-cycle latency overlaps queued tensor
            // work instead of draining the MMA queue at every stage boundary.
            bool ready = false;
            for (int u = blockIdx.x; u < units; u += gridDim.x, ++tcount) {
                const uint32_t bb = tcount & 1u, par = (tcount >> 1) & 1u;
                const uint32_t tm_brick = tmem_base + bb * (P * NT);
                if (KWF) {
                    // ---- kw-fused: slabs are resident (slab index = chunk * 3 + kw), one stage = (chunk, input plane)
                    if (tcount == 0)
                        for (int ph = 0; ph < nphases; ++ph) mbar_wait(&wfull_bar[ph], 0u);
                    constexpr uint32_t kRow16 = kRowBytes >> 4;
                    const uint32_t a_hi = static_cast<uint32_t>(make_smem_desc(0, 10u * kRowBytes, kLayout) >> 32);
                    for (int c = 0; c < a.nchunks; ++c) {
                        const uint32_t sb16 = slabs16 + static_cast<uint32_t>(c) * 3u * slab16;
                        const bool first_phase = (c == 0), last_phase = (c == a.nchunks - 1);
                        auto issue_plane = [&](const int p, const int nblk, const int kd_lo, const bool fresh) {
                            const uint32_t d_tmem = tm_brick + static_cast<uint32_t>(P - 1 - (p - kd_lo)) * NT;
                            const uint32_t idesc = nblk == 3 ? idesc3 : (nblk == 2 ? idesc2 : idesc1);
                            if (fresh) mbar_wait(&tempty_bar[bb * P + p], par ^ 1u);
                            if (!ready) mbar_wait(&ready_bar[stage], phase);
                            tc_fence_after();
                            const uint32_t sa16 = stages16 + static_cast<uint32_t>(stage) * stage16;
                            const uint32_t sbk16 = sb16 + static_cast<uint32_t>(kd_lo) * kTap16;
                            const int nstage = (stage + 1 == a.nstages) ? 0 : stage + 1;
                            const uint32_t nphase = (stage + 1 == a.nstages) ? (phase ^ 1u) : phase;
#pragma unroll
                            for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                                    for (int k = 0; k < CC / 16; ++k) {
                                        const uint32_t ad = sa16 + (kh * 10 + kw) * kRow16 + 2 * k;
                                        const uint32_t bd = sbk16 + kw * slab16 + kh * 3 * kTap16 + 2 * k;
                                        if (kw == 0 && kh == 0 && k == 0 && fresh) {
                                            umma_bf16_lo2(d_tmem, ad, a_hi, bd, desc_hi, idesc1, 0u);
                                            if (nblk > 1)
                                                umma_bf16_lo2(d_tmem + NT, ad, a_hi, bd + kTap16, desc_hi,
                                                              nblk == 3 ? idesc2 : idesc1, 1u);
                                        } else {
                                            umma_bf16_lo2(d_tmem, ad, a_hi, bd, desc_hi, idesc, 1u);
                                        }
                                    }
                                }
                                if (kw == 0) ready = mbar_try_wait(&ready_bar[nstage], nphase);  // probe the next stage
                            }
                            umma_commit(&empty_bar[stage]);
                            if (last_phase && p >= 2) umma_commit(&tfull_bar[bb * P + (p - 2)]);  // plane p-2 is complete
                            stage = nstage;
                            phase = nphase;
                        };
                        issue_plane(0, 1, 0, first_phase);
                        issue_plane(1, 2, 0, first_phase);
                        for (int p = 2; p < P; ++p) issue_plane(p, 3, 0, first_phase);
                        issue_plane(P, 2, 1, false);
                        issue_plane(P + 1, 1, 2, false);
                    }
                    continue;
                }
                for (int ph = 0; ph < nphases; ++ph, ++su) {
                    uint32_t buf;
                    if (resident) {
                        buf = static_cast<uint32_t>(ph);
                        if (tcount == 0) mbar_wait(&wfull_bar[buf], 0u);
                    } else {
                        buf = su % static_cast<uint32_t>(a.nslabbuf);
                        mbar_wait(&wfull_bar[buf], (su / static_cast<uint32_t>(a.nslabbuf)) & 1u);
                    }
                    const uint32_t sb16 = slabs16 + buf * slab16;
                    const bool first_phase = (ph == 0), last_phase = (ph == nphases - 1);
                    // One stage = input plane p.  It feeds output planes q = p - kd, kd in [kd_lo, kd_lo + nblk): their
                    // accumulators are adjacent TMEM column blocks in ascending kd order starting at plane p - kd_lo.
                    // nblk / kd_lo are literals at every call site, so each call compiles to a straight MMA sequence.
                    auto issue_plane = [&](const int p, const int nblk, const int kd_lo, const bool fresh) {
                        const uint32_t d_tmem = tm_brick + static_cast<uint32_t>(P - 1 - (p - kd_lo)) * NT;
                        const uint32_t idesc = nblk == 3 ? idesc3 : (nblk == 2 ? idesc2 : idesc1);
                        if (fresh) mbar_wait(&tempty_bar[bb * P + p], par ^ 1u);  // epilogue has drained plane p's slot
                        if (!ready) mbar_wait(&ready_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t sa16 = stages16 + static_cast<uint32_t>(stage) * stage16;
                        const uint32_t sbk16 = sb16 + static_cast<uint32_t>(kd_lo) * kTap16;
                        const int nstage = (stage + 1 == a.nstages) ? 0 : stage + 1;
                        const uint32_t nphase = (stage + 1 == a.nstages) ? (phase ^ 1u) : phase;
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                            for (int k = 0; k < CC / 16; ++k) {
                                const uint32_t ad = sa16 + kh * kAtom16 + 2 * k;
                                const uint32_t bd = sbk16 + kh * 3 * kTap16 + 2 * k;
                                if (kh == 0 && k == 0 && fresh) {
                                    // the new plane's block must overwrite, the older planes' blocks accumulate: split
                                    umma_bf16_lo(d_tmem, ad, bd, desc_hi, idesc1, 0u);
                                    if (nblk > 1)
                                        umma_bf16_lo(d_tmem + NT, ad, bd + kTap16, desc_hi, nblk == 3 ? idesc2 : idesc1, 1u);
                                } else {
                                    umma_bf16_lo(d_tmem, ad, bd, desc_hi, idesc, 1u);
                                }
                            }
                            if (kh == 0) ready = mbar_try_wait(&ready_bar[nstage], nphase);  // probe the next stage early
                        }
                        umma_commit(&empty_bar[stage]);  // frees the activation slot once these MMAs have read it
                        if (last_phase && p >= 2) umma_commit(&tfull_bar[bb * P + (p - 2)]);  // plane p-2 is complete
                        stage = nstage;
                        phase = nphase;
                    };
                    issue_plane(0, 1, 0, first_phase);
                    issue_plane(1, 2, 0, first_phase);
                    for (int p = 2; p < P; ++p) issue_plane(p, 3, 0, first_phase);
                    issue_plane(P, 2, 1, false);
                    issue_plane(P + 1, 1, 2, false);
                    if (!resident) umma_commit(&wempty_bar[buf]);
                }
            }
        }
        __syncwarp();
    } else if (XF && warp >= 7) {
        // =========================================================== in-consumer norm transform (warps 7..10)
        const uint32_t tt = threadIdx.x - 7 * 32;  // 0..127: the thread's 16-byte column of every 2 KB
        constexpr uint32_t kColMask = CC / 8 - 1;  // 16-byte columns per row: 8 (SW128) / 4 (SW64)
        const uint32_t col = ((tt >> 0) & kColMask) ^ ((tt >> 3) & kColMask);  // physical column ^ address bits [7..]: logical
        const float4* table = reinterpret_cast<const float4*>(a.in_norm);
        const uint32_t nbytes = a.a_stage_bytes;
        int stage = 0;
        uint32_t phase = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const Unit t = decode_unit(a, u, P);
            const int nloads = KWF ? a.nchunks : nphases;
            int c_loaded = -1;
            float sc[8], sh[8], sl[8];
            for (int ph = 0; ph < nloads; ++ph) {
                const int c = KWF ? ph : ph / a.kwn;
                if (c != c_loaded) {  // (scale, shift, slope) of this thread's 8 channels for (batch item, K chunk)
                    const float4* row = table + (static_cast<size_t>(t.n) * a.in_norm_c + c * CC + col * 8);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float4 q = __ldg(row + e);
                        sc[e] = q.x, sh[e] = q.y, sl[e] = q.z;
                    }
                    c_loaded = c;
                }
                for (int p = 0; p < P + 2; ++p) {
                    mbar_wait(&full_bar[stage], phase);
                    const uint32_t base = smem_u32(stages) + static_cast<uint32_t>(stage) * a.a_stage_bytes;
                    for (uint32_t off = tt * 16; off < nbytes; off += 128 * 16) {
                        uint32_t w[4];
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n"
                                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                                     : "r"(base + off));
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float x0, x1;
                            if (a.in_f16) {
                                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
                                x0 = f.x, x1 = f.y;
                            } else {
                                const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
                                x0 = __bfloat162float(b.x), x1 = __bfloat162float(b.y);
                            }
                            float y0 = fmaf(x0, sc[2 * q], sh[2 * q]), y1 = fmaf(x1, sc[2 * q + 1], sh[2 * q + 1]);
                            y0 = y0 > 0.f ? y0 : y0 * sl[2 * q];
                            y1 = y1 > 0.f ? y1 : y1 * sl[2 * q + 1];
                            y0 = (x0 == x0) ? y0 : 0.f;  // NaN = out-of-bounds fill = conv padding
                            y1 = (x1 == x1) ? y1 : 0.f;
                            if (a.in_f16) {
                                const __half2 h = __floats2half2_rn(y0, y1);
                                w[q] = *reinterpret_cast<const uint32_t*>(&h);
                            } else {
                                const __nv_bfloat162 b = __floats2bfloat162_rn(y0, y1);
                                w[q] = *reinterpret_cast<const uint32_t*>(&b);
                            }
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(base + off), "r"(w[0]), "r"(w[1]),
                                     "r"(w[2]), "r"(w[3])
                                     : "memory");
                    }
                    fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's (async proxy) reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&xfull_bar[stage]);
                    if (++stage == a.nstages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else {
        // =========================================================== epilogue (4 warps per group, one TMEM lane
        // quadrant each: a warp may only touch lanes 32 * (warp % 4) ..)
        const int q4 = warp & 3;
        const int group = (warp >= 7) ? 1 : 0;
        const int row = q4 * 32 + lane;
        const int iw = row & 7, ih = row >> 3;
        EpiParams epi;
        epi.sbias = sbias;
        epi.has_bias = a.bias != nullptr;
        epi.stats = a.stats;
        epi.cout = a.cout;
        epi.No = a.tn;
        epi.act = a.act;
        epi.slope = a.slope;
        epi.out_f16 = a.out_f16;
        epi.stats = STATS ? a.stats : nullptr;
        epi.guard = (a.overflow != nullptr && a.out_f16) ? 1 : 0;
        epi.split_stride = 0;
        epi.stage = nullptr;
        epi.stage_wide = epi.stage_sub = 0;
        EpiGuard guard;
        guard.init();
        StatAcc sacc[kChunks];
        // per-thread sums over ALL planes this warp sees of one batch item (STATS only): fp32 sums of a few hundred values;
        // the warp transpose-reduce (62 shuffles + ~190 selects / adds per chunk) runs once per batch item, not per brick —
        // ncu source view of the 16 -> 64 first layer: the reduce was 62 of the 414 warp instructions an epilogue warp
        // issues per plane and chunk, in an epilogue that is issue-bound
        float t1[kChunks][32], t2[kChunks][32];
        const int cb0 = kColSplit ? group * 32 : 0;  // first column of this warp's chunk(s)
#pragma unroll
        for (int j = 0; j < kChunks; ++j) {
            sacc[j].s1 = sacc[j].s2 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) t1[j][i] = t2[j][i] = 0.f;
        }
        int stat_n = -1;  // batch item the running statistics belong to
        uint32_t tcount = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x, ++tcount) {
            const Unit t = decode_unit(a, u, P);
            const uint32_t bb = tcount & 1u, par = (tcount >> 1) & 1u;
            if (STATS && t.n != stat_n) {
#pragma unroll
                for (int j = 0; j < kChunks; ++j) {
                    if (kThreadAcc && stat_n >= 0) {
                        stats_transpose_reduce(t1[j], t2[j], lane, sacc[j]);
#pragma unroll
                        for (int i = 0; i < 32; ++i) t1[j][i] = t2[j][i] = 0.f;
                    }
                    flush_stats(epi, sacc[j], cb0 + j * 32, lane, stat_n);
                }
                stat_n = t.n;
            }
            __nv_bfloat16* obase = a.out + t.n * a.os_n + static_cast<long long>(t.h0 + ih) * a.os_h +
                                   static_cast<long long>(t.w0 + iw) * a.os_w + a.out_c_off;
            for (int q = kColSplit ? 0 : group; q < P; q += kColSplit ? 1 : kEpiGroups) {
                const uint32_t slot = bb * P + static_cast<uint32_t>(q);
                mbar_wait(&tfull_bar[slot], par);
                tc_fence_after();
                __nv_bfloat16* orow = obase + static_cast<long long>(t.d0 + q) * a.os_d;
                const uint32_t t_addr = tmem_base + (bb * P + static_cast<uint32_t>(P - 1 - q)) * NT + (static_cast<uint32_t>(q4 * 32) << 16);
#pragma unroll
                for (int j = 0; j < kChunks; ++j) {
                    const int cb = cb0 + j * 32;
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + cb, v);
                    tmem_ld_wait();
                    epilogue_32cols<kThreadAcc>(v, epi, cb, true, lane, sacc[j], orow, t1[j], t2[j], guard);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[slot]);
            }
        }
        if (STATS) {
#pragma unroll
            for (int j = 0; j < kChunks; ++j) {
                if (kThreadAcc && stat_n >= 0) stats_transpose_reduce(t1[j], t2[j], lane, sacc[j]);
                flush_stats(epi, sacc[j], cb0 + j * 32, lane, stat_n);
            }
        }
        if (epi.guard) guard.flush(a.overflow);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 6) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int CC, int NT, bool STATS, bool KWF, bool XF>
cudaError_t launch_variant(const BrickArgs& a, int grid, size_t smem_bytes, cudaStream_t stream) {
    static unsigned long long attr_done = 0;  // per device
    if (cudaError_t e = ensure_max_smem(conv_brick_kernel<CC, NT, STATS, KWF, XF>, &attr_done, 232448); e != cudaSuccess)
        return e;
    conv_brick_kernel<CC, NT, STATS, KWF, XF><<<grid, brick_threads(CC, NT, STATS, XF), smem_bytes, stream>>>(a);
    return cudaGetLastError();
}

template <int CC, int NT, bool XF>
cudaError_t launch_stats(const BrickArgs& a, int grid, size_t smem_bytes, cudaStream_t stream) {
    if (a.kwf)
        return a.stats != nullptr ? launch_variant<CC, NT, true, true, XF>(a, grid, smem_bytes, stream)
                                  : launch_variant<CC, NT, false, true, XF>(a, grid, smem_bytes, stream);
    return a.stats != nullptr ? launch_variant<CC, NT, true, false, XF>(a, grid, smem_bytes, stream)
                              : launch_variant<CC, NT, false, false, XF>(a, grid, smem_bytes, stream);
}

}  // namespace
}  // namespace bsg
