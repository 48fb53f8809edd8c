// Epilogue shared by the tcgen05 conv kernels: 32 fp32 accumulator columns of one output voxel (one TMEM lane)
// -> bias -> optional norm statistics -> optional LeakyReLU -> 16-bit channels-last store.
// (reference: ConvDropoutNormNonlin.forward, model_architecture/generic_UNet.py:68-72)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "bsg_ptx.cuh"

namespace bsg {

__device__ __forceinline__ void st_global_v8(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                             uint32_t a5, uint32_t a6, uint32_t a7) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3),
                 "r"(a4), "r"(a5), "r"(a6), "r"(a7)
                 : "memory");
}

struct EpiParams {
    const float* sbias;  // shared memory, [cout_pad]
    int has_bias;        // 0: the layer has no bias (nnU-Net's transposed convs): skip the loads and adds
    double* stats;       // [No][cout][2] running (sum, sum of squares) of the pre-activation output, or null
    int cout;            // valid output channels
    int No;              // batch extent
    int act;             // 1: LeakyReLU(slope)
    float slope;
    int out_f16;         // 1: store IEEE fp16 instead of bf16
    int guard;           // 1: track the largest stored magnitude (fp16 range guard, see EpiGuard)
    uint8_t* stage;      // kEpiStage: the 32 x 32 chunk goes here instead of to global memory; the caller issues the TMA tensor
                         // store (tile kernel, tma_out).  stage_wide = 0: rows of 64 bytes (64B swizzle); 1: rows of 128 bytes
                         // (128B swizzle), this chunk being their half `stage_sub`
    int stage_wide, stage_sub;
    int split_stride;    // SPLIT epilogues: channels between the three blocks [hi | hi | lo] of the fp16x3 output
};

// fp16 range guard: IEEE fp16 saturates at 65504 and nothing downstream would notice an inf (the reference's CUDA path
// has the same exposure under autocast).  Every epilogue thread keeps the largest magnitude it stored; at the end of
// its role a value beyond the fp16 range sets *flag, which the engine reads back once per case and answers by
// re-planning the network in bf16 (engine.py).  16 FMNMX3 per 32 columns.
struct EpiGuard {
    float amax;
    __device__ __forceinline__ void init() { amax = 0.f; }
    __device__ __forceinline__ void flush(int* flag) const {
        if (flag != nullptr && !(amax <= 65504.f)) atomicOr(flag, 1);
    }
};

// Per-lane running norm statistics of one 32-column chunk: after the transpose-reduce lane l owns channel co + l.
// They stay in registers across tiles and go to global memory (one atomicAdd pair per lane) only when the batch item
// or the channel block changes: per-tile atomics on the [N][C][2] table serialise in L2 (measured: +2.5..4.8 ms per
// full-resolution layer).  The table is fp64: every flushed partial sum is a deterministic fp32 value (a CTA's share
// of the work and its summation order are static), and fp64 additions of a few hundred such values are exact to
// ~1e-16 whatever order the atomics land in — so the fp32 (scale, shift) derived from them, and with them the whole
// forward, are reproducible from run to run (fp32 atomics moved label decisions on ~3e-4 of the voxels of a GroupNorm
// net between runs: a 1e-7 change of a mean flips fp16 roundings downstream).
struct StatAcc {
    float s1, s2;
};

__device__ __forceinline__ void flush_stats(const EpiParams& e, StatAcc& acc, int co, int lane, int n) {
    if (co + lane < e.cout && n >= 0 && n < e.No && (acc.s1 != 0.f || acc.s2 != 0.f)) {
        double* sp = e.stats + (static_cast<long long>(n) * e.cout + co + lane) * 2;
        atomicAdd(sp, static_cast<double>(acc.s1));
        atomicAdd(sp + 1, static_cast<double>(acc.s2));
    }
    acc.s1 = 0.f;
    acc.s2 = 0.f;
}

// Packed fp32 pairs (sm_100 FADD2 / FFMA2: two fp32 operations per issue slot) for the epilogue's bias add and
// per-thread statistic sums — the first layers' epilogues are bound by instruction issue, not by the tensor pipe.
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1) {  // d = a + b
    asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 ra, ra, rb;\n\t"
        "mov.b64 {%0, %1}, ra;\n\t}\n"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fma2(float& c0, float& c1, float a0, float a1, float b0, float b1) {  // c += a * b
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}\n"
        : "+f"(c0), "+f"(c1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// Per-channel sum / sum of squares over the warp's 32 lanes by transpose-reduce (31 shuffles per array): afterwards
// lane l owns channel l of the chunk, added into its running statistics.
__device__ __forceinline__ void stats_transpose_reduce(float (&s1)[32], float (&s2)[32], int lane, StatAcc& acc) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send1 = up ? s1[i] : s1[i + off];
            const float keep1 = up ? s1[i + off] : s1[i];
            s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, off);
            const float send2 = up ? s2[i] : s2[i + off];
            const float keep2 = up ? s2[i + off] : s2[i];
            s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
        }
    }
    acc.s1 += s1[0];
    acc.s2 += s2[0];
}

// Norm statistics of one 32-column chunk when the rows of a warp span SEVERAL batch items (tile boxes with fewer than
// 32 voxels per item: the <= 2^3 levels of a deep net run with batch > 4 per tile).  One masked transpose-reduce and
// one flush per batch item of the warp; cold path (tiny layers only), kept out of line.
// It re-reads the chunk from TMEM itself (taddr): handing it the caller's register array by reference would force that
// array into local memory for the whole (hot) epilogue.
static __device__ __noinline__ void stats_chunk_grouped(uint32_t taddr, const float* sbias, bool has_bias, double* stats,
                                                        int cout, int No, int co, bool valid, int lane, int vox_per_item,
                                                        int n_first) {
    EpiParams e;
    e.stats = stats;
    e.cout = cout;
    e.No = No;
    uint32_t v[32];
    tmem_ld_32x32(taddr, v);
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) + (has_bias ? sbias[co + i] : 0.f);
    for (int g = 0; g * vox_per_item < 32; ++g) {
        const bool mine = valid && (lane / vox_per_item == g);
        float s1[32], s2[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float x = mine ? f[i] : 0.f;
            s1[i] = x;
            s2[i] = x * x;
        }
        StatAcc acc;
        acc.s1 = acc.s2 = 0.f;
        stats_transpose_reduce(s1, s2, lane, acc);
        flush_stats(e, acc, co, lane, n_first + g);
    }
}

// v: the 32 accumulator columns [co, co+32) of this thread's voxel; orow: the voxel's first output channel;
// valid: voxel inside the tensor.  Norm statistics (when e.stats != null): THREAD_ACC = false reduces this tile's
// values over the warp right away into `acc`; THREAD_ACC = true only adds them to the caller's per-thread sums
// t1 / t2 (32 + 32 registers per chunk), which the caller reduces once per brick with stats_transpose_reduce —
// 64 FMAs per tile and chunk instead of 62 shuffles + ~190 selects/adds.
//
// SPLIT (fp32-equivalent mode, engine dtype "fp32"): the fp32 result y is stored as TWO fp16 numbers, hi = fp16(y) and
// lo = fp16(y - hi) (22 significant bits between them), laid out as three channel blocks [hi | hi | lo] `split_stride`
// channels apart — exactly the K layout the next conv contracts against [w_hi | w_lo | w_hi]:
// y*w ~= hi*w_hi + hi*w_lo + lo*w_hi, three fp16 MMAs with fp32 accumulation per fp32 multiply-add.
// EPI: kEpiDirect (per-thread rows straight to global memory), kEpiSplit (above), kEpiStage (the 32 x 32 chunk goes to the
// shared-memory staging row block e.stage, the caller issues a TMA tensor store).
constexpr int kEpiDirect = 0, kEpiSplit = 1, kEpiStage = 2;
template <bool THREAD_ACC, int EPI = kEpiDirect>
__device__ __forceinline__ void epilogue_32cols(const uint32_t (&v)[32], const EpiParams& e, int co, bool valid, int lane,
                                                StatAcc& acc, __nv_bfloat16* orow, float (&t1)[32], float (&t2)[32],
                                                EpiGuard& guard) {
    float f[32];
    if (e.has_bias) {
        // bias: 8 x ld.shared.v4 (warp-wide broadcast); `e.sbias` is a generic pointer, which would compile to 32
        // generic loads per chunk
        const uint32_t baddr = static_cast<uint32_t>(__cvta_generic_to_shared(e.sbias + co));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float b0, b1, b2, b3;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                         : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3)
                         : "r"(baddr + 16u * i));
            add2(f[4 * i + 0], f[4 * i + 1], __uint_as_float(v[4 * i + 0]), __uint_as_float(v[4 * i + 1]), b0, b1);
            add2(f[4 * i + 2], f[4 * i + 3], __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]), b2, b3);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
    }
    if (e.stats != nullptr) {
        if (THREAD_ACC) {
            if (valid) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    add2(t1[2 * i], t1[2 * i + 1], t1[2 * i], t1[2 * i + 1], f[2 * i], f[2 * i + 1]);
                    fma2(t2[2 * i], t2[2 * i + 1], f[2 * i], f[2 * i + 1], f[2 * i], f[2 * i + 1]);
                }
            }
        } else {
            float s1[32], s2[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float x = valid ? f[i] : 0.f;
                s1[i] = x;
                s2[i] = x * x;
            }
            stats_transpose_reduce(s1, s2, lane, acc);
        }
    }
    if (e.act == 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = f[i] > 0.f ? f[i] : f[i] * e.slope;
    }
    if (!valid && EPI != kEpiStage) return;
    if (e.guard && valid) {
        float m = guard.amax;
#pragma unroll
        for (int i = 0; i < 16; ++i) m = fmaxf(m, fmaxf(fabsf(f[2 * i]), fabsf(f[2 * i + 1])));
        guard.amax = m;
    }
    if constexpr (EPI == kEpiSplit) {
        __half* base = reinterpret_cast<__half*>(orow);
        if (co + 32 <= e.cout) {
            uint32_t ph[16], pl[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
                const float2 hf = __half22float2(h);
                const __half2 l = __floats2half2_rn(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
                ph[i] = *reinterpret_cast<const uint32_t*>(&h);
                pl[i] = *reinterpret_cast<const uint32_t*>(&l);
            }
#pragma unroll
            for (int blk = 0; blk < 3; ++blk) {
                uint4* d4 = reinterpret_cast<uint4*>(base + co + blk * e.split_stride);
                const uint32_t* p = blk == 2 ? pl : ph;
#pragma unroll
                for (int i = 0; i < 4; ++i) d4[i] = make_uint4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (co + i < e.cout) {
                    const __half h = __float2half_rn(f[i]);
                    const __half l = __float2half_rn(f[i] - __half2float(h));
                    base[co + i] = h;
                    base[co + i + e.split_stride] = h;
                    base[co + i + 2 * e.split_stride] = l;
                }
        }
        return;
    }
    if (co + 32 <= e.cout || EPI == kEpiStage) {
        // one uniform branch around the whole block: a per-element `out_f16 ? half : bf16` is if-converted into BOTH
        // F2FP conversions plus a select, and the conversion pipe is what bounds the store-heavy epilogues (ncu on the
        // transposed conv: 82 % busy)
        uint32_t pk[16];
        if (e.out_f16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
                pk[i] = *reinterpret_cast<uint32_t*>(&h);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                pk[i] = *reinterpret_cast<uint32_t*>(&b);
            }
        }
        if constexpr (EPI == kEpiStage) {
            // staging row of this thread: 64 bytes at lane * 64 (or its half of the 128 bytes at lane * 128), the 16-byte
            // pieces XOR-swizzled the way the store's tensor map undoes (CU_TENSOR_MAP_SWIZZLE_64B: address bits [4, 6)
            // ^ [7, 9); _128B: bits [4, 7) ^ [7, 10)) — conflict-free 128-bit shared-memory stores.  Rows outside the
            // tensor and channels >= cout are clipped by the tensor store.
            const uint32_t ulane = static_cast<uint32_t>(lane);
            const uint32_t row = smem_u32(e.stage) + (e.stage_wide ? ulane * 128u : ulane * 64u);
            const uint32_t x = e.stage_wide ? (ulane & 7u) : ((ulane >> 1) & 3u);
            const uint32_t p0 = e.stage_wide ? static_cast<uint32_t>(e.stage_sub) * 4u : 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(row + (((p0 + static_cast<uint32_t>(i)) ^ x) << 4)),
                             "r"(pk[4 * i]), "r"(pk[4 * i + 1]), "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                             : "memory");
            return;
        }
        __nv_bfloat16* dst = orow + co;
        if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
            // two 256-bit stores (STG.256, sm_100): whole 32-byte sectors, half the L2 write requests of 4 x 128-bit
            st_global_v8(dst, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
            st_global_v8(dst + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
        } else {
            uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
            for (int i = 0; i < 4; ++i) d4[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (co + i < e.cout) {
                if (e.out_f16)
                    reinterpret_cast<__half*>(orow)[co + i] = __float2half_rn(f[i]);
                else
                    orow[co + i] = __float2bfloat16_rn(f[i]);
            }
    }
}

}  // namespace bsg
