// Epilogue shared by the tcgen05 conv kernels: 32 fp32 accumulator columns of one output voxel (one TMEM lane)
// -> bias -> optional norm statistics -> optional LeakyReLU -> 16-bit channels-last store.
// (reference: ConvDropoutNormNonlin.forward, model_architecture/generic_UNet.py:68-72)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace bsg {

struct EpiParams {
    const float* sbias;  // shared memory, [cout_pad]
    float* stats;        // [No][cout][2] running (sum, sum of squares) of the pre-activation output, or null
    int cout;            // valid output channels
    int No;              // batch extent
    int act;             // 1: LeakyReLU(slope)
    float slope;
    int out_f16;         // 1: store IEEE fp16 instead of bf16
};

// Per-lane running norm statistics of one 32-column chunk: after the transpose-reduce lane l owns channel co + l.
// They stay in registers across tiles and go to global memory (one atomicAdd pair per lane) only when the batch item
// or the channel block changes: per-tile atomics on the [N][C][2] table serialise in L2 (measured: +2.5..4.8 ms per
// full-resolution layer).
struct StatAcc {
    float s1, s2;
};

__device__ __forceinline__ void flush_stats(const EpiParams& e, StatAcc& acc, int co, int lane, int n) {
    if (co + lane < e.cout && n >= 0 && n < e.No && (acc.s1 != 0.f || acc.s2 != 0.f)) {
        float* sp = e.stats + (static_cast<long long>(n) * e.cout + co + lane) * 2;
        atomicAdd(sp, acc.s1);
        atomicAdd(sp + 1, acc.s2);
    }
    acc.s1 = 0.f;
    acc.s2 = 0.f;
}

// v: the 32 accumulator columns [co, co+32) of this thread's voxel; orow: the voxel's first output channel;
// valid: voxel inside the tensor; acc: this lane's running statistics for the chunk (used when e.stats != null).
__device__ __forceinline__ void epilogue_32cols(const uint32_t (&v)[32], const EpiParams& e, int co, bool valid, int lane,
                                                StatAcc& acc, __nv_bfloat16* orow) {
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) + e.sbias[co + i];
    if (e.stats != nullptr) {
        // per-channel sum / sum of squares over this warp's 32 voxels: transpose-reduce (31 shuffles each)
        float s1[32], s2[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float x = valid ? f[i] : 0.f;
            s1[i] = x;
            s2[i] = x * x;
        }
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
    if (e.act == 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = f[i] > 0.f ? f[i] : f[i] * e.slope;
    }
    if (!valid) return;
    if (co + 32 <= e.cout) {
        uint4* dst = reinterpret_cast<uint4*>(orow + co);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint4 u;
            if (e.out_f16) {
                __half2 p0 = __floats2half2_rn(f[8 * i + 0], f[8 * i + 1]);
                __half2 p1 = __floats2half2_rn(f[8 * i + 2], f[8 * i + 3]);
                __half2 p2 = __floats2half2_rn(f[8 * i + 4], f[8 * i + 5]);
                __half2 p3 = __floats2half2_rn(f[8 * i + 6], f[8 * i + 7]);
                u.x = *reinterpret_cast<uint32_t*>(&p0);
                u.y = *reinterpret_cast<uint32_t*>(&p1);
                u.z = *reinterpret_cast<uint32_t*>(&p2);
                u.w = *reinterpret_cast<uint32_t*>(&p3);
            } else {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(f[8 * i + 0], f[8 * i + 1]);
                __nv_bfloat162 p1 = __floats2bfloat162_rn(f[8 * i + 2], f[8 * i + 3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(f[8 * i + 4], f[8 * i + 5]);
                __nv_bfloat162 p3 = __floats2bfloat162_rn(f[8 * i + 6], f[8 * i + 7]);
                u.x = *reinterpret_cast<uint32_t*>(&p0);
                u.y = *reinterpret_cast<uint32_t*>(&p1);
                u.z = *reinterpret_cast<uint32_t*>(&p2);
                u.w = *reinterpret_cast<uint32_t*>(&p3);
            }
            dst[i] = u;
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
