// Fused HBM-bound passes around the conv stacks of predict_3D (nnU-Net v1 `_internal_predict_3D_3Dconv_tiled` and
// `_internal_maybe_mirror_and_pred_3D`; call site
// run_brats2021_inference_singlethread.py:97-106):
//   gather_patch_tta      tile crop + 8-way mirror flips + fp32 -> bf16 channels-last (input side of the TTA)
//   norm_finalize / norm_apply_lrelu   InstanceNorm / GroupNorm from the conv epilogue's (sum, sumsq) + LeakyReLU
//   head_tta_accumulate   1x1x1 seg head + sigmoid/softmax + un-flip + mean over mirrors + Gaussian weight +
//                         accumulate into the full-volume fp32 accumulator
//   finalize              / weight-sum, mean over folds, ordered-threshold or argmax decision -> uint8 labels
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "bsg_common.cuh"

namespace bsg {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxMirrors = 8;
constexpr int kMaxClasses = 8;

struct MirrorSet {
    int n;
    int code[kMaxMirrors];  // bit0: flip x (tensor dim 4), bit1: flip y (dim 3), bit2: flip z (dim 2)
};

// 16 packed channels (8 pairs) of a kw-packed voxel from its three w-neighbours' C channels: forward order
// [x(w-1) | x(w) | x(w+1) | 0] and the order a copy flipped along x sees, [x(w+1) | x(w) | x(w-1) | 0].
template <int C>
__device__ __forceinline__ void kwpack_pairs(const float (&v)[3][5], int out_f16, uint32_t (&pkf)[8], uint32_t (&pkr)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float f[2], r[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int e = 2 * j + q;  // packed channel e: neighbour k = e / C, channel c = e % C (compile-time)
            const int k = e / C, c = e % C;
            f[q] = e < 3 * C ? v[k][c] : 0.f;
            r[q] = e < 3 * C ? v[2 - k][c] : 0.f;
        }
        if (out_f16) {
            __half2 p = __floats2half2_rn(f[0], f[1]), q2 = __floats2half2_rn(r[0], r[1]);
            pkf[j] = *reinterpret_cast<uint32_t*>(&p);
            pkr[j] = *reinterpret_cast<uint32_t*>(&q2);
        } else {
            __nv_bfloat162 p = __floats2bfloat162_rn(f[0], f[1]), q2 = __floats2bfloat162_rn(r[0], r[1]);
            pkf[j] = *reinterpret_cast<uint32_t*>(&p);
            pkr[j] = *reinterpret_cast<uint32_t*>(&q2);
        }
    }
}

// out[m][d][h][w][c] = vol[c][z0 + fz(d)][y0 + fy(h)][x0 + fx(w)], c < C; channels C..cpad-1 are zero.
// grid (ceil(P1*P2 / 256), P0): one thread per SOURCE voxel of the tile — it is read and converted once and stored to
// its position in each of the (up to 8) mirrored copies; a warp's 32 consecutive w land on 32 consecutive (or
// reversed) voxels of every copy, so each store instruction still covers one contiguous 1 KB run.  32-bit index
// math, 256-bit stores.
__global__ void __launch_bounds__(kThreads) gather_patch_kernel(const float* __restrict__ vol, int C, int Z, int Y,
                                                                int X, int z0, int y0, int x0, int P0, int P1, int P2,
                                                                const MirrorSet ms, __nv_bfloat16* __restrict__ out,
                                                                int cpad, int out_f16, int kwpack) {
    const int hw = blockIdx.x * blockDim.x + threadIdx.x;
    if (hw >= P1 * P2) return;
    const int d = blockIdx.y;
    const int h = hw / P2, w = hw - h * P2;
    const size_t plane = static_cast<size_t>(Z) * Y * X;
    const float* src = vol + (static_cast<size_t>(z0 + d) * Y + (y0 + h)) * X + (x0 + w);
    if (kwpack == 2) {
        // fp16x3 split of the fp32 input (engine dtype "fp32"): channels [hi (C) | hi (C) | lo (C) | 0...]
        for (int m = 0; m < ms.n; ++m) {
            const int code = ms.code[m];
            const int ow = (code & 1) ? P2 - 1 - w : w;
            const int oh = (code & 2) ? P1 - 1 - h : h;
            const int od = (code & 4) ? P0 - 1 - d : d;
            __half* dst = reinterpret_cast<__half*>(out) + (((static_cast<size_t>(m) * P0 + od) * P1 + oh) * P2 + ow) * cpad;
            for (int c = 0; c < cpad; ++c) dst[c] = __float2half_rn(0.f);
            for (int c = 0; c < C; ++c) {
                const float v = __ldg(src + c * plane);
                const __half hi = __float2half_rn(v);
                dst[c] = hi;
                dst[C + c] = hi;
                dst[2 * C + c] = __float2half_rn(v - __half2float(hi));
            }
        }
        return;
    }
    if (kwpack) {
        // kw-packed layout for the network's first conv (3*C <= 16): channel k*C + c of a voxel holds channel c of its
        // w-neighbour k-1 IN THE COPY's orientation (zero outside the tile = the conv's zero padding), so that the
        // 3x3x3 conv over C channels becomes a 3x3x1 conv over 3*C: 9 taps of K = 16 instead of 27 (the 4 BraTS
        // modalities are padded to K = 16 either way).  A copy flipped along x sees the neighbours swapped.
        float v[3][5];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int ww = w + k - 1;
            const bool inb = ww >= 0 && ww < P2;
#pragma unroll
            for (int c = 0; c < 5; ++c) v[k][c] = (inb && c < C) ? __ldg(src + (k - 1) + c * plane) : 0.f;
        }
        uint32_t pkf[8], pkr[8];
        switch (C) {
            case 1: kwpack_pairs<1>(v, out_f16, pkf, pkr); break;
            case 2: kwpack_pairs<2>(v, out_f16, pkf, pkr); break;
            case 3: kwpack_pairs<3>(v, out_f16, pkf, pkr); break;
            case 4: kwpack_pairs<4>(v, out_f16, pkf, pkr); break;
            default: kwpack_pairs<5>(v, out_f16, pkf, pkr); break;
        }
        for (int m = 0; m < ms.n; ++m) {
            const int code = ms.code[m];
            const int ow = (code & 1) ? P2 - 1 - w : w;
            const int oh = (code & 2) ? P1 - 1 - h : h;
            const int od = (code & 4) ? P0 - 1 - d : d;
            __nv_bfloat16* dst = out + (((static_cast<size_t>(m) * P0 + od) * P1 + oh) * P2 + ow) * cpad;
            const uint32_t* pk = (code & 1) ? pkr : pkf;
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(dst), "r"(pk[0]), "r"(pk[1]),
                         "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                         : "memory");
        }
        return;
    }
    for (int c0 = 0; c0 < cpad; c0 += 16) {
        uint32_t pk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int ca = c0 + 2 * k, cb = ca + 1;
            const float fa = ca < C ? __ldg(src + ca * plane) : 0.f;
            const float fb = cb < C ? __ldg(src + cb * plane) : 0.f;
            if (out_f16) {
                __half2 p = __floats2half2_rn(fa, fb);
                pk[k] = *reinterpret_cast<uint32_t*>(&p);
            } else {
                __nv_bfloat162 p = __floats2bfloat162_rn(fa, fb);
                pk[k] = *reinterpret_cast<uint32_t*>(&p);
            }
        }
        for (int m = 0; m < ms.n; ++m) {
            // the source voxel (d, h, w) is element (fz(d), fy(h), fx(w)) of the copy flipped by code m
            const int code = ms.code[m];
            const int ow = (code & 1) ? P2 - 1 - w : w;
            const int oh = (code & 2) ? P1 - 1 - h : h;
            const int od = (code & 4) ? P0 - 1 - d : d;
            __nv_bfloat16* dst = out + (((static_cast<size_t>(m) * P0 + od) * P1 + oh) * P2 + ow) * cpad + c0;
            if (c0 + 16 <= cpad && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(dst), "r"(pk[0]), "r"(pk[1]),
                             "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                             : "memory");
            } else {
                *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                if (c0 + 8 < cpad) *reinterpret_cast<uint4*>(dst + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- norms
// stats [N][C][2] (sum, sumsq over `count` voxels) -> per (n,c) affine: y = x*scale + shift.
// groups == 0: InstanceNorm (per channel); groups > 0: GroupNorm (C/groups channels share statistics).
// out_stride == 2: scale_shift [N][C][2]; out_stride == 4: rows [coff, coff+C) of a consumer table [N][ctot][4] =
// (scale, shift, slope, 0) (bsg_norm_finalize_table).
__global__ void norm_finalize_kernel(const double* __restrict__ stats, int N, int C, int groups, double count, float eps,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float* __restrict__ scale_shift, int out_stride, int ctot, int coff, float slope) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * C) return;
    const int n = i / C, c = i % C;
    double s1, s2, cnt;
    if (groups > 0) {
        const int cg = C / groups, g = c / cg;
        s1 = s2 = 0.0;
        for (int k = 0; k < cg; ++k) {
            s1 += stats[(static_cast<size_t>(n) * C + g * cg + k) * 2];
            s2 += stats[(static_cast<size_t>(n) * C + g * cg + k) * 2 + 1];
        }
        cnt = count * cg;
    } else {
        s1 = stats[static_cast<size_t>(i) * 2];
        s2 = stats[static_cast<size_t>(i) * 2 + 1];
        cnt = count;
    }
    const double mean = s1 / cnt;
    double var = s2 / cnt - mean * mean;  // biased variance, as torch's instance_norm / group_norm
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + static_cast<double>(eps));
    const double ga = gamma ? static_cast<double>(gamma[c]) : 1.0, be = beta ? static_cast<double>(beta[c]) : 0.0;
    float* o = scale_shift + (static_cast<size_t>(n) * ctot + coff + c) * out_stride;
    o[0] = static_cast<float>(ga * rstd);
    o[1] = static_cast<float>(be - mean * ga * rstd);
    if (out_stride == 4) {
        o[2] = slope;
        o[3] = 0.f;
    }
}

// in place on a channel slice [coff, coff+C) of a (N, V, ctot) 16-bit buffer: x <- lrelu(x*scale + shift).
// One 16-byte group (8 channels) per thread and step.  The launch makes the thread stride a multiple of the groups
// per voxel, so a thread keeps the SAME 8 channels of the SAME batch item (blockIdx.y) for its whole run: its 16
// scale/shift values live in registers and the loop body is one load, 8 FMAs, one store — no index division.
template <bool IN_F16, bool OUT_F16>
__global__ void __launch_bounds__(kThreads) norm_apply_kernel(__nv_bfloat16* __restrict__ x, size_t V, int C, int ctot,
                                                              int coff, const float* __restrict__ scale_shift,
                                                              float slope) {
    const int c8n = C / 8;
    const int n = blockIdx.y;
    const size_t j0 = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;  // multiple of c8n (host guarantees it)
    const int c8 = static_cast<int>(j0 % c8n);
    size_t v = j0 / c8n;
    const size_t vstep = stride / c8n;
    float sc[8], sh[8];
    {
        const float2* ss = reinterpret_cast<const float2*>(scale_shift + (static_cast<size_t>(n) * C + c8 * 8) * 2);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2 t = __ldg(ss + k);
            sc[k] = t.x;
            sh[k] = t.y;
        }
    }
    uint4* p = reinterpret_cast<uint4*>(x + (static_cast<size_t>(n) * V + v) * ctot + coff + c8 * 8);
    const size_t pstep = vstep * ctot / 8;  // uint4 units
    for (; v < V; v += vstep, p += pstep) {
        uint4 u = *p;
        uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float a, b;
            if (IN_F16) {
                const float2 t = __half22float2(*reinterpret_cast<__half2*>(&w[k]));
                a = t.x;
                b = t.y;
            } else {
                const __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&w[k]);
                a = __bfloat162float(t.x);
                b = __bfloat162float(t.y);
            }
            a = fmaf(a, sc[2 * k], sh[2 * k]);
            b = fmaf(b, sc[2 * k + 1], sh[2 * k + 1]);
            a = a > 0.f ? a : a * slope;
            b = b > 0.f ? b : b * slope;
            if (OUT_F16) {
                __half2 o = __floats2half2_rn(a, b);
                w[k] = *reinterpret_cast<uint32_t*>(&o);
            } else {
                __nv_bfloat162 o = __floats2bfloat162_rn(a, b);
                w[k] = *reinterpret_cast<uint32_t*>(&o);
            }
        }
        *p = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// fp16x3 split layout (engine dtype "fp32"): x = hi + lo from blocks 0 and 2 of [hi | hi | lo], normalise + LeakyReLU in
// fp32, write the three blocks back.  One 8-channel group per thread and step, same thread-constant channel mapping.
__global__ void __launch_bounds__(kThreads) norm_apply_split_kernel(__half* __restrict__ x, size_t V, int C, int ctot,
                                                                    int coff, const float* __restrict__ scale_shift,
                                                                    float slope) {
    const int c8n = C / 8;
    const int n = blockIdx.y;
    const size_t j0 = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;  // multiple of c8n (host guarantees it)
    const int c8 = static_cast<int>(j0 % c8n);
    const size_t vstep = stride / c8n;
    float sc[8], sh[8];
    {
        const float2* ss = reinterpret_cast<const float2*>(scale_shift + (static_cast<size_t>(n) * C + c8 * 8) * 2);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2 t = __ldg(ss + k);
            sc[k] = t.x;
            sh[k] = t.y;
        }
    }
    for (size_t v = j0 / c8n; v < V; v += vstep) {
        __half* p = x + (static_cast<size_t>(n) * V + v) * ctot + coff + c8 * 8;
        const uint4 uh = *reinterpret_cast<const uint4*>(p), ul = *reinterpret_cast<const uint4*>(p + 2 * C);
        const uint32_t wh[4] = {uh.x, uh.y, uh.z, uh.w}, wl[4] = {ul.x, ul.y, ul.z, ul.w};
        uint32_t oh[4], ol[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&wh[k]));
            const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&wl[k]));
            float a = fmaf(fh.x + fl.x, sc[2 * k], sh[2 * k]), b = fmaf(fh.y + fl.y, sc[2 * k + 1], sh[2 * k + 1]);
            a = a > 0.f ? a : a * slope;
            b = b > 0.f ? b : b * slope;
            const __half2 h = __floats2half2_rn(a, b);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
            oh[k] = *reinterpret_cast<const uint32_t*>(&h);
            ol[k] = *reinterpret_cast<const uint32_t*>(&l);
        }
        const uint4 vh = make_uint4(oh[0], oh[1], oh[2], oh[3]);
        *reinterpret_cast<uint4*>(p) = vh;
        *reinterpret_cast<uint4*>(p + C) = vh;
        *reinterpret_cast<uint4*>(p + 2 * C) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    }
}

// ---------------------------------------------------------------------------------------------- head + TTA + accumulate
constexpr int kMaxHeadCh = 64;
struct HeadParams {
    float w[kMaxClasses][kMaxHeadCh];
    float b[kMaxClasses];
    int ncls;
    int cfeat;   // feature channels feeding the head (multiple of 8, <= 64)
    int nonlin;  // 0: sigmoid, 1: softmax over classes, 2: identity
    float mirror_weight;  // 1 / num_results of the full TTA (the mirrors of a tile may be split over launches)
};

// grid (ceil(P1*P2 / 256), P0): one thread per tile voxel.  NCLS / CFEAT are compile-time bounds (the valid counts are
// hp.ncls <= NCLS, hp.cfeat <= CFEAT, padded weights are zero), so the head weights become constant-bank operands of
// fully unrolled FMAs.
template <int NCLS, int CFEAT>
__global__ void __launch_bounds__(kThreads) head_tta_accumulate_kernel(
    const __nv_bfloat16* __restrict__ feat, int ctot, int P0, int P1, int P2, const MirrorSet ms,
    const __grid_constant__ HeadParams hp, const int feat_f16, const float* __restrict__ gauss, float* __restrict__ acc, int Z, int Y, int X,
    int z0, int y0, int x0, const float* __restrict__ norm_ss, const float norm_slope, const int lo_off) {
    const int hw = blockIdx.x * blockDim.x + threadIdx.x;
    if (hw >= P1 * P2) return;
    const int d = blockIdx.y;
    const int h = hw / P2, w = hw - h * P2;
    const size_t plane = static_cast<size_t>(Z) * Y * X;
    const float inv = hp.mirror_weight;
    float res[NCLS];
#pragma unroll
    for (int k = 0; k < NCLS; ++k) res[k] = 0.f;
#pragma unroll 2
    for (int m = 0; m < ms.n; ++m) {
        const int code = ms.code[m];
        // prediction m was computed on the flipped tile: its voxel for (d,h,w) sits at the flipped position
        const int sw = (code & 1) ? P2 - 1 - w : w;
        const int sh = (code & 2) ? P1 - 1 - h : h;
        const int sd = (code & 4) ? P0 - 1 - d : d;
        const size_t v = ((static_cast<size_t>(m) * P0 + sd) * P1 + sh) * P2 + sw;
        const uint4* fp = reinterpret_cast<const uint4*>(feat + v * ctot);
        uint4 u[CFEAT / 8];
#pragma unroll
        for (int q = 0; q < CFEAT / 8; ++q) u[q] = (q * 8 < hp.cfeat) ? __ldg(fp + q) : make_uint4(0, 0, 0, 0);
        float logit[NCLS];
#pragma unroll
        for (int k = 0; k < NCLS; ++k) logit[k] = hp.b[k];
#pragma unroll
        for (int q = 0; q < CFEAT / 8; ++q) {
            const uint32_t ww[4] = {u[q].x, u[q].y, u[q].z, u[q].w};
            float f[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (feat_f16) {
                    const float2 h2 = __half22float2(*reinterpret_cast<const __half2*>(&ww[k]));
                    f[2 * k] = h2.x;
                    f[2 * k + 1] = h2.y;
                } else {
                    const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&ww[k]);
                    f[2 * k] = __bfloat162float(b2.x);
                    f[2 * k + 1] = __bfloat162float(b2.y);
                }
            }
            if (lo_off != 0) {  // fp16x3 split features: f = hi + lo (blocks 0 and 2 of [hi | hi | lo])
                const uint4 ul = (q * 8 < hp.cfeat) ? __ldg(reinterpret_cast<const uint4*>(feat + v * ctot + lo_off) + q)
                                                    : make_uint4(0, 0, 0, 0);
                const uint32_t wl[4] = {ul.x, ul.y, ul.z, ul.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 l2 = __half22float2(*reinterpret_cast<const __half2*>(&wl[k]));
                    f[2 * k] += l2.x;
                    f[2 * k + 1] += l2.y;
                }
            }
            if (norm_ss != nullptr) {
                // deferred InstanceNorm / GroupNorm + LeakyReLU of the last conv block (the features are its raw
                // output): y = lrelu(x * scale[m][c] + shift[m][c]) — saves the block's separate normalise pass
                const float4* ss = reinterpret_cast<const float4*>(norm_ss + (static_cast<size_t>(m) * hp.cfeat + q * 8) * 2);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 t = (q * 8 < hp.cfeat) ? __ldg(ss + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float a = fmaf(f[2 * c], t.x, t.y), b = fmaf(f[2 * c + 1], t.z, t.w);
                    f[2 * c] = a > 0.f ? a : a * norm_slope;
                    f[2 * c + 1] = b > 0.f ? b : b * norm_slope;
                }
            }
#pragma unroll
            for (int k = 0; k < NCLS; ++k) {
#pragma unroll
                for (int c = 0; c < 8; ++c) logit[k] = fmaf(hp.w[k][q * 8 + c], f[c], logit[k]);
            }
        }
        if (hp.nonlin == 0) {
#pragma unroll
            for (int k = 0; k < NCLS; ++k) res[k] += inv * (1.0f / (1.0f + expf(-logit[k])));
        } else if (hp.nonlin == 1) {
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < NCLS; ++k)
                if (k < hp.ncls) mx = fmaxf(mx, logit[k]);
            float e[NCLS], sum = 0.f;
#pragma unroll
            for (int k = 0; k < NCLS; ++k) {
                e[k] = (k < hp.ncls) ? expf(logit[k] - mx) : 0.f;
                sum += e[k];
            }
#pragma unroll
            for (int k = 0; k < NCLS; ++k) res[k] += inv * (e[k] / sum);
        } else {
#pragma unroll
            for (int k = 0; k < NCLS; ++k) res[k] += inv * logit[k];
        }
    }
    const float g = gauss ? __ldg(gauss + static_cast<size_t>(d) * P1 * P2 + hw) : 1.0f;
    float* ap = acc + (static_cast<size_t>(z0 + d) * Y + (y0 + h)) * X + (x0 + w);
#pragma unroll
    for (int k = 0; k < NCLS; ++k)
        if (k < hp.ncls) ap[k * plane] += res[k] * g;
}

// ---------------------------------------------------------------------------------------------- finalize
struct FinalizeParams {
    const float* acc[16];  // K accumulators (folds / models to average), each [ncls][nvox]
    int K;
    int ncls;
    int mode;                // 0: argmax, 1: ordered threshold > 0.5 (regions)
    int order[kMaxClasses];  // regions_class_order
};

// fl32(a / w) > 0.5 for w > 0, without the division: round-to-nearest-even maps the quotient to a float above 0.5 iff it
// lies above the midpoint of 0.5 and its successor, 0.5 * (1 + 2^-24) (the tie goes to 0.5, the even one); a and w have
// 24-bit significands, so w * (0.5 + 2^-25) is exact in fp64 and the comparison is exact.
__device__ __forceinline__ bool exceeds_half(float a, float w) {
    return static_cast<double>(a) > static_cast<double>(w) * (0.5 + 0x1p-25);
}

// One voxel's decision from its class probabilities (argmax, or the ordered > 0.5 assignment of the region trainers).
template <int MAXC>
__device__ __forceinline__ int decide_label(const FinalizeParams& fp, const float (&p)[MAXC]) {
    int lab = 0;
    if (fp.mode == 0) {
        float best = p[0];
#pragma unroll
        for (int k = 1; k < MAXC; ++k)
            if (k < fp.ncls && p[k] > best) {
                best = p[k];
                lab = k;
            }
    } else {
#pragma unroll
        for (int k = 0; k < MAXC; ++k)
            if (k < fp.ncls && p[k] > 0.5f) lab = fp.order[k];
    }
    return lab;
}

// VEC voxels per thread and step (VEC = 4: 128-bit loads of the accumulators / weight sums, 32-bit label stores; the
// host picks it when nvox and every pointer allow).  class_probabilities = aggregated_results /
// aggregated_nb_of_predictions (IEEE division, as numpy), then np.mean over the folds.
// MAXC: compile-time bound of the class count (4 for the BraTS region / label sets, else 8) — with arrays of 8 classes the
// kernel needed 96 registers, two blocks per SM, and its 32 KB of loads in flight per SM held it at 37 % of HBM bandwidth.
// K1 (one accumulator, the common case): all class loads of a voxel group are issued before the first division — inside
// the runtime fold loop the compiler keeps each load next to its use and a thread walks its four streams one memory
// round trip after the other (measured 22 % of HBM bandwidth that way).
template <int VEC, bool K1, int MAXC>
__global__ void __launch_bounds__(kThreads, MAXC == 4 ? 4 : 2) finalize_kernel(const FinalizeParams fp, const float* __restrict__ wsum,
                                                            size_t nvox, float* __restrict__ probs,
                                                            uint8_t* __restrict__ seg) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    const size_t ngroups = nvox / VEC;
    for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
        const size_t i = g * VEC;
        float wv[VEC], p[VEC][MAXC];
        if constexpr (VEC == 4) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(wsum + i));
            wv[0] = t.x, wv[1] = t.y, wv[2] = t.z, wv[3] = t.w;
        } else {
            wv[0] = __ldcs(wsum + i);
        }
        if constexpr (K1) {
            float a[MAXC][VEC];
            // regions decision without probabilities: fl(a / w) > 0.5 is decided EXACTLY by one fp64 product and compare
            // (see exceeds_half) — the 12 IEEE divisions per thread made this pass ALU-bound at 18-22 % of HBM bandwidth
            const bool by_compare = fp.mode == 1 && probs == nullptr;
#pragma unroll
            for (int k = 0; k < MAXC; ++k) {
                if (k < fp.ncls) {
                    if constexpr (VEC == 4) {
                        const float4 t = __ldcs(reinterpret_cast<const float4*>(fp.acc[0] + k * nvox + i));
                        a[k][0] = t.x, a[k][1] = t.y, a[k][2] = t.z, a[k][3] = t.w;
                    } else {
                        a[k][0] = __ldcs(fp.acc[0] + k * nvox + i);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < MAXC; ++k) {
                if (k < fp.ncls) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
                        p[v][k] = by_compare ? (exceeds_half(a[k][v], wv[v]) ? 1.f : 0.f) : __fdiv_rn(a[k][v], wv[v]);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < MAXC; ++k) {
                if (k < fp.ncls) {
                    float s[VEC];
                    for (int j = 0; j < fp.K; ++j) {
                        float a[VEC];
                        if constexpr (VEC == 4) {
                            const float4 t = __ldcs(reinterpret_cast<const float4*>(fp.acc[j] + k * nvox + i));
                            a[0] = t.x, a[1] = t.y, a[2] = t.z, a[3] = t.w;
                        } else {
                            a[0] = __ldcs(fp.acc[j] + k * nvox + i);
                        }
#pragma unroll
                        for (int v = 0; v < VEC; ++v) s[v] = j == 0 ? __fdiv_rn(a[v], wv[v]) : s[v] + __fdiv_rn(a[v], wv[v]);
                    }
#pragma unroll
                    for (int v = 0; v < VEC; ++v) p[v][k] = __fdiv_rn(s[v], static_cast<float>(fp.K));
                }
            }
        }
        if (probs) {
#pragma unroll
            for (int k = 0; k < MAXC; ++k) {
                if (k < fp.ncls) {
                    if constexpr (VEC == 4)
                        __stcs(reinterpret_cast<float4*>(probs + k * nvox + i), make_float4(p[0][k], p[1][k], p[2][k], p[3][k]));
                    else
                        probs[k * nvox + i] = p[0][k];
                }
            }
        }
        if (seg) {
            if constexpr (VEC == 4) {
                uint32_t packed = 0;
#pragma unroll
                for (int v = 0; v < VEC; ++v) packed |= static_cast<uint32_t>(decide_label<MAXC>(fp, p[v]) & 255) << (8 * v);
                *reinterpret_cast<uint32_t*>(seg + i) = packed;
            } else {
                seg[i] = static_cast<uint8_t>(decide_label<MAXC>(fp, p[0]));
            }
        }
    }
}

int fill_mirrors(MirrorSet* ms, const int* codes, int n) {
    BSG_REQUIRE(n >= 1 && n <= kMaxMirrors && codes != nullptr, "mirror count %d (1..8)", n);
    ms->n = n;
    for (int i = 0; i < kMaxMirrors; ++i) ms->code[i] = i < n ? (codes[i] & 7) : 0;
    return BSG_OK;
}

}  // namespace
}  // namespace bsg

using namespace bsg;

extern "C" {

int bsg_gather_patch_tta(const float* vol, int C, int Z, int Y, int X, int z0, int y0, int x0, int P0, int P1, int P2,
                         const int* mirror_codes_host, int nmirrors, void* out_bf16, int cpad, int out_f16, int kwpack,
                         void* stream) {
    BSG_REQUIRE(vol != nullptr && out_bf16 != nullptr, "null argument");
    BSG_REQUIRE(cpad % 8 == 0 && cpad >= C, "cpad %d must be a multiple of 8 and >= C=%d", cpad, C);
    BSG_REQUIRE(kwpack != 1 || (cpad == 16 && 3 * C <= 16 && (reinterpret_cast<uintptr_t>(out_bf16) & 31) == 0),
                "kwpack needs 3 * C <= 16, cpad == 16 and a 32-byte aligned output");
    BSG_REQUIRE(kwpack != 2 || (3 * C <= cpad && out_f16), "the fp16x3 split layout needs 3 * C <= cpad and fp16 output");
    BSG_REQUIRE(z0 >= 0 && y0 >= 0 && x0 >= 0 && z0 + P0 <= Z && y0 + P1 <= Y && x0 + P2 <= X,
                "tile exceeds the volume");
    MirrorSet ms;
    int rc = fill_mirrors(&ms, mirror_codes_host, nmirrors);
    if (rc != BSG_OK) return rc;
    dim3 grid(static_cast<unsigned>(ceil_div(P1 * P2, kThreads)), static_cast<unsigned>(P0));
    gather_patch_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        vol, C, Z, Y, X, z0, y0, x0, P0, P1, P2, ms, static_cast<__nv_bfloat16*>(out_bf16), cpad, out_f16, kwpack);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_norm_finalize(const double* stats, int N, int C, int groups, double count, float eps, const float* gamma,
                      const float* beta, float* scale_shift, void* stream) {
    BSG_REQUIRE(stats != nullptr && scale_shift != nullptr, "null argument");
    BSG_REQUIRE(groups >= 0 && (groups == 0 || C % groups == 0), "C %d not divisible by groups %d", C, groups);
    norm_finalize_kernel<<<ceil_div(N * C, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        stats, N, C, groups, count, eps, gamma, beta, scale_shift, 2, C, 0, 0.f);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_norm_finalize_table(const double* stats, int N, int C, int groups, double count, float eps, const float* gamma,
                            const float* beta, float slope, float* table, int ctot, int coff, void* stream) {
    BSG_REQUIRE(stats != nullptr && table != nullptr, "null argument");
    BSG_REQUIRE(groups >= 0 && (groups == 0 || C % groups == 0), "C %d not divisible by groups %d", C, groups);
    BSG_REQUIRE(coff >= 0 && coff + C <= ctot && (reinterpret_cast<uintptr_t>(table) & 15) == 0, "bad table slice");
    norm_finalize_kernel<<<ceil_div(N * C, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        stats, N, C, groups, count, eps, gamma, beta, table, 4, ctot, coff, slope);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_norm_apply_lrelu(void* x_bf16, size_t voxels_per_item, int N, int C, int ctot, int coff,
                         const float* scale_shift, float slope, int in_f16, int out_f16, void* stream) {
    BSG_REQUIRE(x_bf16 != nullptr && scale_shift != nullptr, "null argument");
    BSG_REQUIRE(C % 8 == 0 && ctot % 8 == 0 && coff % 8 == 0, "channel counts must be multiples of 8");
    // thread stride = blocks * 256 must be a multiple of the 16-byte groups per voxel (C / 8)
    const int c8n = C / 8;
    int g = c8n, r = kThreads;  // gcd(c8n, 256)
    while (r != 0) {
        const int t = g % r;
        g = r;
        r = t;
    }
    const int unit = c8n / g;  // smallest block count whose thread total is a multiple of c8n
    const size_t groups = voxels_per_item * static_cast<size_t>(c8n);
    size_t blocks = (groups + kThreads - 1) / kThreads;
    const size_t cap = static_cast<size_t>(sm_count_cached()) * 16 / (N > 0 ? N : 1) + 1;
    if (blocks > cap) blocks = cap;
    blocks = (blocks + unit - 1) / unit * unit;
    dim3 grid(static_cast<unsigned>(blocks), static_cast<unsigned>(N));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    __nv_bfloat16* xp = static_cast<__nv_bfloat16*>(x_bf16);
    if (in_f16 && out_f16)
        norm_apply_kernel<true, true><<<grid, kThreads, 0, st>>>(xp, voxels_per_item, C, ctot, coff, scale_shift, slope);
    else if (in_f16)
        norm_apply_kernel<true, false><<<grid, kThreads, 0, st>>>(xp, voxels_per_item, C, ctot, coff, scale_shift, slope);
    else if (out_f16)
        norm_apply_kernel<false, true><<<grid, kThreads, 0, st>>>(xp, voxels_per_item, C, ctot, coff, scale_shift, slope);
    else
        norm_apply_kernel<false, false><<<grid, kThreads, 0, st>>>(xp, voxels_per_item, C, ctot, coff, scale_shift, slope);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

static int head_launch(const void* feat_bf16, int feat_f16, int cfeat, int ctot, int P0, int P1, int P2,
                       const int* mirror_codes_host, int nmirrors, float mirror_weight, const float* head_w_host,
                       const float* head_b_host, int ncls, int nonlin, const float* gauss, float* acc, int Z, int Y, int X,
                       int z0, int y0, int x0, const float* norm_scale_shift, float norm_slope, int lo_off, void* stream) {
    BSG_REQUIRE(feat_bf16 != nullptr && head_w_host != nullptr && acc != nullptr, "null argument");
    BSG_REQUIRE(cfeat % 8 == 0 && cfeat >= 8 && cfeat <= kMaxHeadCh, "head input channels %d (8..64, multiple of 8)",
                cfeat);
    BSG_REQUIRE(ctot % 8 == 0, "ctot must be a multiple of 8");
    BSG_REQUIRE(ncls >= 1 && ncls <= kMaxClasses, "ncls %d (1..%d)", ncls, kMaxClasses);
    BSG_REQUIRE(nonlin >= 0 && nonlin <= 2, "nonlin %d", nonlin);
    BSG_REQUIRE(z0 >= 0 && y0 >= 0 && x0 >= 0 && z0 + P0 <= Z && y0 + P1 <= Y && x0 + P2 <= X,
                "tile exceeds the volume");
    MirrorSet ms;
    int rc = fill_mirrors(&ms, mirror_codes_host, nmirrors);
    if (rc != BSG_OK) return rc;
    HeadParams hp;
    memset(&hp, 0, sizeof(hp));
    hp.ncls = ncls;
    hp.cfeat = cfeat;
    hp.nonlin = nonlin;
    hp.mirror_weight = mirror_weight;
    for (int k = 0; k < ncls; ++k) {
        for (int c = 0; c < cfeat; ++c) hp.w[k][c] = head_w_host[k * cfeat + c];
        hp.b[k] = head_b_host ? head_b_host[k] : 0.f;
    }
    dim3 grid(static_cast<unsigned>(ceil_div(P1 * P2, kThreads)), static_cast<unsigned>(P0));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const __nv_bfloat16* fp = static_cast<const __nv_bfloat16*>(feat_bf16);
#define BSG_HEAD_LAUNCH(NC, CF)                                                                                     \
    head_tta_accumulate_kernel<NC, CF><<<grid, kThreads, 0, st>>>(fp, ctot, P0, P1, P2, ms, hp, feat_f16, gauss, acc, Z, \
                                                                  Y, X, z0, y0, x0, norm_scale_shift, norm_slope, lo_off)
    if (ncls <= 3 && cfeat <= 32)  // the BraTS region heads: 3 classes x 32 features — a fourth, padded class is 25 % more FMAs
        BSG_HEAD_LAUNCH(3, 32);
    else if (ncls <= 4 && cfeat <= 32)
        BSG_HEAD_LAUNCH(4, 32);
    else if (ncls <= 4)
        BSG_HEAD_LAUNCH(4, 64);
    else if (cfeat <= 32)
        BSG_HEAD_LAUNCH(8, 32);
    else
        BSG_HEAD_LAUNCH(8, 64);
#undef BSG_HEAD_LAUNCH
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_head_tta_accumulate(const void* feat_bf16, int feat_f16, int cfeat, int ctot, int P0, int P1, int P2,
                            const int* mirror_codes_host, int nmirrors, float mirror_weight,
                            const float* head_w_host, const float* head_b_host, int ncls, int nonlin,
                            const float* gauss, float* acc, int Z, int Y, int X, int z0, int y0, int x0,
                            const float* norm_scale_shift, float norm_slope, void* stream) {
    return head_launch(feat_bf16, feat_f16, cfeat, ctot, P0, P1, P2, mirror_codes_host, nmirrors, mirror_weight, head_w_host,
                       head_b_host, ncls, nonlin, gauss, acc, Z, Y, X, z0, y0, x0, norm_scale_shift, norm_slope, 0, stream);
}

int bsg_head_tta_accumulate_split(const void* feat16, int cfeat, int ctot, int P0, int P1, int P2,
                                  const int* mirror_codes_host, int nmirrors, float mirror_weight,
                                  const float* head_w_host, const float* head_b_host, int ncls, int nonlin,
                                  const float* gauss, float* acc, int Z, int Y, int X, int z0, int y0, int x0, void* stream) {
    BSG_REQUIRE(ctot >= 3 * cfeat, "split features need ctot >= 3 * cfeat");
    return head_launch(feat16, 1, cfeat, ctot, P0, P1, P2, mirror_codes_host, nmirrors, mirror_weight, head_w_host,
                       head_b_host, ncls, nonlin, gauss, acc, Z, Y, X, z0, y0, x0, nullptr, 0.f, 2 * cfeat, stream);
}

int bsg_norm_apply_lrelu_split(void* x, size_t voxels_per_item, int N, int C, int ctot, int coff, const float* scale_shift,
                               float slope, void* stream) {
    BSG_REQUIRE(x != nullptr && scale_shift != nullptr, "null argument");
    BSG_REQUIRE(C % 8 == 0 && ctot % 8 == 0 && coff % 8 == 0 && coff + 3 * C <= ctot, "bad split channel slice");
    const int c8n = C / 8;
    int g = c8n, r = kThreads;  // gcd(c8n, 256)
    while (r != 0) {
        const int t = g % r;
        g = r;
        r = t;
    }
    const int unit = c8n / g;
    const size_t groups = voxels_per_item * static_cast<size_t>(c8n);
    size_t blocks = (groups + kThreads - 1) / kThreads;
    const size_t cap = static_cast<size_t>(sm_count_cached()) * 16 / (N > 0 ? N : 1) + 1;
    if (blocks > cap) blocks = cap;
    blocks = (blocks + unit - 1) / unit * unit;
    dim3 grid(static_cast<unsigned>(blocks), static_cast<unsigned>(N));
    norm_apply_split_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<__half*>(x), voxels_per_item, C, ctot, coff, scale_shift, slope);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_finalize(const float* const* acc_list_host, int K, const float* wsum, int ncls, size_t nvox, int mode,
                 const int* order_host, float* probs, uint8_t* seg, void* stream) {
    BSG_REQUIRE(acc_list_host != nullptr && wsum != nullptr, "null argument");
    BSG_REQUIRE(K >= 1 && K <= 16, "K %d (1..16)", K);
    BSG_REQUIRE(ncls >= 1 && ncls <= kMaxClasses, "ncls %d", ncls);
    BSG_REQUIRE(mode == 0 || (mode == 1 && order_host != nullptr), "mode %d", mode);
    FinalizeParams fp;
    memset(&fp, 0, sizeof(fp));
    for (int j = 0; j < K; ++j) fp.acc[j] = acc_list_host[j];
    fp.K = K;
    fp.ncls = ncls;
    fp.mode = mode;
    for (int k = 0; k < ncls; ++k) fp.order[k] = order_host ? order_host[k] : k;
    if (nvox == 0) return BSG_OK;
    auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    bool vec = nvox % 4 == 0 && aligned16(wsum) && aligned16(probs) && (reinterpret_cast<uintptr_t>(seg) & 3) == 0;
    for (int j = 0; j < K; ++j) vec = vec && aligned16(acc_list_host[j]);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // one 4-voxel group per thread (no grid-stride tail): ~8.7 k blocks for a BraTS volume
    const size_t ngroups = vec ? nvox / 4 : nvox;
    const unsigned blocks = static_cast<unsigned>((ngroups + kThreads - 1) / kThreads);
#define BSG_FINALIZE_LAUNCH(MAXC)                                                              \
    do {                                                                                       \
        if (vec && K == 1)                                                                     \
            finalize_kernel<4, true, MAXC><<<blocks, kThreads, 0, s>>>(fp, wsum, nvox, probs, seg);  \
        else if (vec)                                                                          \
            finalize_kernel<4, false, MAXC><<<blocks, kThreads, 0, s>>>(fp, wsum, nvox, probs, seg); \
        else if (K == 1)                                                                       \
            finalize_kernel<1, true, MAXC><<<blocks, kThreads, 0, s>>>(fp, wsum, nvox, probs, seg);  \
        else                                                                                   \
            finalize_kernel<1, false, MAXC><<<blocks, kThreads, 0, s>>>(fp, wsum, nvox, probs, seg); \
    } while (0)
    if (ncls <= 4)
        BSG_FINALIZE_LAUNCH(4);
    else
        BSG_FINALIZE_LAUNCH(8);
#undef BSG_FINALIZE_LAUNCH
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

}  // extern "C"
