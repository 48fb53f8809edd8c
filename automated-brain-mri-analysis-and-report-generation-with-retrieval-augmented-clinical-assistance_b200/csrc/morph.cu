// Voxel operations of the feature-extraction steps that sit next to the hot path:
//   * iterated 6-connected binary erosion / dilation       (scipy.ndimage.binary_erosion / binary_dilation defaults:
//                                                            feature_extraction/step4_morphology.py:146,227,252-254)
//   * exact Euclidean distance transform                    (scipy.ndimage.distance_transform_edt, step4:160-161)
//   * gradient magnitude of the signed distance on the mask surface, summed (np.gradient + std/mean, step4:164-183)
//   * masked intensity moments, compaction and exact order statistics by radix select
//                                                           (feature_extraction/utils.py:27-68, step4:233-262,318-348)
// All HBM-bound or latency-bound passes over a <= 16 M voxel volume; none is GEMM-shaped.
#include <math_constants.h>

#include "bsg_common.cuh"

namespace bsg {
namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------ 6-connected morphology
// out = AND (erosion) / OR (dilation) of the voxel and its six face neighbours; outside the volume counts as 0
// (scipy border_value=0), so an erosion always peels the volume faces.
__global__ void __launch_bounds__(kThreads) morph6_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                          int d0, int d1, int d2, int dilate) {
    // 32-bit index arithmetic (the host checks n < 2^31): 64-bit div / mod per voxel cost more than the 7 byte loads
    const unsigned n = static_cast<unsigned>(d0) * d1 * d2;
    const unsigned s1 = static_cast<unsigned>(d2), s0 = static_cast<unsigned>(d1) * d2;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned row = i / s1;
        const int i2 = static_cast<int>(i - row * s1), i0 = static_cast<int>(row / d1),
                  i1 = static_cast<int>(row - static_cast<unsigned>(i0) * d1);
        const bool c = in[i] != 0;
        const bool a0 = i0 > 0 && in[i - s0] != 0, b0 = i0 + 1 < d0 && in[i + s0] != 0;
        const bool a1 = i1 > 0 && in[i - s1] != 0, b1 = i1 + 1 < d1 && in[i + s1] != 0;
        const bool a2 = i2 > 0 && in[i - 1] != 0, b2 = i2 + 1 < d2 && in[i + 1] != 0;
        const bool r = dilate ? (c || a0 || b0 || a1 || b1 || a2 || b2) : (c && a0 && b0 && a1 && b1 && a2 && b2);
        out[i] = r ? 1 : 0;
    }
}

// out = a & ~b (b may be null: out = a != 0), every byte normalised to 0 / 1
__global__ void __launch_bounds__(kThreads) mask_andnot_kernel(const uint8_t* __restrict__ a,
                                                               const uint8_t* __restrict__ b, size_t n,
                                                               uint8_t* __restrict__ out) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (a[i] != 0 && !(b != nullptr && b[i] != 0)) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------ exact EDT
// Squared distance to the nearest zero voxel, one axis at a time (axis 0, then 1, then 2 — the order in which SciPy
// accumulates (delta_k * sampling_k)^2, so anisotropic results round the same way):
//     out[i] = min_j ( f[j] + ((i - j) * s)^2 )      along the axis
// by brute force over j from shared memory: exact, branch-free, and at n <= 256 it is ~10 GFLOP fp64 for a BraTS
// volume, i.e. cheaper than the three passes' memory traffic.  A block owns a bundle of kBundle lines that are
// adjacent in memory (axes 0 / 1: neighbouring i2; axis 2: neighbouring rows, loaded along the row).
constexpr int kBundle = 16;
constexpr int kPitch = kBundle + 1;

struct EdtPass {
    int n;            // extent along the axis
    long long step;   // element stride along the axis
    long long nlines; // number of lines
    int inner;        // lines come in runs of `inner` consecutive elements (axes 0/1: d2 resp. d2; axis 2: 1)
    long long run_stride;  // start of run r = (r / runs_per_outer) * outer_stride + (r % runs_per_outer) * run_stride
    long long runs_per_outer;
    long long outer_stride;
    double sampling;
    int first;        // 1: input is the mask (f = mask ? +inf : 0)
    int last;         // 1: write sqrt
    int along_row;    // 1: the axis is the contiguous one (bundle = neighbouring rows)
};

__device__ __forceinline__ long long line_base(const EdtPass& p, long long line) {
    const long long run = line / p.inner, within = line - run * p.inner;
    return (run / p.runs_per_outer) * p.outer_stride + (run % p.runs_per_outer) * p.run_stride + within;
}

__global__ void __launch_bounds__(kThreads) edt_pass_kernel(const uint8_t* __restrict__ mask,
                                                            const double* __restrict__ fin, double* __restrict__ fout,
                                                            EdtPass p) {
    extern __shared__ double sh[];  // g[n][kPitch], then sq[n]
    double* g = sh;
    double* sq = sh + static_cast<size_t>(p.n) * kPitch;
    const long long nbundles = (p.nlines + kBundle - 1) / kBundle;
    for (int k = threadIdx.x; k < p.n; k += kThreads) {
        const double t = __dmul_rn(static_cast<double>(k), p.sampling);
        sq[k] = __dmul_rn(t, t);
    }
    for (long long b = blockIdx.x; b < nbundles; b += gridDim.x) {
        __syncthreads();
        // load: element (l, j) of the bundle
        for (int e = threadIdx.x; e < kBundle * p.n; e += kThreads) {
            const int l = p.along_row ? e / p.n : e % kBundle;
            const int j = p.along_row ? e % p.n : e / kBundle;
            const long long line = b * kBundle + l;
            double v = CUDART_INF;
            if (line < p.nlines) {
                const long long idx = line_base(p, line) + j * p.step;
                v = p.first ? (mask[idx] != 0 ? CUDART_INF : 0.0) : fin[idx];
            }
            g[j * kPitch + l] = v;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < kBundle * p.n; e += kThreads) {
            const int l = p.along_row ? e / p.n : e % kBundle;
            const int i = p.along_row ? e % p.n : e / kBundle;
            const long long line = b * kBundle + l;
            if (line >= p.nlines) continue;
            double best = CUDART_INF;
            for (int j = 0; j < p.n; ++j) {
                const int d = i - j;
                best = fmin(best, __dadd_rn(g[j * kPitch + l], sq[d < 0 ? -d : d]));
            }
            fout[line_base(p, line) + i * p.step] = p.last ? sqrt(best) : best;
        }
    }
}

// ------------------------------------------------------------------------------------------ border regularity
// Over the surface voxels (mask & ~erode6(mask)): g = |grad(dist_in - dist_out)| with np.gradient's differences
// (central inside, one-sided at the faces, unit spacing); out[0] += 1, out[1] += g - center, out[2] += (g - center)^2.
__device__ __forceinline__ double grad1(const double* __restrict__ a, const double* __restrict__ b, size_t i, int pos,
                                        int n, size_t step) {
    if (n < 2) return 0.0;
    if (pos == 0) return (a[i + step] - b[i + step]) - (a[i] - b[i]);
    if (pos == n - 1) return (a[i] - b[i]) - (a[i - step] - b[i - step]);
    return ((a[i + step] - b[i + step]) - (a[i - step] - b[i - step])) / 2.0;
}

__global__ void __launch_bounds__(kThreads) surface_gradient_kernel(const uint8_t* __restrict__ mask,
                                                                    const double* __restrict__ din,
                                                                    const double* __restrict__ dout, int d0, int d1,
                                                                    int d2, double center, double* __restrict__ out) {
    __shared__ double sh[kThreads / 32][3];
    const size_t n = static_cast<size_t>(d0) * d1 * d2;
    const size_t s1 = static_cast<size_t>(d2), s0 = static_cast<size_t>(d1) * d2;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    double c = 0.0, a1 = 0.0, a2 = 0.0;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (mask[i] == 0) continue;
        const int i2 = static_cast<int>(i % d2), i1 = static_cast<int>((i / d2) % d1), i0 = static_cast<int>(i / s0);
        const bool interior = i0 > 0 && i0 + 1 < d0 && i1 > 0 && i1 + 1 < d1 && i2 > 0 && i2 + 1 < d2 &&
                              mask[i - s0] != 0 && mask[i + s0] != 0 && mask[i - s1] != 0 && mask[i + s1] != 0 &&
                              mask[i - 1] != 0 && mask[i + 1] != 0;
        if (interior) continue;
        const double g0 = grad1(din, dout, i, i0, d0, s0), g1 = grad1(din, dout, i, i1, d1, s1),
                     g2 = grad1(din, dout, i, i2, d2, 1);
        const double g = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(g0, g0), __dmul_rn(g1, g1)), __dmul_rn(g2, g2)));
        const double t = g - center;
        c += 1.0;
        a1 += t;
        a2 += t * t;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
        sh[w][0] = c;
        sh[w][1] = a1;
        sh[w][2] = a2;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int k = 0; k < kThreads / 32; ++k) t += sh[k][threadIdx.x];
        atomicAdd(out + threadIdx.x, t);
    }
}

// ------------------------------------------------------------------------------------------ masked intensities
// float <-> order-preserving unsigned key
__device__ __forceinline__ uint32_t f2key(float x) {
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// selected(i) = mask != null ? mask[i] != 0 : data[i] > 0         (`data[mask > 0]`, `data[data > 0]`)
__device__ __forceinline__ bool selected(const float* __restrict__ data, const uint8_t* __restrict__ mask, size_t i) {
    return mask != nullptr ? mask[i] != 0 : data[i] > 0.f;
}

// out[0] += count, out[1] += sum(x - center), out[2] += sum((x - center)^2) in fp64; keys[0] = min key, keys[1] = max
__global__ void __launch_bounds__(kThreads) intensity_moments_kernel(const float* __restrict__ data,
                                                                  const uint8_t* __restrict__ mask, size_t n,
                                                                  double center, double* __restrict__ out,
                                                                  uint32_t* __restrict__ keys) {
    __shared__ double sh[kThreads / 32][3];
    __shared__ uint32_t shk[kThreads / 32][2];
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    double c = 0.0, a1 = 0.0, a2 = 0.0;
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (!selected(data, mask, i)) continue;
        const float x = data[i];
        const double t = static_cast<double>(x) - center;
        c += 1.0;
        a1 += t;
        a2 += t * t;
        const uint32_t k = f2key(x);
        kmin = min(kmin, k);
        kmax = max(kmax, k);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
        sh[w][0] = c;
        sh[w][1] = a1;
        sh[w][2] = a2;
        shk[w][0] = kmin;
        shk[w][1] = kmax;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int k = 0; k < kThreads / 32; ++k) t += sh[k][threadIdx.x];
        atomicAdd(out + threadIdx.x, t);
    }
    if (threadIdx.x == 32) {
        uint32_t a = 0xffffffffu, b = 0u;
        for (int k = 0; k < kThreads / 32; ++k) {
            a = min(a, shk[k][0]);
            b = max(b, shk[k][1]);
        }
        atomicMin(keys, a);
        atomicMax(keys + 1, b);
    }
}

__global__ void keys_to_float_kernel(const uint32_t* __restrict__ keys, int nk, float* __restrict__ out) {
    if (threadIdx.x < nk) out[threadIdx.x] = key2f(keys[threadIdx.x]);
}

// Order-free compaction of the selected values as sortable keys (warp-aggregated atomics on one counter).
__global__ void __launch_bounds__(kThreads) masked_compact_kernel(const float* __restrict__ data,
                                                                  const uint8_t* __restrict__ mask, size_t n,
                                                                  uint32_t* __restrict__ keys_out,
                                                                  unsigned long long* __restrict__ count) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    const size_t nround = (n + stride - 1) / stride * stride;  // whole warps stay in the loop for the ballots
    const int lane = threadIdx.x & 31;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nround; i += stride) {
        const bool sel = i < n && selected(data, mask, i);
        const uint32_t bal = __ballot_sync(0xffffffffu, sel);
        if (bal == 0) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(count, static_cast<unsigned long long>(__popc(bal)));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (sel) keys_out[base + __popc(bal & ((1u << lane) - 1u))] = f2key(data[i]);
    }
}

// Radix select, 8 bits per pass from the top, all requested ranks at once.
//   state[r] = {prefix (bits decided so far), remaining rank inside the prefix's bucket}
constexpr int kMaxRanks = 8;
struct SelectState {
    uint32_t prefix[kMaxRanks];
    unsigned long long rank[kMaxRanks];
};

__global__ void __launch_bounds__(kThreads) select_hist_kernel(const uint32_t* __restrict__ keys, size_t count,
                                                               const SelectState* __restrict__ st, int nranks,
                                                               int shift, uint32_t* __restrict__ hist /*[R][256]*/) {
    __shared__ uint32_t sh[kMaxRanks * 256];
    for (int k = threadIdx.x; k < nranks * 256; k += kThreads) sh[k] = 0;
    __syncthreads();
    const uint32_t himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
    uint32_t pre[kMaxRanks];
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r) pre[r] = r < nranks ? st->prefix[r] : 0u;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
        const uint32_t k = keys[i];
        const uint32_t digit = (k >> shift) & 255u;
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r)
            if (r < nranks && ((k ^ pre[r]) & himask) == 0u) atomicAdd(&sh[r * 256 + digit], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nranks * 256; k += kThreads)
        if (sh[k]) atomicAdd(&hist[k], sh[k]);
}

__global__ void select_pick_kernel(SelectState* __restrict__ st, int nranks, int shift, uint32_t* __restrict__ hist,
                                   float* __restrict__ out) {
    const int r = threadIdx.x;
    if (r < nranks) {
        unsigned long long k = st->rank[r], cum = 0;
        int b = 0;
        for (; b < 255; ++b) {
            const unsigned long long h = hist[r * 256 + b];
            if (cum + h > k) break;
            cum += h;
        }
        st->prefix[r] |= static_cast<uint32_t>(b) << shift;
        st->rank[r] = k - cum;
        if (shift == 0) out[r] = key2f(st->prefix[r]);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nranks * 256; k += blockDim.x) hist[k] = 0;
}

// count of selected voxels with  x1 < t1  &&  x2 > t2  &&  x3 < t3   (comparisons in fp64, as numpy does for a
// float64 array against a float64 scalar); a null x_k skips that test
__global__ void __launch_bounds__(kThreads) masked_threshold_count_kernel(const float* __restrict__ x1,
                                                                          const float* __restrict__ x2,
                                                                          const float* __restrict__ x3,
                                                                          const uint8_t* __restrict__ mask, size_t n,
                                                                          double t1, double t2, double t3,
                                                                          unsigned long long* __restrict__ out) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    unsigned long long c = 0;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (mask[i] == 0) continue;
        const bool ok = (x1 == nullptr || static_cast<double>(x1[i]) < t1) &&
                        (x2 == nullptr || static_cast<double>(x2[i]) > t2) &&
                        (x3 == nullptr || static_cast<double>(x3[i]) < t3);
        c += ok ? 1 : 0;
    }
    c = __reduce_add_sync(0xffffffffu, static_cast<unsigned>(c));
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

}  // namespace
}  // namespace bsg

using namespace bsg;

extern "C" {

int bsg_binary_morph6(const uint8_t* in, uint8_t* out, uint8_t* tmp, int d0, int d1, int d2, int dilate, int iterations,
                      void* stream) {
    BSG_REQUIRE(in != nullptr && out != nullptr && d0 > 0 && d1 > 0 && d2 > 0, "bad argument");
    BSG_REQUIRE(iterations >= 1, "iterations %d (repeat-until-stable is not supported)", iterations);
    BSG_REQUIRE(static_cast<size_t>(d0) * d1 * d2 < (1ull << 31), "volume too large (>= 2^31 voxels)");
    BSG_REQUIRE(in != out && (iterations == 1 || (tmp != nullptr && tmp != in && tmp != out)),
                "in / out / tmp must be distinct buffers (tmp is needed for iterations > 1)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t n = static_cast<size_t>(d0) * d1 * d2;
    const int grid = grid_for(n, kThreads, 16);
    // ping-pong so that the last iteration lands in `out`
    const uint8_t* src = in;
    for (int it = 0; it < iterations; ++it) {
        uint8_t* dst = ((iterations - 1 - it) % 2 == 0) ? out : tmp;
        morph6_kernel<<<grid, kThreads, 0, s>>>(src, dst, d0, d1, d2, dilate);
        src = dst;
    }
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_mask_andnot(const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out, void* stream) {
    BSG_REQUIRE(a != nullptr && out != nullptr, "null argument");
    mask_andnot_kernel<<<grid_for(n, kThreads, 16), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, b, n, out);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_edt(const uint8_t* mask, int d0, int d1, int d2, const double* sampling, double* out, double* tmp, void* stream) {
    BSG_REQUIRE(mask != nullptr && out != nullptr && tmp != nullptr && out != tmp && d0 > 0 && d1 > 0 && d2 > 0,
                "bad argument");
    const int nmax = d0 > d1 ? (d0 > d2 ? d0 : d2) : (d1 > d2 ? d1 : d2);
    const size_t smem_max = (static_cast<size_t>(nmax) * kPitch + nmax) * sizeof(double);
    BSG_REQUIRE(smem_max <= 200 * 1024, "extent %d too large for the EDT line buffer", nmax);
    static unsigned long long attr_done = 0;  // per device
    BSG_CUDA_OK(ensure_max_smem(edt_pass_kernel, &attr_done, 200 * 1024));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long D0 = d0, D1 = d1, D2 = d2;
    const double sm[3] = {sampling ? sampling[0] : 1.0, sampling ? sampling[1] : 1.0, sampling ? sampling[2] : 1.0};
    // axis 0: mask -> out; axis 1: out -> tmp; axis 2: tmp -> out (sqrt)
    EdtPass p[3];
    p[0] = {d0, D1 * D2, D1 * D2, static_cast<int>(D1 * D2 > 0x7fffffff ? 0 : D1 * D2), 0, 1, 0, sm[0], 1, 0, 0};
    p[1] = {d1, D2, D0 * D2, d2, 0, 1, D1 * D2, sm[1], 0, 0, 0};
    p[2] = {d2, 1, D0 * D1, 1, D2, D0 * D1, 0, sm[2], 0, 1, 1};
    BSG_REQUIRE(p[0].inner > 0, "plane too large");
    const double* src[3] = {nullptr, out, tmp};
    double* dst[3] = {out, tmp, out};
    for (int a = 0; a < 3; ++a) {
        const long long nb = (p[a].nlines + kBundle - 1) / kBundle;
        const int grid = static_cast<int>(nb < 4ll * sm_count_cached() ? nb : 4ll * sm_count_cached());
        const size_t smem = (static_cast<size_t>(p[a].n) * kPitch + p[a].n) * sizeof(double);
        edt_pass_kernel<<<grid, kThreads, smem, s>>>(mask, src[a], dst[a], p[a]);
    }
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_surface_gradient_sums(const uint8_t* mask, const double* dist_in, const double* dist_out, int d0, int d1, int d2,
                              double center, double* out3, void* stream) {
    BSG_REQUIRE(mask != nullptr && dist_in != nullptr && dist_out != nullptr && out3 != nullptr, "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    BSG_CUDA_OK(cudaMemsetAsync(out3, 0, 3 * sizeof(double), s));
    const size_t n = static_cast<size_t>(d0) * d1 * d2;
    surface_gradient_kernel<<<grid_for(n, kThreads, 8), kThreads, 0, s>>>(mask, dist_in, dist_out, d0, d1, d2, center,
                                                                           out3);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_intensity_moments(const float* data, const uint8_t* mask, size_t n, double center, double* out3, float* minmax,
                       void* stream) {
    BSG_REQUIRE(data != nullptr && out3 != nullptr && minmax != nullptr, "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    BSG_CUDA_OK(cudaMemsetAsync(out3, 0, 3 * sizeof(double), s));
    // the two floats of `minmax` double as the key accumulators: min starts at all ones, max at zero
    uint32_t* keys = reinterpret_cast<uint32_t*>(minmax);
    BSG_CUDA_OK(cudaMemsetAsync(keys, 0xff, sizeof(uint32_t), s));
    BSG_CUDA_OK(cudaMemsetAsync(keys + 1, 0, sizeof(uint32_t), s));
    intensity_moments_kernel<<<grid_for(n, kThreads, 8), kThreads, 0, s>>>(data, mask, n, center, out3, keys);
    keys_to_float_kernel<<<1, 32, 0, s>>>(keys, 2, minmax);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_masked_compact_keys(const float* data, const uint8_t* mask, size_t n, uint32_t* keys_out,
                            unsigned long long* count_dev, void* stream) {
    BSG_REQUIRE(data != nullptr && keys_out != nullptr && count_dev != nullptr, "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    BSG_CUDA_OK(cudaMemsetAsync(count_dev, 0, sizeof(unsigned long long), s));
    masked_compact_kernel<<<grid_for(n, kThreads, 8), kThreads, 0, s>>>(data, mask, n, keys_out, count_dev);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

size_t bsg_select_workspace_bytes(void) { return sizeof(SelectState) + kMaxRanks * 256 * sizeof(uint32_t); }

int bsg_select_ranks(const uint32_t* keys, size_t count, const unsigned long long* ranks_host, int nranks,
                     float* out_dev, void* workspace, size_t workspace_bytes, void* stream) {
    BSG_REQUIRE(keys != nullptr && ranks_host != nullptr && out_dev != nullptr && workspace != nullptr, "null argument");
    BSG_REQUIRE(nranks >= 1 && nranks <= kMaxRanks, "nranks %d (1..%d)", nranks, kMaxRanks);
    BSG_REQUIRE(workspace_bytes >= bsg_select_workspace_bytes(), "workspace too small");
    BSG_REQUIRE(count > 0, "empty selection");
    SelectState init;
    for (int r = 0; r < kMaxRanks; ++r) {
        init.prefix[r] = 0;
        init.rank[r] = 0;
    }
    for (int r = 0; r < nranks; ++r) {
        BSG_REQUIRE(ranks_host[r] < count, "rank %llu outside the %zu selected values", ranks_host[r], count);
        init.rank[r] = ranks_host[r];
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SelectState* st = static_cast<SelectState*>(workspace);
    uint32_t* hist = reinterpret_cast<uint32_t*>(st + 1);
    // the state is tiny: a synchronous-with-respect-to-host pageable copy is fine (the caller syncs for `count` anyway)
    BSG_CUDA_OK(cudaMemcpyAsync(st, &init, sizeof(init), cudaMemcpyHostToDevice, s));
    BSG_CUDA_OK(cudaMemsetAsync(hist, 0, kMaxRanks * 256 * sizeof(uint32_t), s));
    const int grid = grid_for(count, kThreads, 4);
    for (int shift = 24; shift >= 0; shift -= 8) {
        select_hist_kernel<<<grid, kThreads, 0, s>>>(keys, count, st, nranks, shift, hist);
        select_pick_kernel<<<1, 256, 0, s>>>(st, nranks, shift, hist, out_dev);
    }
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_masked_threshold_count(const float* x1, const float* x2, const float* x3, const uint8_t* mask, size_t n, double t1,
                               double t2, double t3, unsigned long long* out_dev, void* stream) {
    BSG_REQUIRE(mask != nullptr && out_dev != nullptr, "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    BSG_CUDA_OK(cudaMemsetAsync(out_dev, 0, sizeof(unsigned long long), s));
    masked_threshold_count_kernel<<<grid_for(n, kThreads, 8), kThreads, 0, s>>>(x1, x2, x3, mask, n, t1, t2, t3, out_dev);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

}  // extern "C"
