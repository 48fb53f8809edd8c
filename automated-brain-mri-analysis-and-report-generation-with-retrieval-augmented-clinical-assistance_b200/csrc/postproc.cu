// Voxel post-processing kernels: label ensemble / remap, joint label histogram (Dice), 26-connected component
// labelling in SciPy raster order with per-component statistics, 6-connected surface count and masked moments.
// All HBM-bound byte/integer passes: 16-byte vectorised loads, shared-memory staged union-find, warp-aggregated
// atomics, grids sized in multiples of the SM count.  Results are bit-exact with the reference's NumPy/SciPy code.
#include "bsg_common.cuh"

namespace bsg {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// ---------------------------------------------------------------------------------------------- LUT kernels
struct Lut256 {
    uint8_t v[256];
};

__global__ void __launch_bounds__(kThreads) lut_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                          size_t n, const Lut256 lut) {
    __shared__ uint8_t s[256];
    if (threadIdx.x < 256) s[threadIdx.x] = lut.v[threadIdx.x];
    __syncthreads();
    const size_t nvec = n / 16;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint4 x = ld_stream(reinterpret_cast<const uint4*>(in) + i);
        uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t a = w[k];
            w[k] = static_cast<uint32_t>(s[a & 255]) | (static_cast<uint32_t>(s[(a >> 8) & 255]) << 8) |
                   (static_cast<uint32_t>(s[(a >> 16) & 255]) << 16) | (static_cast<uint32_t>(s[a >> 24]) << 24);
        }
        reinterpret_cast<uint4*>(out)[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (size_t i = nvec * 16 + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = s[in[i]];
}

// np.round((a+b)/2.0).astype(uint8): round-half-to-even of the mean of two labels
__device__ __forceinline__ uint32_t mean_half_even(uint32_t a, uint32_t b) {
    const uint32_t s = a + b;
    uint32_t r = s >> 1;
    if (s & 1u) r += (r & 1u);
    return r & 255u;
}

__global__ void __launch_bounds__(kThreads) pair_round_kernel(const uint8_t* __restrict__ a,
                                                              const uint8_t* __restrict__ b, uint8_t* __restrict__ out,
                                                              size_t n, const Lut256 post) {
    __shared__ uint8_t s[256];
    if (threadIdx.x < 256) s[threadIdx.x] = post.v[threadIdx.x];
    __syncthreads();
    const size_t nvec = n / 16;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint4 xa = ld_stream(reinterpret_cast<const uint4*>(a) + i);
        uint4 xb = ld_stream(reinterpret_cast<const uint4*>(b) + i);
        uint32_t wa[4] = {xa.x, xa.y, xa.z, xa.w}, wb[4] = {xb.x, xb.y, xb.z, xb.w}, wo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t o = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                o |= static_cast<uint32_t>(s[mean_half_even((wa[k] >> (8 * j)) & 255u, (wb[k] >> (8 * j)) & 255u)])
                     << (8 * j);
            wo[k] = o;
        }
        reinterpret_cast<uint4*>(out)[i] = make_uint4(wo[0], wo[1], wo[2], wo[3]);
    }
    for (size_t i = nvec * 16 + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = s[mean_half_even(a[i], b[i])];
}

// np.round(x).astype(np.uint8) for float inputs (labels arrive as float64 from get_fdata() in the reference)
template <typename T>
__global__ void __launch_bounds__(kThreads) round_to_u8_kernel(const T* __restrict__ in, uint8_t* __restrict__ out,
                                                               size_t n) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = static_cast<uint8_t>(static_cast<long long>(rint(static_cast<double>(in[i]))));
}

// ---------------------------------------------------------------------------------------------- joint histogram
// hist[p*16+g] += #voxels with pred==p, gt==g (p,g < 16); bad += #voxels with a label >= 16.
__global__ void __launch_bounds__(kThreads) joint_hist_kernel(const uint8_t* __restrict__ pred,
                                                              const uint8_t* __restrict__ gt, size_t n,
                                                              unsigned long long* __restrict__ hist,
                                                              unsigned long long* __restrict__ bad) {
    __shared__ unsigned int sh[256];
    __shared__ unsigned int sbad;
    sh[threadIdx.x] = 0;
    if (threadIdx.x == 0) sbad = 0;
    __syncthreads();
    unsigned int zero_pairs = 0, nbad = 0;
    const size_t nvec = n / 16;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint4 xp = ld_stream(reinterpret_cast<const uint4*>(pred) + i);
        uint4 xg = ld_stream(reinterpret_cast<const uint4*>(gt) + i);
        uint32_t wp[4] = {xp.x, xp.y, xp.z, xp.w}, wg[4] = {xg.x, xg.y, xg.z, xg.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if ((wp[k] | wg[k]) == 0u) {  // background fast path: 4 (0,0) pairs, no atomics
                zero_pairs += 4;
                continue;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t p = (wp[k] >> (8 * j)) & 255u, g = (wg[k] >> (8 * j)) & 255u;
                if ((p | g) == 0u)
                    ++zero_pairs;
                else if ((p | g) < 16u)
                    atomicAdd(&sh[p * 16 + g], 1u);
                else
                    ++nbad;
            }
        }
    }
    for (size_t i = nvec * 16 + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t p = pred[i], g = gt[i];
        if ((p | g) == 0u)
            ++zero_pairs;
        else if ((p | g) < 16u)
            atomicAdd(&sh[p * 16 + g], 1u);
        else
            ++nbad;
    }
    zero_pairs = __reduce_add_sync(0xffffffffu, zero_pairs);
    nbad = __reduce_add_sync(0xffffffffu, nbad);
    if ((threadIdx.x & 31) == 0) {
        if (zero_pairs) atomicAdd(&sh[0], zero_pairs);
        if (nbad) atomicAdd(&sbad, nbad);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], static_cast<unsigned long long>(sh[threadIdx.x]));
    if (threadIdx.x == 0 && sbad) atomicAdd(bad, static_cast<unsigned long long>(sbad));
}

// ---------------------------------------------------------------------------------------------- fused label pass
// out = post[round_half_even((a + b) / 2)] (run_brats2021_inference_singlethread.py:305 + convert_labels_to_brats.py)
// AND the joint histogram of (out, gt) (evaluate_segmentation.py:25-32) in ONE read of the three label volumes: the
// three separate passes are launch-bound at BraTS size (24 us each for 18-27 MB).  Two 16-byte groups per thread and
// step (six independent loads in flight).
__global__ void __launch_bounds__(kThreads) ensemble_hist_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                                 const uint8_t* __restrict__ gt, uint8_t* __restrict__ out,
                                                                 size_t n, const Lut256 post,
                                                                 unsigned long long* __restrict__ hist,
                                                                 unsigned long long* __restrict__ bad) {
    __shared__ uint8_t s[256];
    __shared__ unsigned int sh[256];
    __shared__ unsigned int sbad;
    s[threadIdx.x] = post.v[threadIdx.x];
    sh[threadIdx.x] = 0;
    if (threadIdx.x == 0) sbad = 0;
    __syncthreads();
    unsigned int zero_pairs = 0, nbad = 0;
    auto tally = [&](uint32_t p, uint32_t g) {
        if ((p | g) == 0u)
            ++zero_pairs;
        else if ((p | g) < 16u)
            atomicAdd(&sh[p * 16 + g], 1u);
        else
            ++nbad;
    };
    auto group = [&](const uint4 xa, const uint4 xb, const uint4 xg, size_t i) {
        const uint32_t wa[4] = {xa.x, xa.y, xa.z, xa.w}, wb[4] = {xb.x, xb.y, xb.z, xb.w}, wg[4] = {xg.x, xg.y, xg.z, xg.w};
        uint32_t wo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t o = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                o |= static_cast<uint32_t>(s[mean_half_even((wa[k] >> (8 * j)) & 255u, (wb[k] >> (8 * j)) & 255u)])
                     << (8 * j);
            wo[k] = o;
            if ((o | wg[k]) == 0u) {
                zero_pairs += 4;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) tally((o >> (8 * j)) & 255u, (wg[k] >> (8 * j)) & 255u);
            }
        }
        reinterpret_cast<uint4*>(out)[i] = make_uint4(wo[0], wo[1], wo[2], wo[3]);
    };
    const size_t nvec = n / 16;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    const uint4 *a4 = reinterpret_cast<const uint4*>(a), *b4 = reinterpret_cast<const uint4*>(b),
                *g4 = reinterpret_cast<const uint4*>(gt);
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; i + stride < nvec; i += 2 * stride) {
        const uint4 xa0 = ld_stream(a4 + i), xb0 = ld_stream(b4 + i), xg0 = ld_stream(g4 + i);
        const uint4 xa1 = ld_stream(a4 + i + stride), xb1 = ld_stream(b4 + i + stride), xg1 = ld_stream(g4 + i + stride);
        group(xa0, xb0, xg0, i);
        group(xa1, xb1, xg1, i + stride);
    }
    if (i < nvec) group(ld_stream(a4 + i), ld_stream(b4 + i), ld_stream(g4 + i), i);
    for (size_t t = nvec * 16 + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += stride) {
        const uint32_t o = s[mean_half_even(a[t], b[t])];
        out[t] = static_cast<uint8_t>(o);
        tally(o, gt[t]);
    }
    zero_pairs = __reduce_add_sync(0xffffffffu, zero_pairs);
    nbad = __reduce_add_sync(0xffffffffu, nbad);
    if ((threadIdx.x & 31) == 0) {
        if (zero_pairs) atomicAdd(&sh[0], zero_pairs);
        if (nbad) atomicAdd(&sbad, nbad);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], static_cast<unsigned long long>(sh[threadIdx.x]));
    if (threadIdx.x == 0 && sbad) atomicAdd(bad, static_cast<unsigned long long>(sbad));
}

// ---------------------------------------------------------------------------------------------- CCL (26-conn)
// Foreground test: bit v of maskbits set <=> label value v (< 32) is foreground; labels >= 32 are background.
__device__ __forceinline__ bool is_fg(uint8_t v, uint32_t maskbits) { return v < 32 && ((maskbits >> v) & 1u); }

__device__ __forceinline__ int uf_find(const int* L, int x) {
    int p = L[x];
    while (p != x) {
        x = p;
        p = L[x];
    }
    return x;
}
__device__ __forceinline__ int uf_find_volatile(int* L, int x) {
    int p = reinterpret_cast<volatile int*>(L)[x];
    while (p != x) {
        x = p;
        p = reinterpret_cast<volatile int*>(L)[x];
    }
    return x;
}
// link the larger root under the smaller one: roots end up being each component's minimum raster index
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    bool done;
    do {
        a = uf_find_volatile(L, a);
        b = uf_find_volatile(L, b);
        if (a < b) {
            const int old = atomicMin(&L[b], a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            const int old = atomicMin(&L[a], b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

// tile of the block-local pass: 4 x 8 x 64 voxels (d2 fastest), 256 threads x 8 voxels
constexpr int kT0 = 4, kT1 = 8, kT2 = 64, kTileVox = kT0 * kT1 * kT2;

// Pass 1: union-find inside each tile, entirely in shared memory; writes L[i] = global index of the local root
// (or -1 for background).
// maxd: largest |o0|+|o1|+|o2| of a neighbour offset — 1: 6-connected (scipy.ndimage default structure, used by
// binary_fill_holes), 2: 18-connected, 3: 26-connected.
__global__ void __launch_bounds__(kThreads) ccl_local_kernel(const uint8_t* __restrict__ vol, uint32_t maskbits, int d0,
                                                             int d1, int d2, int* __restrict__ L, int maxd) {
    __shared__ int sl[kTileVox];
    __shared__ uint8_t sf[kTileVox];
    const int t2n = (d2 + kT2 - 1) / kT2, t1n = (d1 + kT1 - 1) / kT1, t0n = (d0 + kT0 - 1) / kT0;
    const int ntiles = t0n * t1n * t2n;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b2 = (tile % t2n) * kT2, b1 = ((tile / t2n) % t1n) * kT1, b0 = (tile / (t2n * t1n)) * kT0;
        for (int l = threadIdx.x; l < kTileVox; l += kThreads) {
            const int i2 = l % kT2, i1 = (l / kT2) % kT1, i0 = l / (kT2 * kT1);
            const int g0 = b0 + i0, g1 = b1 + i1, g2 = b2 + i2;
            bool fg = false;
            if (g0 < d0 && g1 < d1 && g2 < d2)
                fg = is_fg(vol[(static_cast<size_t>(g0) * d1 + g1) * d2 + g2], maskbits);
            sf[l] = fg;
            sl[l] = fg ? l : -1;
        }
        __syncthreads();
        for (int l = threadIdx.x; l < kTileVox; l += kThreads) {
            if (!sf[l]) continue;
            const int i2 = l % kT2, i1 = (l / kT2) % kT1, i0 = l / (kT2 * kT1);
            // 13 backward neighbours (smaller raster index)
#pragma unroll
            for (int k = 0; k < 13; ++k) {
                const int o0 = k < 9 ? -1 : 0;
                const int o1 = k < 9 ? (k / 3 - 1) : (k < 12 ? -1 : 0);
                const int o2 = k < 9 ? (k % 3 - 1) : (k < 12 ? (k - 9 - 1) : -1);
                if ((o0 != 0) + (o1 != 0) + (o2 != 0) > maxd) continue;
                const int j0 = i0 + o0, j1 = i1 + o1, j2 = i2 + o2;
                if (j0 < 0 || j1 < 0 || j1 >= kT1 || j2 < 0 || j2 >= kT2) continue;
                const int m = (j0 * kT1 + j1) * kT2 + j2;
                if (sf[m]) uf_union(sl, l, m);
            }
        }
        __syncthreads();
        for (int l = threadIdx.x; l < kTileVox; l += kThreads) {
            const int i2 = l % kT2, i1 = (l / kT2) % kT1, i0 = l / (kT2 * kT1);
            const int g0 = b0 + i0, g1 = b1 + i1, g2 = b2 + i2;
            if (g0 < d0 && g1 < d1 && g2 < d2) {
                int out = -1;
                if (sf[l]) {
                    const int r = uf_find(sl, l);
                    const int r2 = r % kT2, r1 = (r / kT2) % kT1, r0 = r / (kT2 * kT1);
                    out = ((b0 + r0) * d1 + (b1 + r1)) * d2 + (b2 + r2);
                }
                L[(static_cast<size_t>(g0) * d1 + g1) * d2 + g2] = out;
            }
        }
        __syncthreads();
    }
}

// Pass 2: merge across tile faces with global atomics (only voxels on a low face of their tile do any work).
__global__ void __launch_bounds__(kThreads) ccl_border_kernel(int d0, int d1, int d2, int* __restrict__ L, int maxd) {
    const size_t n = static_cast<size_t>(d0) * d1 * d2;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (L[i] < 0) continue;
        const int i2 = static_cast<int>(i % d2), i1 = static_cast<int>((i / d2) % d1),
                  i0 = static_cast<int>(i / (static_cast<size_t>(d2) * d1));
        const bool f0 = (i0 % kT0) == 0, f1lo = (i1 % kT1) == 0, f1hi = (i1 % kT1) == kT1 - 1, f2lo = (i2 % kT2) == 0,
                   f2hi = (i2 % kT2) == kT2 - 1;
        if (!(f0 || f1lo || f1hi || f2lo || f2hi)) continue;
#pragma unroll
        for (int k = 0; k < 13; ++k) {
            const int o0 = k < 9 ? -1 : 0;
            const int o1 = k < 9 ? (k / 3 - 1) : (k < 12 ? -1 : 0);
            const int o2 = k < 9 ? (k % 3 - 1) : (k < 12 ? (k - 9 - 1) : -1);
            if ((o0 != 0) + (o1 != 0) + (o2 != 0) > maxd) continue;
            const int j0 = i0 + o0, j1 = i1 + o1, j2 = i2 + o2;
            if (j0 < 0 || j1 < 0 || j1 >= d1 || j2 < 0 || j2 >= d2) continue;
            // same tile => already merged by the local pass
            if (j0 / kT0 == i0 / kT0 && j1 / kT1 == i1 / kT1 && j2 / kT2 == i2 / kT2) continue;
            const size_t m = (static_cast<size_t>(j0) * d1 + j1) * d2 + j2;
            if (L[m] >= 0) uf_union(L, static_cast<int>(i), static_cast<int>(m));
        }
    }
}

// Pass 3: flatten + count roots per chunk (chunk = kChunk consecutive voxels in raster order).
constexpr int kChunk = 2048;
__global__ void __launch_bounds__(kThreads) ccl_flatten_count_kernel(size_t n, int* __restrict__ L,
                                                                     int* __restrict__ chunk_roots) {
    __shared__ int scount;
    const size_t nchunks = (n + kChunk - 1) / kChunk;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        if (threadIdx.x == 0) scount = 0;
        __syncthreads();
        int mine = 0;
        for (int k = threadIdx.x; k < kChunk; k += kThreads) {
            const size_t i = c * kChunk + k;
            if (i < n && L[i] >= 0) {
                const int r = uf_find(L, static_cast<int>(i));
                L[i] = r;  // roots are fixed points, so concurrent flattening is benign
                mine += (r == static_cast<int>(i));
            }
        }
        mine = __reduce_add_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&scount, mine);
        __syncthreads();
        if (threadIdx.x == 0) chunk_roots[c] = scount;
        __syncthreads();
    }
}

// Pass 4: exclusive scan of the per-chunk root counts (single block; <= a few thousand chunks).
__global__ void __launch_bounds__(1024) ccl_scan_kernel(int nchunks, int* __restrict__ chunk_roots,
                                                        int* __restrict__ ncomp) {
    __shared__ int swarp[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nchunks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nchunks ? chunk_roots[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) swarp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = swarp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            swarp[threadIdx.x] = w;
        }
        __syncthreads();
        const int warp_off = (threadIdx.x >> 5) ? swarp[(threadIdx.x >> 5) - 1] : 0;
        const int incl = x + warp_off;
        if (i < nchunks) chunk_roots[i] = carry + incl - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *ncomp = carry;
}

// Pass 5: roots get their 1-based SciPy id = rank among roots in raster order; stored as -(id) - 1 ... we write
// ids into a separate array keyed by voxel: id_at_root[i] (only roots).  One warp-ballot scan per 32 voxels.
__global__ void __launch_bounds__(kThreads) ccl_assign_kernel(size_t n, const int* __restrict__ L,
                                                              const int* __restrict__ chunk_off,
                                                              int* __restrict__ labels) {
    __shared__ int swarp[kThreads / 32];
    const size_t nchunks = (n + kChunk - 1) / kChunk;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        int running = chunk_off[c];
        for (int base = 0; base < kChunk; base += kThreads) {
            const size_t i = c * kChunk + base + threadIdx.x;
            const bool root = (i < n) && (L[i] == static_cast<int>(i));
            const unsigned bal = __ballot_sync(0xffffffffu, root);
            const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
            if (lane == 0) swarp[w] = __popc(bal);
            __syncthreads();
            int off = 0, tot = 0;
#pragma unroll
            for (int k = 0; k < kThreads / 32; ++k) {
                const int cnt = swarp[k];
                if (k < w) off += cnt;
                tot += cnt;
            }
            if (root) labels[i] = running + off + __popc(bal & ((1u << lane) - 1u)) + 1;
            running += tot;
            __syncthreads();
        }
    }
}

// Pass 6: every foreground voxel takes the id of its root; background -> 0.  Also accumulates the per-component
// statistics of feature_extraction/step3_multiplicity.py:63-121 with warp-aggregated atomics.
struct CompStats {  // one per component, 1-based id -> index id-1
    unsigned long long count, s0, s1, s2, n1, n2, n3;
    int mn0, mn1, mn2, mx0, mx1, mx2;
    int pad0, pad1;
};
static_assert(sizeof(CompStats) == 88, "CompStats layout is part of the C ABI");

__global__ void __launch_bounds__(kThreads) ccl_stats_init_kernel(CompStats* st, int cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) {
        CompStats z;
        z.count = z.s0 = z.s1 = z.s2 = z.n1 = z.n2 = z.n3 = 0;
        z.mn0 = z.mn1 = z.mn2 = 0x7fffffff;
        z.mx0 = z.mx1 = z.mx2 = -1;
        z.pad0 = z.pad1 = 0;
        st[i] = z;
    }
}

__global__ void __launch_bounds__(kThreads) ccl_finalize_kernel(const uint8_t* __restrict__ vol, int d0, int d1,
                                                                int d2, const int* __restrict__ L,
                                                                int* __restrict__ labels, CompStats* __restrict__ st,
                                                                int cap) {
    const size_t n = static_cast<size_t>(d0) * d1 * d2;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    const size_t nround = (n + 31) / 32 * 32;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nround; i += stride) {
        int id = 0;
        int root = -1;
        if (i < n) root = L[i];
        // ids of roots were written by ccl_assign_kernel before this kernel started; non-root voxels read them
        if (root >= 0) id = (root == static_cast<int>(i)) ? labels[i] : labels[root];
        // lanes that contribute statistics; ids beyond `cap` are only counted (the caller re-runs with a larger
        // table), and must not be named in the masks of the warp collectives below
        const bool part = st != nullptr && id > 0 && id <= cap;
        const unsigned active = __ballot_sync(0xffffffffu, part);
        if (i < n && root != static_cast<int>(i)) labels[i] = id;  // (roots already hold their id)
        if (active == 0u) continue;
        if (part) {
            const int i2 = static_cast<int>(i % d2), i1 = static_cast<int>((i / d2) % d1),
                      i0 = static_cast<int>(i / (static_cast<size_t>(d2) * d1));
            const uint8_t v = vol[i];
            const unsigned peers = __match_any_sync(active, id);
            const int leader = __ffs(peers) - 1;
            const int lane = threadIdx.x & 31;
            // reduce over the peer group
            unsigned cnt = __popc(peers);
            unsigned s0 = __reduce_add_sync(peers, static_cast<unsigned>(i0));
            unsigned s1 = __reduce_add_sync(peers, static_cast<unsigned>(i1));
            unsigned s2 = __reduce_add_sync(peers, static_cast<unsigned>(i2));
            unsigned n1 = __popc(__ballot_sync(peers, v == 1) & peers);
            unsigned n2 = __popc(__ballot_sync(peers, v == 2) & peers);
            unsigned n3 = __popc(__ballot_sync(peers, v == 3) & peers);
            int mn0 = __reduce_min_sync(peers, i0), mx0 = __reduce_max_sync(peers, i0);
            int mn1 = __reduce_min_sync(peers, i1), mx1 = __reduce_max_sync(peers, i1);
            int mn2 = __reduce_min_sync(peers, i2), mx2 = __reduce_max_sync(peers, i2);
            if (lane == leader) {
                CompStats* s = st + (id - 1);
                atomicAdd(&s->count, static_cast<unsigned long long>(cnt));
                atomicAdd(&s->s0, static_cast<unsigned long long>(s0));
                atomicAdd(&s->s1, static_cast<unsigned long long>(s1));
                atomicAdd(&s->s2, static_cast<unsigned long long>(s2));
                if (n1) atomicAdd(&s->n1, static_cast<unsigned long long>(n1));
                if (n2) atomicAdd(&s->n2, static_cast<unsigned long long>(n2));
                if (n3) atomicAdd(&s->n3, static_cast<unsigned long long>(n3));
                atomicMin(&s->mn0, mn0);
                atomicMin(&s->mn1, mn1);
                atomicMin(&s->mn2, mn2);
                atomicMax(&s->mx0, mx0);
                atomicMax(&s->mx1, mx1);
                atomicMax(&s->mx2, mx2);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- masked moments
// For each of nmask label sets (bit masks): count, first and second coordinate moments, bounding box, and the
// 6-connected surface count (voxels whose 6-neighbourhood leaves the mask or the image — SciPy binary_erosion with
// border_value=0, feature_extraction/step4_morphology.py:42-45).
struct MaskMoments {
    unsigned long long count, s0, s1, s2, s00, s11, s22, s01, s02, s12, surface;
    int mn0, mn1, mn2, mx0, mx1, mx2;
    int pad[2];
};
static_assert(sizeof(MaskMoments) == 120, "MaskMoments layout is part of the C ABI");
constexpr int kMaxMasks = 8;
struct MaskSet {
    uint32_t bits[kMaxMasks];
    uint32_t want_surface;  // bit m set: compute the surface count of mask m
    int nmask;
};

__global__ void __launch_bounds__(kThreads) moments_init_kernel(MaskMoments* out, int nmask) {
    const int i = threadIdx.x;
    if (i < nmask) {
        MaskMoments z;
        memset(&z, 0, sizeof(z));
        z.mn0 = z.mn1 = z.mn2 = 0x7fffffff;
        z.mx0 = z.mx1 = z.mx2 = -1;
        out[i] = z;
    }
}

__global__ void __launch_bounds__(kThreads) moments_kernel(const uint8_t* __restrict__ vol, int d0, int d1, int d2,
                                                           const MaskSet ms, MaskMoments* __restrict__ out) {
    const size_t n = static_cast<size_t>(d0) * d1 * d2;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    const size_t nround = (n + 31) / 32 * 32;
    const size_t s1 = d2, s0 = static_cast<size_t>(d1) * d2;
    // per-thread partials, flushed with warp reductions at the end
    unsigned long long acc[kMaxMasks][11];
    int bb[kMaxMasks][6];
#pragma unroll
    for (int m = 0; m < kMaxMasks; ++m) {
#pragma unroll
        for (int k = 0; k < 11; ++k) acc[m][k] = 0;
        bb[m][0] = bb[m][1] = bb[m][2] = 0x7fffffff;
        bb[m][3] = bb[m][4] = bb[m][5] = -1;
    }
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nround; i += stride) {
        if (i >= n) continue;
        const uint8_t v = vol[i];
        if (v == 0 || v >= 32) continue;
        const int i2 = static_cast<int>(i % d2), i1 = static_cast<int>((i / d2) % d1),
                  i0 = static_cast<int>(i / s0);
        uint8_t nb[6];
        bool have_nb = false;
#pragma unroll
        for (int m = 0; m < kMaxMasks; ++m) {
            if (m >= ms.nmask) break;
            if (!((ms.bits[m] >> v) & 1u)) continue;
            acc[m][0] += 1;
            acc[m][1] += i0;
            acc[m][2] += i1;
            acc[m][3] += i2;
            acc[m][4] += static_cast<unsigned long long>(i0) * i0;
            acc[m][5] += static_cast<unsigned long long>(i1) * i1;
            acc[m][6] += static_cast<unsigned long long>(i2) * i2;
            acc[m][7] += static_cast<unsigned long long>(i0) * i1;
            acc[m][8] += static_cast<unsigned long long>(i0) * i2;
            acc[m][9] += static_cast<unsigned long long>(i1) * i2;
            bb[m][0] = min(bb[m][0], i0);
            bb[m][1] = min(bb[m][1], i1);
            bb[m][2] = min(bb[m][2], i2);
            bb[m][3] = max(bb[m][3], i0);
            bb[m][4] = max(bb[m][4], i1);
            bb[m][5] = max(bb[m][5], i2);
            if ((ms.want_surface >> m) & 1u) {
                if (!have_nb) {
                    // out-of-image neighbours count as background (value 0)
                    nb[0] = i0 > 0 ? vol[i - s0] : 0;
                    nb[1] = i0 < d0 - 1 ? vol[i + s0] : 0;
                    nb[2] = i1 > 0 ? vol[i - s1] : 0;
                    nb[3] = i1 < d1 - 1 ? vol[i + s1] : 0;
                    nb[4] = i2 > 0 ? vol[i - 1] : 0;
                    nb[5] = i2 < d2 - 1 ? vol[i + 1] : 0;
                    have_nb = true;
                }
                bool interior = true;
#pragma unroll
                for (int k = 0; k < 6; ++k) interior = interior && is_fg(nb[k], ms.bits[m]);
                if (!interior) acc[m][10] += 1;
            }
        }
    }
#pragma unroll
    for (int m = 0; m < kMaxMasks; ++m) {
        if (m >= ms.nmask) break;
        const unsigned any = __ballot_sync(0xffffffffu, acc[m][0] != 0);
        if (any == 0u) continue;
        unsigned long long r[11];
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            unsigned long long x = acc[m][k];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            r[k] = x;
        }
        int b[6];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            b[k] = __reduce_min_sync(0xffffffffu, bb[m][k]);
            b[k + 3] = __reduce_max_sync(0xffffffffu, bb[m][k + 3]);
        }
        if ((threadIdx.x & 31) == 0) {
            unsigned long long* o = reinterpret_cast<unsigned long long*>(out + m);
#pragma unroll
            for (int k = 0; k < 11; ++k)
                if (r[k]) atomicAdd(o + k, r[k]);
            atomicMin(&out[m].mn0, b[0]);
            atomicMin(&out[m].mn1, b[1]);
            atomicMin(&out[m].mn2, b[2]);
            atomicMax(&out[m].mx0, b[3]);
            atomicMax(&out[m].mx1, b[4]);
            atomicMax(&out[m].mx2, b[5]);
        }
    }
}

}  // namespace
}  // namespace bsg

using namespace bsg;

extern "C" {

int bsg_label_lut_u8(const uint8_t* in, uint8_t* out, size_t n, const uint8_t* lut_host, void* stream) {
    BSG_REQUIRE(in != nullptr && out != nullptr && lut_host != nullptr, "null argument");
    if (n == 0) return BSG_OK;
    BSG_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "label volumes must be 16-byte aligned");
    Lut256 lut;
    memcpy(lut.v, lut_host, 256);
    lut_u8_kernel<<<grid_for(n / 16 + 1, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(in, out, n, lut);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_label_pair_round_u8(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n, const uint8_t* post_lut_host,
                            void* stream) {
    BSG_REQUIRE(a != nullptr && b != nullptr && out != nullptr, "null argument");
    if (n == 0) return BSG_OK;
    BSG_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) &
                 15) == 0,
                "label volumes must be 16-byte aligned");
    Lut256 lut;
    for (int i = 0; i < 256; ++i) lut.v[i] = post_lut_host ? post_lut_host[i] : static_cast<uint8_t>(i);
    pair_round_kernel<<<grid_for(n / 16 + 1, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, b, out, n,
                                                                                                          lut);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_label_pair_round_hist_u8(const uint8_t* a, const uint8_t* b, const uint8_t* gt, uint8_t* out, size_t n,
                                 const uint8_t* post_lut_host, unsigned long long* hist256, unsigned long long* bad,
                                 void* stream) {
    BSG_REQUIRE(a != nullptr && b != nullptr && gt != nullptr && out != nullptr && hist256 != nullptr && bad != nullptr,
                "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    BSG_CUDA_OK(cudaMemsetAsync(hist256, 0, 256 * sizeof(unsigned long long), s));
    BSG_CUDA_OK(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), s));
    if (n == 0) return BSG_OK;
    BSG_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(gt) |
                  reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "label volumes must be 16-byte aligned");
    Lut256 lut;
    for (int i = 0; i < 256; ++i) lut.v[i] = post_lut_host ? post_lut_host[i] : static_cast<uint8_t>(i);
    ensemble_hist_kernel<<<grid_for(n / 32 + 1, kThreads, 8), kThreads, 0, s>>>(a, b, gt, out, n, lut, hist256, bad);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_round_to_u8(const void* in, int dtype, uint8_t* out, size_t n, void* stream) {
    BSG_REQUIRE(in != nullptr && out != nullptr, "null argument");
    if (n == 0) return BSG_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int g = grid_for(n, kThreads);
    if (dtype == 0)
        round_to_u8_kernel<float><<<g, kThreads, 0, s>>>(static_cast<const float*>(in), out, n);
    else if (dtype == 1)
        round_to_u8_kernel<double><<<g, kThreads, 0, s>>>(static_cast<const double*>(in), out, n);
    else
        return set_error(BSG_EINVAL, "dtype %d (0 = f32, 1 = f64)", dtype);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_joint_hist_u8(const uint8_t* pred, const uint8_t* gt, size_t n, unsigned long long* hist256,
                      unsigned long long* bad, void* stream) {
    BSG_REQUIRE(pred != nullptr && gt != nullptr && hist256 != nullptr && bad != nullptr, "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    BSG_CUDA_OK(cudaMemsetAsync(hist256, 0, 256 * sizeof(unsigned long long), s));
    BSG_CUDA_OK(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), s));
    if (n == 0) return BSG_OK;
    BSG_REQUIRE(((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(gt)) & 15) == 0,
                "label volumes must be 16-byte aligned");
    joint_hist_kernel<<<grid_for(n / 16 + 1, kThreads * 4, 4), kThreads, 0, s>>>(pred, gt, n, hist256, bad);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

size_t bsg_ccl26_workspace_bytes(int d0, int d1, int d2) {
    const size_t n = static_cast<size_t>(d0) * d1 * d2;
    const size_t nchunks = (n + kChunk - 1) / kChunk;
    return n * sizeof(int) + (nchunks + 64) * sizeof(int);
}

int bsg_ccl26_stats(const uint8_t* vol, int d0, int d1, int d2, uint32_t maskbits, int* labels, int* ncomp_dev,
                    void* comp_stats, int stats_cap, void* workspace, size_t workspace_bytes, void* stream) {
    return bsg_ccl_stats(vol, d0, d1, d2, maskbits, 26, labels, ncomp_dev, comp_stats, stats_cap, workspace,
                         workspace_bytes, stream);
}

int bsg_ccl_stats(const uint8_t* vol, int d0, int d1, int d2, uint32_t maskbits, int connectivity, int* labels,
                  int* ncomp_dev, void* comp_stats, int stats_cap, void* workspace, size_t workspace_bytes,
                  void* stream) {
    BSG_REQUIRE(connectivity == 6 || connectivity == 18 || connectivity == 26, "connectivity %d (6, 18 or 26)",
                connectivity);
    const int maxd = connectivity == 6 ? 1 : (connectivity == 18 ? 2 : 3);
    BSG_REQUIRE(vol != nullptr && labels != nullptr && ncomp_dev != nullptr && workspace != nullptr, "null argument");
    BSG_REQUIRE(d0 > 0 && d1 > 0 && d2 > 0, "empty volume");
    const size_t n = static_cast<size_t>(d0) * d1 * d2;
    BSG_REQUIRE(n < (static_cast<size_t>(1) << 31), "volume too large for int32 labels");
    if (workspace_bytes < bsg_ccl26_workspace_bytes(d0, d1, d2))
        return set_error(BSG_ENOMEM, "workspace %zu < %zu bytes", workspace_bytes, bsg_ccl26_workspace_bytes(d0, d1, d2));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int* L = static_cast<int*>(workspace);
    int* chunk = L + n;
    const int nchunks = static_cast<int>((n + kChunk - 1) / kChunk);
    const int ntiles = ceil_div(d0, kT0) * ceil_div(d1, kT1) * ceil_div(d2, kT2);
    const int sms = sm_count_cached();
    ccl_local_kernel<<<ntiles < sms * 8 ? ntiles : sms * 8, kThreads, 0, s>>>(vol, maskbits, d0, d1, d2, L, maxd);
    ccl_border_kernel<<<grid_for(n, kThreads), kThreads, 0, s>>>(d0, d1, d2, L, maxd);
    ccl_flatten_count_kernel<<<nchunks < sms * 8 ? nchunks : sms * 8, kThreads, 0, s>>>(n, L, chunk);
    ccl_scan_kernel<<<1, 1024, 0, s>>>(nchunks, chunk, ncomp_dev);
    BSG_CUDA_OK(cudaMemsetAsync(labels, 0, n * sizeof(int), s));
    ccl_assign_kernel<<<nchunks < sms * 8 ? nchunks : sms * 8, kThreads, 0, s>>>(n, L, chunk, labels);
    CompStats* st = static_cast<CompStats*>(comp_stats);
    if (st != nullptr && stats_cap > 0)
        ccl_stats_init_kernel<<<ceil_div(stats_cap, kThreads), kThreads, 0, s>>>(st, stats_cap);
    ccl_finalize_kernel<<<grid_for(n, kThreads), kThreads, 0, s>>>(vol, d0, d1, d2, L, labels,
                                                                   stats_cap > 0 ? st : nullptr, stats_cap);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_masked_moments(const uint8_t* vol, int d0, int d1, int d2, const uint32_t* maskbits_host, int nmask,
                       uint32_t want_surface, void* out_moments, void* stream) {
    BSG_REQUIRE(vol != nullptr && maskbits_host != nullptr && out_moments != nullptr, "null argument");
    BSG_REQUIRE(nmask >= 1 && nmask <= kMaxMasks, "nmask %d (1..%d)", nmask, kMaxMasks);
    BSG_REQUIRE(d0 > 0 && d1 > 0 && d2 > 0, "empty volume");
    MaskSet ms;
    memset(&ms, 0, sizeof(ms));
    for (int i = 0; i < nmask; ++i) {
        BSG_REQUIRE((maskbits_host[i] & 1u) == 0, "label 0 cannot be foreground");
        ms.bits[i] = maskbits_host[i];
    }
    ms.want_surface = want_surface;
    ms.nmask = nmask;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MaskMoments* out = static_cast<MaskMoments*>(out_moments);
    moments_init_kernel<<<1, kThreads, 0, s>>>(out, nmask);
    const size_t n = static_cast<size_t>(d0) * d1 * d2;
    moments_kernel<<<grid_for(n, kThreads * 8, 2), kThreads, 0, s>>>(vol, d0, d1, d2, ms, out);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

}  // extern "C"
