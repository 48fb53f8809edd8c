// GPU preprocessing of one case, the step right before predict_3D (trainer.preprocess_patient,
// run_brats2021_inference_singlethread.py:89; UPSTREAM nnU-Net v1 GenericPreprocessor / crop_to_nonzero, SURVEY.md
// Appendix A.8): non-zero mask over the modalities, binary_fill_holes, bounding box, per-channel z-score inside the
// mask ("nonCT" scheme, use_mask_for_norm), crop.  1 mm isotropic BraTS data: no resampling.
#include "bsg_common.cuh"

namespace bsg {
namespace {

constexpr int kThreads = 256;

// mask[i] = any_c(vol[c][i] != 0)      (create_nonzero_mask: `nonzero_mask = nonzero_mask | (data[c] != 0)`)
__global__ void __launch_bounds__(kThreads) nonzero_mask_kernel(const float* __restrict__ vol, int C, size_t n,
                                                                uint8_t* __restrict__ mask) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        bool nz = false;
        for (int c = 0; c < C; ++c) nz = nz || (__ldg(vol + c * n + i) != 0.f);
        mask[i] = nz ? 1 : 0;
    }
}

// flag[id] = 1 for every background component (6-connected labels of mask == 0) that touches a face of the volume
__global__ void __launch_bounds__(kThreads) border_flag_kernel(const int* __restrict__ labels, int d0, int d1, int d2,
                                                               uint8_t* __restrict__ flag) {
    const size_t n = static_cast<size_t>(d0) * d1 * d2;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int id = labels[i];
        if (id <= 0) continue;
        const int i2 = static_cast<int>(i % d2), i1 = static_cast<int>((i / d2) % d1),
                  i0 = static_cast<int>(i / (static_cast<size_t>(d2) * d1));
        if (i0 == 0 || i0 == d0 - 1 || i1 == 0 || i1 == d1 - 1 || i2 == 0 || i2 == d2 - 1) flag[id] = 1;
    }
}

// binary_fill_holes: a background voxel is a hole iff its component never reaches the border
__global__ void __launch_bounds__(kThreads) fill_holes_kernel(const int* __restrict__ labels,
                                                              const uint8_t* __restrict__ flag, size_t n,
                                                              uint8_t* __restrict__ mask) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int id = labels[i];
        if (id > 0 && !flag[id]) mask[i] = 1;
    }
}

// per channel: sum and sum of squares (fp64) and count of vol[c][i] over mask != 0
__global__ void __launch_bounds__(kThreads) masked_channel_stats_kernel(const float* __restrict__ vol, int C, size_t n,
                                                                        const uint8_t* __restrict__ mask,
                                                                        double* __restrict__ out /* [C][3] */) {
    __shared__ double sh[kThreads / 32][3];
    const int c = blockIdx.y;
    const float* v = vol + static_cast<size_t>(c) * n;
    double s1 = 0.0, s2 = 0.0, cnt = 0.0;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (mask[i]) {
            const double x = static_cast<double>(__ldg(v + i));
            s1 += x;
            s2 += x * x;
            cnt += 1.0;
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
        sh[w][0] = s1;
        sh[w][1] = s2;
        sh[w][2] = cnt;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int k = 0; k < kThreads / 32; ++k) t += sh[k][threadIdx.x];
        atomicAdd(out + c * 3 + threadIdx.x, t);
    }
}

// out[c][z][y][x] over the crop box: mask ? (v - mean_c) / (std_c + 1e-8) : 0, float32 arithmetic as numpy does it
__global__ void __launch_bounds__(kThreads) crop_normalize_kernel(const float* __restrict__ vol, int C, int Z, int Y, int X,
                                                                  const uint8_t* __restrict__ mask, int z0, int y0,
                                                                  int x0, int cz, int cy, int cx,
                                                                  const float* __restrict__ mean_std /* [C][2] */,
                                                                  float* __restrict__ out,
                                                                  uint8_t* __restrict__ mask_out) {
    const size_t cn = static_cast<size_t>(cz) * cy * cx;
    const size_t n = static_cast<size_t>(Z) * Y * X;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < cn; i += stride) {
        const int x = static_cast<int>(i % cx), y = static_cast<int>((i / cx) % cy),
                  z = static_cast<int>(i / (static_cast<size_t>(cx) * cy));
        const size_t src = (static_cast<size_t>(z0 + z) * Y + (y0 + y)) * X + (x0 + x);
        const bool m = mask[src] != 0;
        if (mask_out) mask_out[i] = m ? 1 : 0;
        for (int c = 0; c < C; ++c) {
            float r = 0.f;
            if (m) {
                const float mean = mean_std[2 * c], sd = mean_std[2 * c + 1];
                r = __fdiv_rn(__fsub_rn(__ldg(vol + c * n + src), mean), __fadd_rn(sd, 1e-8f));
            }
            out[c * cn + i] = r;
        }
    }
}

}  // namespace
}  // namespace bsg

using namespace bsg;

extern "C" {

int bsg_nonzero_mask(const float* vol, int C, int Z, int Y, int X, uint8_t* mask, void* stream) {
    BSG_REQUIRE(vol != nullptr && mask != nullptr && C > 0 && Z > 0 && Y > 0 && X > 0, "bad argument");
    const size_t n = static_cast<size_t>(Z) * Y * X;
    nonzero_mask_kernel<<<grid_for(n, kThreads, 16), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(vol, C, n, mask);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_fill_holes_u8(uint8_t* mask, int d0, int d1, int d2, int* labels, int* ncomp_dev, uint8_t* flags, size_t flags_cap,
                      void* workspace, size_t workspace_bytes, void* stream) {
    BSG_REQUIRE(mask != nullptr && labels != nullptr && ncomp_dev != nullptr && flags != nullptr && workspace != nullptr,
                "null argument");
    const size_t n = static_cast<size_t>(d0) * d1 * d2;
    BSG_REQUIRE(flags_cap >= n / 2 + 2, "flags buffer too small (%zu < %zu)", flags_cap, n / 2 + 2);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // 6-connected components of the background (value 0 -> maskbits bit 0); ids 1..ncomp <= n/2 + 1
    const int rc = bsg_ccl_stats(mask, d0, d1, d2, 1u, 6, labels, ncomp_dev, nullptr, 0, workspace, workspace_bytes, stream);
    if (rc != BSG_OK) return rc;
    BSG_CUDA_OK(cudaMemsetAsync(flags, 0, flags_cap, s));
    border_flag_kernel<<<grid_for(n, kThreads, 16), kThreads, 0, s>>>(labels, d0, d1, d2, flags);
    fill_holes_kernel<<<grid_for(n, kThreads, 16), kThreads, 0, s>>>(labels, flags, n, mask);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_masked_channel_stats(const float* vol, int C, size_t n, const uint8_t* mask, double* out, void* stream) {
    BSG_REQUIRE(vol != nullptr && mask != nullptr && out != nullptr && C > 0 && C <= 64, "bad argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    BSG_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(double) * 3 * C, s));
    dim3 grid(static_cast<unsigned>(grid_for(n, kThreads, 8)), static_cast<unsigned>(C));
    masked_channel_stats_kernel<<<grid, kThreads, 0, s>>>(vol, C, n, mask, out);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_crop_normalize(const float* vol, int C, int Z, int Y, int X, const uint8_t* mask, int z0, int y0, int x0, int cz,
                       int cy, int cx, const float* mean_std, float* out, uint8_t* mask_out, void* stream) {
    BSG_REQUIRE(vol != nullptr && mask != nullptr && mean_std != nullptr && out != nullptr, "null argument");
    BSG_REQUIRE(z0 >= 0 && y0 >= 0 && x0 >= 0 && cz > 0 && cy > 0 && cx > 0 && z0 + cz <= Z && y0 + cy <= Y && x0 + cx <= X,
                "crop box outside the volume");
    const size_t cn = static_cast<size_t>(cz) * cy * cx;
    crop_normalize_kernel<<<grid_for(cn, kThreads, 16), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        vol, C, Z, Y, X, mask, z0, y0, x0, cz, cy, cx, mean_std, out, mask_out);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

}  // extern "C"
