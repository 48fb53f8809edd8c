// Error plumbing shared by all translation units of libbrainseg_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include "../../include/brainseg_b200.h"

namespace bsg {

int set_error(int code, const char* fmt, ...);

#define BSG_CUDA_OK(expr)                                                                               \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return ::bsg::set_error(BSG_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                    __FILE__, __LINE__);                                                \
    } while (0)

#define BSG_REQUIRE(cond, ...)                                        \
    do {                                                              \
        if (!(cond)) return ::bsg::set_error(BSG_EINVAL, __VA_ARGS__); \
    } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
// SM count of the CURRENT device (one process may drive several: engines are keyed by device on the Python side)
inline int sm_count_cached() {
    static int n[kMaxDevices] = {0};
    const int dev = current_device();
    if (n[dev] == 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        n[dev] = v > 0 ? v : 148;
    }
    return n[dev];
}
// cudaFuncAttributeMaxDynamicSharedMemorySize applies per device: `done` is the per-kernel bit set of devices served
template <typename K>
inline cudaError_t ensure_max_smem(K kernel, unsigned long long* done, int bytes) {
    const int dev = current_device();
    if ((*done >> dev) & 1ull) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) *done |= 1ull << dev;
    return e;
}

// Blocks for a grid-stride kernel: enough to cover `work` items at `per_block` each, capped at `waves` resident
// blocks per SM (a multiple of the SM count, so the last wave is full).
inline int grid_for(size_t work, int per_block, int waves = 8) {
    size_t blocks = (work + per_block - 1) / per_block;
    const size_t cap = static_cast<size_t>(sm_count_cached()) * waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

}  // namespace bsg
