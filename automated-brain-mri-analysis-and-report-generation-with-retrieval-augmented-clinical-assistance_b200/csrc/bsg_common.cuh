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
inline int sm_count_cached() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
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
