// Brick conv kernel, instantiations WITHOUT the in-consumer norm transform + the launcher (see conv_brick_kernel.cuh).
#include "conv_brick_kernel.cuh"

namespace bsg {

size_t conv_brick_smem_bytes(const BrickArgs& a) {
    return static_cast<size_t>(a.nslabbuf) * a.slab_bytes + static_cast<size_t>(a.nstages) * a.a_stage_bytes +
           1024 /*barriers*/ + 256 /*bias*/ + 1024 /*align*/;
}

cudaError_t launch_conv_brick(const BrickArgs& a, int cc, int nt, int grid, size_t smem_bytes, cudaStream_t stream) {
    if (a.in_norm != nullptr) return launch_conv_brick_xf(a, cc, nt, grid, smem_bytes, stream);
    if (nt == 32) {
        if (cc == 64) return launch_stats<64, 32, false>(a, grid, smem_bytes, stream);
        if (cc == 32) return launch_stats<32, 32, false>(a, grid, smem_bytes, stream);
        return launch_stats<16, 32, false>(a, grid, smem_bytes, stream);
    }
    if (cc == 64) return launch_stats<64, 64, false>(a, grid, smem_bytes, stream);
    if (cc == 32) return launch_stats<32, 64, false>(a, grid, smem_bytes, stream);
    return launch_stats<16, 64, false>(a, grid, smem_bytes, stream);
}

}  // namespace bsg
