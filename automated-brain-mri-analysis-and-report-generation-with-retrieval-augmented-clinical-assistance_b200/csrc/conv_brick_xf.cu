// Brick conv kernel, instantiations WITH the in-consumer norm transform (XF; see conv_brick_kernel.cuh): the input is
// the raw output of an InstanceNorm / GroupNorm block and is normalised + LeakyReLU'd in shared memory on its way to
// the tensor core (reference: ConvDropoutNormNonlin.forward, model_architecture/generic_UNet.py:68-72).
#include "conv_brick_kernel.cuh"

namespace bsg {

cudaError_t launch_conv_brick_xf(const BrickArgs& a, int cc, int nt, int grid, size_t smem_bytes, cudaStream_t stream) {
    if (cc == 16) return cudaErrorInvalidValue;  // planner never asks for it
    if (nt == 32) {
        if (cc == 64) return launch_stats<64, 32, true>(a, grid, smem_bytes, stream);
        return launch_stats<32, 32, true>(a, grid, smem_bytes, stream);
    }
    if (cc == 64) return launch_stats<64, 64, true>(a, grid, smem_bytes, stream);
    return launch_stats<32, 64, true>(a, grid, smem_bytes, stream);
}

}  // namespace bsg
