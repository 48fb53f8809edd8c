// Tile conv kernel (see conv_tc_kernel.cuh): instantiations with the plain 16-bit epilogue + the launcher.
#include "conv_tc_kernel.cuh"

namespace bsg {

size_t conv_tc_smem_bytes(const ConvArgs& a) {
    return static_cast<size_t>(a.nstages) * (a.a_stage_bytes + a.b_stage_bytes) + 1024 /*barriers*/ + 2048 /*bias*/ +
           1024 /*align*/ + (a.tma_out ? kTmaOutSmemBytes : 0);
}

cudaError_t launch_conv_tc(const ConvArgs& a, int grid, size_t smem_bytes, cudaStream_t stream) {
    if (a.split_stride != 0) return launch_conv_tc_split(a, grid, smem_bytes, stream);
    if (a.tma_out) return launch_conv_tc_tma(a, grid, smem_bytes, stream);
    if (a.mb == 2) return launch_conv_tc_mb(a, grid, smem_bytes, stream);
    return launch_modes<kEpiDirect>(a, grid, smem_bytes, stream);
}

}  // namespace bsg
