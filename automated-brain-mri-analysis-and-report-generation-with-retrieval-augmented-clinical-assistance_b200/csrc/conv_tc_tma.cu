// Tile conv kernel (see conv_tc_kernel.cuh): instantiations whose epilogue stages every warp's 32 voxels x 32 channels in
// shared memory and writes them with TMA tensor stores (bsg_conv_desc.tma_store = 1).
#include "conv_tc_kernel.cuh"

namespace bsg {

cudaError_t launch_conv_tc_tma(const ConvArgs& a, int grid, size_t smem_bytes, cudaStream_t stream) {
    return launch_modes<kEpiStage>(a, grid, smem_bytes, stream);
}

}  // namespace bsg
