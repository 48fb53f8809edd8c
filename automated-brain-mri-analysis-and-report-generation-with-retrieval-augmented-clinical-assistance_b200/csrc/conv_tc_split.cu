// Tile conv kernel (see conv_tc_kernel.cuh): instantiations whose epilogue stores the fp32 result as the fp16x3 split
// [hi | hi | lo] (engine dtype "fp32": fp32-equivalent products as three fp16 MMAs; conv_epilogue.cuh).
#include "conv_tc_kernel.cuh"

namespace bsg {

cudaError_t launch_conv_tc_split(const ConvArgs& a, int grid, size_t smem_bytes, cudaStream_t stream) {
    return launch_modes<kEpiSplit>(a, grid, smem_bytes, stream);
}

}  // namespace bsg
