// Tile conv kernel (see conv_tc_kernel.cuh): instantiations with M blocking — two M tiles (adjacent output planes) per
// work item and weight stage, two accumulators per TMEM buffer; direct 16-bit epilogue.
#include "conv_tc_kernel.cuh"

namespace bsg {

cudaError_t launch_conv_tc_mb(const ConvArgs& a, int grid, size_t smem_bytes, cudaStream_t stream) {
    return launch_modes<kEpiDirect, 2>(a, grid, smem_bytes, stream);
}

}  // namespace bsg
