// Argument block of the "brick" variant of the tcgen05 conv kernel (conv_brick.cu) for the wide, shallow layers of
// Generic_UNet (stride-1 Conv3d k3, Cout <= 64: the full- and half-resolution levels, where most of the time goes).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace bsg {

// One work unit = a brick of P consecutive output planes (8 w x 16 h voxels each, one batch item).  All P fp32
// accumulators of a brick live side by side in TMEM (2 bricks double-buffered = 512 columns), so every activation
// box that TMA brings in (8 x 18 haloed rows of one input plane, one kw shift, one 16/32/64-channel chunk) feeds all
// nine (kd, kh) taps it takes part in — the three kd taps as ONE MMA of N = 3*NT over adjacent accumulators — and
// the weights sit in shared memory as per-phase slabs (phase = (channel chunk, kw); 9 taps x Cout x CC each) that
// stay resident when all of them fit.
struct BrickArgs {
    CUtensorMap mapA;  // 5-D (C, W, H, D, N) activations, box (CC, 8 | 10, 18, 1, 1), OOB zero fill = conv padding
    CUtensorMap mapW;  // 5-D (Cin_pad, Cout_pad, kh, kw, kd) view of the 27-tap weights, box (CC, NT, 1, 1, 3)
    int tw, th, tb, tn;  // unit grid: W/8, H/16, D/P, batch
    int P;               // planes per brick = 256 / NT
    int D;
    int nchunks;   // Cin / CC
    int kwn;       // taps along w: 3, or 1 for a 3x3x1 kernel (kw-packed first layer: the w neighbours sit in the channels)
    int nphases;   // kwn * nchunks, phase = chunk * kwn + kw
    int nslabbuf;  // weight slab buffers in shared memory; >= nphases: resident for the whole launch
    int kwf;       // 1: kw-fused — one haloed 10 w x 18 h box per (plane, chunk) serves all 27 taps (needs resident slabs)
    int nstages;   // activation ring depth
    uint32_t a_stage_bytes;  // multiple of 1024
    uint32_t a_tx_bytes;     // bytes one activation box delivers
    uint32_t slab_bytes;     // 9 * NT * CC * 2
    // epilogue
    __nv_bfloat16* out;
    long long os_n, os_d, os_h, os_w;
    int out_c_off;
    int cout, cout_pad;
    const float* bias;
    float slope;
    int act;
    double* stats;
    int out_f16;
    int* overflow;  // device flag, set when a stored fp16 value left the fp16 range (null: no guard)
    const float* in_norm;  // XF: [N][in_norm_c][4] (scale, shift, LeakyReLU slope, 0) applied to the INPUT in shared
                           // memory (the producer's deferred norm pass); mapA then carries NaN out-of-bounds fill
    int in_norm_c;         // channels per batch item of that table (the input buffer's slice width)
    int in_f16;  // 1: activations and weights are IEEE fp16 instead of bf16
};

constexpr int kBrickThreads = 224;  // 4 epilogue warps + activation producer + weight producer + MMA issuer
// Epilogue warp groups: CC == 16 (the 4-channel network input: few MMAs per plane, the epilogue is the bottleneck) runs a
// second group on warps 7..10.  The groups take alternate planes — except with norm statistics at NT = 64, where they
// take the two 32-column halves of every plane instead: a thread then keeps 64 statistic sums, not 128, which fits the
// 352-thread register budget (186) without falling back to the per-tile shuffle reduction (~4x the instructions).
__host__ __device__ constexpr int brick_epi_groups(int cc, int nt, bool stats) { return cc == 16 ? 2 : 1; }
// ... the XF instantiations use warps 7-10 for the in-consumer norm transform
__host__ __device__ constexpr int brick_threads(int cc, int nt, bool stats, bool xf) {
    return (brick_epi_groups(cc, nt, stats) == 2 || xf) ? kBrickThreads + 128 : kBrickThreads;
}
cudaError_t launch_conv_brick_xf(const BrickArgs& a, int cc, int nt, int grid, size_t smem_bytes, cudaStream_t stream);
cudaError_t launch_conv_brick(const BrickArgs& a, int cc, int nt, int grid, size_t smem_bytes, cudaStream_t stream);
size_t conv_brick_smem_bytes(const BrickArgs& a);

}  // namespace bsg
