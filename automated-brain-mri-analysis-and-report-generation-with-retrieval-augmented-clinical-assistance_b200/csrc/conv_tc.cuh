// Argument block shared by the host planner (conv_plan.cu) and the tcgen05 implicit-GEMM kernel (conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace bsg {

// One launch = one Conv3d(k3, s1|s2, p1) / ConvTranspose3d(k2,s2) / 1x1x1 GEMM layer over a whole batch of
// channels-last (N,D,H,W,C) bf16 activations.  GEMM view: M = output voxels (128 per tile, a bw x bh x bd x bn
// box), N = output channels (ntile per CTA pass), K = taps x input channels (cc per pipeline stage).
struct ConvArgs {
    CUtensorMap mapA[8];  // stride 1: [0] only.  stride 2: one half-resolution view per input parity (pd,ph,pw).
    CUtensorMap mapW;     // packed weights, dims (Cin_pad, Nrows, ntaps), tap order (kd, kw, kh)
    CUtensorMap mapWh;    // pair mode: box (cc, ntile / 2, 1) of the same tensor
    CUtensorMap mapO[8];  // tma_out: the output tensor (cout, Wo, Ho, Do, No) in TILE coordinates, box = one epilogue warp's
                          // 32 voxels x 32 channels; transposed conv: one double-strided view per output parity (pd,ph,pw)
    int tma_out;          // 1: epilogue stores through shared memory + TMA tensor stores
    int store_cols;       // tma_out: channels per store row, 64 (128-byte rows = whole lines) when cout_pad and the N tile
                          // are multiples of 64, else 32
    int pair;             // 1: launched as 2-CTA clusters; the two CTAs work on neighbouring M tiles of the same N tile
                          //    in lock-step, each fetches half of every weight stage and multicasts it to both
    int bw, bh, bd, bn;   // output tile box, bw*bh*bd*bn == 128
    int mb;               // M blocking: M tiles (adjacent planes) per work item and weight stage, 1 or 2 (2 needs bd == bn == 1,
                          // ntile <= 128: two accumulators x two buffers fill the 512 TMEM columns)
    int tw, th, td, tn;   // tile counts per dimension
    int n_ntiles, ntile;  // N tiling (ntile % 32 == 0, ntile <= 256)
    int Wo, Ho, Do, No;   // extents of the tile coordinate space
    int nchunks, cc;      // K chunks per tap; channels per chunk (16 / 32 / 64 -> swizzle 32B / 64B / 128B)
    int ntaps;            // 27 or 1
    int stride;           // 1 or 2
    int khshift;          // 1: an A stage holds bh+2 rows of h (box (cc, 8, bh+2, 1, 1)) and serves 3 kh taps
    int taps3;            // 1: a stage holds the three kh taps of one (kd, kw) as three separate activation boxes and one
                          //    3-tap weight box (any box shape, stride 1 or 2): 3x fewer, larger pipeline steps
    int nstages;
    uint32_t a_stage_bytes, b_stage_bytes;  // smem footprint of one stage, both multiples of 1024
    uint32_t stage_tx_bytes;                // bytes the two TMA boxes of a stage actually deliver
    uint32_t tmem_cols;                     // power of two >= 2*ntile
    // epilogue
    __nv_bfloat16* out;
    long long os_n, os_d, os_h, os_w;  // output element strides
    int out_mul;                       // 1; 2 = transposed-conv scatter (N tiles enumerate parity x Cout)
    int cout_pad;                      // padded Cout (per parity)
    int cout;                          // valid Cout
    int out_c_off;                     // channel offset inside the output tensor (concat buffers)
    const float* bias;                 // [cout_pad] or null
    float slope;
    int act;       // 0: none, 1: LeakyReLU(slope)
    double* stats;  // [No][cout][2] running (sum, sum of squares) of the pre-activation output, or null
    int out_f16;   // 1: store IEEE fp16 instead of bf16 (raw pre-norm outputs: 3 more mantissa bits, same bytes)
    int* overflow;  // device flag, set when a stored fp16 value left the fp16 range (null: no guard)
    int split_stride;  // != 0: fp16x3 split output [hi | hi | lo], blocks this many channels apart (conv_epilogue.cuh)
    int in_f16;    // 1: activations and weights are IEEE fp16 instead of bf16
};

constexpr unsigned kTmaOutSmemBytes = 8 * 2 * 4096 + 1024;  // two staging buffers per epilogue warp + alignment
cudaError_t launch_conv_tc(const ConvArgs& a, int grid, size_t smem_bytes, cudaStream_t stream);  // pair: grid even
cudaError_t launch_conv_tc_tma(const ConvArgs& a, int grid, size_t smem_bytes, cudaStream_t stream);
cudaError_t launch_conv_tc_mb(const ConvArgs& a, int grid, size_t smem_bytes, cudaStream_t stream);
cudaError_t launch_conv_tc_split(const ConvArgs& a, int grid, size_t smem_bytes, cudaStream_t stream);
size_t conv_tc_smem_bytes(const ConvArgs& a);

}  // namespace bsg
