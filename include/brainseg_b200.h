/*
 * brainseg_b200 — C ABI of libbrainseg_b200.so (sm_100a only).
 *
 * Drop-in boundary for the one hot path this repo accelerates: nnU-Net BraTS-2021 Generic_UNet sliding-window
 * inference (predict_3D) plus the voxel post-processing that consumes it.  The reference is pure Python and has no
 * FFI of its own; every entry point below names the reference call site (file:line under the reference tree) whose
 * arithmetic it replaces.  See INTEGRATION.md for the ctypes stubs a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success or a negative BSG_E* code; bsg_last_error() gives the message
 *     (thread-local).  Nothing throws across the ABI.
 *   - pointers are DEVICE pointers unless the name says host; the caller owns all memory (PyTorch tensors);
 *     the library never allocates or frees caller buffers (conv plans own only their descriptors).
 *   - `stream` is a cudaStream_t passed as void*.
 *   - activations are channels-last (N, D, H, W, C) bf16; a tensor may be a channel slice [c_off, c_off+C) of a
 *     wider buffer whose per-voxel channel count is c_tot (concat buffers, generic_UNet.py:438).
 *   - label volumes are uint8, C-order, any 3-D shape (d0, d1, d2).
 */
#ifndef BRAINSEG_B200_H
#define BRAINSEG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSG_OK 0
#define BSG_EINVAL (-1)   /* bad argument / unsupported shape */
#define BSG_ECUDA (-2)    /* CUDA runtime or driver error */
#define BSG_EARCH (-3)    /* device is not sm_100 */
#define BSG_ELABEL (-4)   /* label value outside the supported range */
#define BSG_ENOMEM (-5)   /* caller workspace too small */

int bsg_version(void);
/* Copies the calling thread's last error message into buf (NUL-terminated); returns its length. */
size_t bsg_last_error(char* buf, size_t cap);
/* 0 when the current device is compute capability 10.x, BSG_EARCH otherwise. */
int bsg_check_device(void);
int bsg_sm_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * 3-D conv stacks — Generic_UNet.forward, model_architecture/generic_UNet.py:423-446
 * ------------------------------------------------------------------------------------------------------------ */

enum { BSG_CONV_K3 = 0, BSG_CONVT_K2S2 = 1, BSG_CONV_K1 = 2 };
enum { BSG_ACT_NONE = 0, BSG_ACT_LRELU = 1 };

typedef struct bsg_conv_desc {
    int kind;            /* BSG_CONV_K3: nn.Conv3d k3 p1 (generic_UNet.py:56; stride 2 = conv pooling :285-288)
                            BSG_CONVT_K2S2: nn.ConvTranspose3d k2 s2 bias=False (:363-364)
                            BSG_CONV_K1: 1x1x1 conv */
    int stride;          /* 1 or 2 (K3 only) */
    int N, D, H, W;      /* INPUT extents */
    int cin;             /* input channels used (multiple of 16; zero-padded by the caller) */
    const void* in;      /* bf16, points at channel 0 of the slice */
    int in_ctot;         /* channels per voxel of the input buffer */
    int cout;            /* valid output channels */
    void* out;           /* bf16, points at channel 0 of the output buffer (not the slice) */
    int out_ctot;        /* channels per voxel of the output buffer */
    int out_coff;        /* first channel written */
    const void* weights; /* bf16 [ntaps][rows][cin], tap order (kd, kw, kh); rows = cout_pad (K3/K1) or
                            8*cout_pad (CONVT, parity-major (kd,kh,kw)); cout_pad = cout rounded up to 32 */
    const float* bias;   /* fp32 [cout_pad] or NULL */
    int act;             /* BSG_ACT_* applied after bias (used when the norm is folded / absent) */
    float slope;         /* LeakyReLU negative slope (generic_UNet.py:39) */
    double* stats;       /* fp64 [N][cout][2] += (sum, sum of squares) of the pre-activation output, or NULL;
                            feeds InstanceNorm / GroupNorm (generic_UNet.py:62-65).  fp64 so that the order in which
                            the CTAs' partial sums arrive cannot change the result (run-to-run reproducibility) */
    int out_f16;         /* 1: store the output as IEEE fp16 instead of bf16 (raw pre-norm values that
                            bsg_norm_apply_lrelu then rewrites in place as bf16) */
    int use_khshift;     /* -1 auto, 0 off, 1 on: halo reuse of the h taps inside shared memory */
    int max_ctas;        /* 0 = one CTA per SM */
    int in_f16;          /* 1: activations AND weights are IEEE fp16 instead of bf16 (same tensor-pipe rate; used for the
                            InstanceNorm / GroupNorm stacks, whose activations are bounded by construction) */
    int algo;            /* -1 auto, 0 tile kernel (one 128-voxel tile per accumulator), 1 brick kernel when the layer
                            suits it (stride-1 k3, Cout <= 64, W % 8 == 0, H % 16 == 0, D % (256/Cout_pad) == 0) */
    int pair;            /* tile kernel only: -1 auto, 0 off, 1 on when possible — launch as 2-CTA clusters whose CTAs work
                            on neighbouring tiles in lock-step and share every weight stage through TMA multicast */
    int* overflow;       /* device int or NULL: set to 1 when a value stored as fp16 (out_f16) left the fp16 range
                            (|x| > 65504): the caller's cue to re-plan the network in bf16 */
    const float* in_norm; /* NULL, or fp32 [N][in_norm_c][4] = (scale, shift, LeakyReLU slope, 0) per batch item and INPUT
                            channel: `in` is then the RAW output of an InstanceNorm / GroupNorm block whose
                            bsg_norm_apply_lrelu pass was skipped, and the conv applies y = lrelu(x*scale + shift) to
                            its input on the fly, in shared memory (generic_UNet.py:68-72 of the producing block;
                            rows written by bsg_norm_finalize_table).  Brick kernel only: bsg_conv_plan_create returns
                            BSG_EINVAL when the layer does not qualify and the caller keeps the separate pass. */
    int in_norm_c;       /* channels per batch item of the in_norm table (>= cin) */
    int in_norm_cc;      /* 0: planner's choice of K chunk for an in_norm plan; 32: force 32-channel chunks where the weight
                            slabs stay resident (measurement switch: slower) */
    int out_split_stride; /* 0, or S: fp32-equivalent mode — the fp32 result y is stored as the fp16 pair hi = fp16(y),
                            lo = fp16(y - hi) in THREE channel blocks [hi | hi | lo] at out_coff, out_coff + S,
                            out_coff + 2S; the next conv contracts them against weights stacked [w_hi | w_lo | w_hi] on its
                            input channels (y*w = hi*w_hi + hi*w_lo + lo*w_hi + O(2^-22)).  fp16 operands, tile kernel. */
    int kw_taps;         /* 0 / 3: 3x3x3 kernel.  1: 3x3x1 kernel (kd, kh taps only), weights [9 taps (kd, kh)][cout_pad][cin]:
                            the network's first conv on an input whose w neighbours were packed into the channels by
                            bsg_gather_patch_tta(kwpack = 1) — 9 taps of K = 16 instead of 27.  Brick kernel only. */
    int tma_store;       /* tile kernel epilogue through shared memory + TMA tensor stores (each epilogue warp stages its 32
                            voxels x 32 / 64 channels and writes them with one cp.async.bulk.tensor store).  -1 / 0: planner's
                            choice — on for transposed convs whose store rows are whole 128-byte lines (Cout_pad % 64 == 0),
                            where it measured 8-12 % faster than the direct per-thread rows; 1: on; 2: off. */
    int mblock;          /* tile kernel M blocking (two M tiles = adjacent planes per work item and weight stage; needs the 8 x 16
                            tile box and an N tile <= 128): -1 / 0 planner's choice (layers with enough tiles: from 4 waves of CTAs on at
                            stride 2, from 16 on at stride 1), 1 on where possible, 2 off */
} bsg_conv_desc;

typedef struct bsg_conv_plan bsg_conv_plan;

/* sizeof(bsg_conv_desc) as the library was compiled: lets a binding check its own struct layout. */
size_t bsg_conv_desc_size(void);
int bsg_conv_plan_create(const bsg_conv_desc* desc, bsg_conv_plan** plan);
int bsg_conv_plan_run(const bsg_conv_plan* plan, void* stream);
void bsg_conv_plan_destroy(bsg_conv_plan* plan);
/* Introspection for tests / roofline accounting: fills tile box, N tile, stages, grid, smem bytes, flops. */
typedef struct bsg_conv_info {
    int bw, bh, bd, bn, ntile, n_ntiles, cc, nstages, khshift, grid;
    size_t smem_bytes;
    double flops; /* algorithmic 2*MACs on valid channels */
} bsg_conv_info;
int bsg_conv_plan_info(const bsg_conv_plan* plan, bsg_conv_info* info);

/* ------------------------------------------------------------------------------------------------------------
 * Voxel post-processing (consumers of predict_3D's label volume)
 * ------------------------------------------------------------------------------------------------------------ */

/* out[i] = lut[in[i]].  Replaces the 4 boolean-mask passes of convert_labels_to_brats.py:34-55
 * (lut = {0,2,1,3,0...} for brats2025, {0,2,1,4,0...} for brats2021).  lut_host: 256 bytes of HOST memory. */
int bsg_label_lut_u8(const uint8_t* in, uint8_t* out, size_t n, const uint8_t* lut_host, void* stream);

/* out[i] = post[ round_half_even((a[i] + b[i]) / 2) ]: the two-model label ensemble
 * np.round((seg1+seg2)/2.0).astype(np.uint8) of run_brats2021_inference_singlethread.py:305, optionally fused with a
 * following label remap (post_lut_host: 256 host bytes, NULL = identity). */
int bsg_label_pair_round_u8(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n, const uint8_t* post_lut_host,
                            void* stream);

/* bsg_label_pair_round_u8 and bsg_joint_hist_u8(out, gt) in ONE pass over the three label volumes: the two-model
 * ensemble + remap (run_brats2021_inference_singlethread.py:305, convert_labels_to_brats.py:34-55) and the Dice bins of
 * the result against a ground truth (evaluate_segmentation.py:25-32). */
int bsg_label_pair_round_hist_u8(const uint8_t* a, const uint8_t* b, const uint8_t* gt, uint8_t* out, size_t n,
                                 const uint8_t* post_lut_host, unsigned long long* hist256, unsigned long long* bad,
                                 void* stream);

/* np.round(x).astype(np.uint8) (convert_labels_to_brats.py:37, feature_extraction/utils.py:169);
 * dtype 0 = float32, 1 = float64. */
int bsg_round_to_u8(const void* in, int dtype, uint8_t* out, size_t n, void* stream);

/* Joint histogram of two label volumes: hist256[p*16+g] = #voxels with pred==p and gt==g (p, g < 16); *bad counts
 * voxels carrying a label >= 16.  Every TP/FP/FN/TN of evaluate_segmentation.py:25-32 (per label) and of the WT/TC/ET
 * compounds (:129-151) is a sum of bins.  hist256 (256 x u64) and bad (1 x u64) are device buffers. */
int bsg_joint_hist_u8(const uint8_t* pred, const uint8_t* gt, size_t n, unsigned long long* hist256,
                      unsigned long long* bad, void* stream);

/* Per-component record written by bsg_ccl26_stats (88 bytes). */
typedef struct bsg_comp_stats {
    unsigned long long count;      /* comp_mask.sum()                       step3_multiplicity.py:65 */
    unsigned long long s0, s1, s2; /* coordinate sums -> centroids          :71-76 */
    unsigned long long n1, n2, n3; /* voxels with label 1 / 2 / 3           :104-110 */
    int mn0, mn1, mn2, mx0, mx1, mx2; /* bounding box                      :86-93 */
    int pad0, pad1;
} bsg_comp_stats;

/* 26-connected component labelling (scipy.ndimage.label with generate_binary_structure(3,3),
 * feature_extraction/step3_multiplicity.py:58-59, :222-223): foreground = voxels whose value v < 32 has bit v set in
 * maskbits (seg>0 -> 0xFFFFFFFE, seg==3 -> 1<<3).  labels (int32, same shape) receive SciPy's numbering: components
 * numbered 1.. in C-order raster order of their first voxel.  *ncomp_dev (device int) receives the component count.
 * comp_stats (device, stats_cap records, may be NULL with stats_cap 0) receives the per-component statistics; if the
 * count exceeds stats_cap only the first stats_cap components are recorded.  workspace: device scratch of at least
 * bsg_ccl26_workspace_bytes(). */
size_t bsg_ccl26_workspace_bytes(int d0, int d1, int d2);
/* Same with a choice of connectivity: 26 (above), 18 (generate_binary_structure(3, 2), step6_normal_structures.py:66-67)
 * or 6 (scipy.ndimage's default structure — the one binary_fill_holes uses in nnU-Net v1 create_nonzero_mask,
 * SURVEY App. A.8).  Label value 0 may be the foreground (maskbits bit 0). */
int bsg_ccl_stats(const uint8_t* vol, int d0, int d1, int d2, uint32_t maskbits, int connectivity, int* labels,
                  int* ncomp_dev, void* comp_stats, int stats_cap, void* workspace, size_t workspace_bytes,
                  void* stream);
int bsg_ccl26_stats(const uint8_t* vol, int d0, int d1, int d2, uint32_t maskbits, int* labels, int* ncomp_dev,
                    void* comp_stats, int stats_cap, void* workspace, size_t workspace_bytes, void* stream);

/* Per-mask record written by bsg_masked_moments (120 bytes). */
typedef struct bsg_mask_moments {
    unsigned long long count;                        /* mask.sum()            utils.py:181-183 */
    unsigned long long s0, s1, s2;                   /* centroid              utils.py:186-197 */
    unsigned long long s00, s11, s22, s01, s02, s12; /* -> np.cov             step4_morphology.py:100-104 */
    unsigned long long surface;                      /* (mask & ~binary_erosion(mask)).sum()   step4:42-45 */
    int mn0, mn1, mn2, mx0, mx1, mx2;                /* bounding box          utils.py:200-216 */
    int pad[2];
} bsg_mask_moments;

/* One pass over a label volume computing, for up to 8 label sets (bit masks, bit v = label v, v in 1..31), the voxel
 * count, first/second coordinate moments, bounding box and — for masks flagged in want_surface — the 6-connected
 * surface-voxel count with SciPy's border_value=0 semantics.  maskbits_host: nmask host words; out: device records. */
int bsg_masked_moments(const uint8_t* vol, int d0, int d1, int d2, const uint32_t* maskbits_host, int nmask,
                       uint32_t want_surface, void* out_moments, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Preprocessing of one case — trainer.preprocess_patient (run_brats2021_inference_singlethread.py:89); UPSTREAM nnU-Net
 * v1 crop_to_nonzero + GenericPreprocessor "nonCT" normalisation with use_mask_for_norm.
 * vol: fp32 (C, Z, Y, X).
 * ------------------------------------------------------------------------------------------------------------ */

/* mask[z][y][x] = any modality != 0 (create_nonzero_mask before hole filling). */
int bsg_nonzero_mask(const float* vol, int C, int Z, int Y, int X, uint8_t* mask, void* stream);
/* scipy.ndimage.binary_fill_holes(mask) in place: background voxels whose 6-connected component does not reach a face
 * of the volume become 1.  labels (int32, same shape), ncomp_dev (1 int), flags (>= n/2 + 2 bytes) and workspace
 * (bsg_ccl26_workspace_bytes) are device scratch. */
int bsg_fill_holes_u8(uint8_t* mask, int d0, int d1, int d2, int* labels, int* ncomp_dev, uint8_t* flags, size_t flags_cap,
                      void* workspace, size_t workspace_bytes, void* stream);
/* out[c] = (sum, sum of squares, count) in fp64 of vol[c][i] over mask[i] != 0 — data[c][mask].mean() / .std(). */
int bsg_masked_channel_stats(const float* vol, int C, size_t n, const uint8_t* mask, double* out, void* stream);
/* out[c] over the crop box (z0, y0, x0) + (cz, cy, cx): mask ? (v - mean_c) / (std_c + 1e-8) : 0 in float32 arithmetic
 * (mean_std: device fp32 [C][2]); mask_out (may be NULL) receives the cropped mask. */
int bsg_crop_normalize(const float* vol, int C, int Z, int Y, int X, const uint8_t* mask, int z0, int y0, int x0, int cz,
                       int cy, int cx, const float* mean_std, float* out, uint8_t* mask_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Voxel operations of the remaining feature-extraction steps.  Masks are uint8 volumes
 * (non-zero = set; outputs are 0 / 1), intensities fp32 volumes, all C-order (d0, d1, d2), device memory.
 * ------------------------------------------------------------------------------------------------------------ */

/* scipy.ndimage.binary_erosion / binary_dilation with the default 6-connected structure, border_value = 0 and
 * `iterations` >= 1 repetitions (feature_extraction/step4_morphology.py:146,227,252-254; step1:225; step2:373).
 * in / out / tmp are distinct buffers; tmp (same size) is only touched when iterations > 1. */
int bsg_binary_morph6(const uint8_t* in, uint8_t* out, uint8_t* tmp, int d0, int d1, int d2, int dilate, int iterations,
                      void* stream);
/* out = a & ~b (`dilated & ~wt_mask`, step4:228); b may be NULL (out = a != 0). */
int bsg_mask_andnot(const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out, void* stream);
/* scipy.ndimage.distance_transform_edt(mask, sampling) (step4:160-161, step6:206): out (fp64) = exact Euclidean
 * distance of every non-zero voxel to the nearest zero voxel, 0 on zero voxels; sampling: 3 host doubles or NULL (1,1,1).
 * Bit-exact with SciPy for isotropic sampling; +inf everywhere if the mask has no zero voxel (SciPy's result is
 * unspecified there).  tmp: fp64 scratch of the same size. */
int bsg_edt(const uint8_t* mask, int d0, int d1, int d2, const double* sampling, double* out, double* tmp, void* stream);
/* Border-regularity reduction (analyze_border_regularity, step4:146-176): over the surface voxels mask & ~erode6(mask),
 * g = |np.gradient(dist_in - dist_out)|; out3 (device fp64) = {count, sum(g - center), sum((g - center)^2)}. */
int bsg_surface_gradient_sums(const uint8_t* mask, const double* dist_in, const double* dist_out, int d0, int d1, int d2,
                              double center, double* out3, void* stream);
/* Intensity statistics of data[mask > 0] (get_intensity_stats, feature_extraction/utils.py:27-51) or, with mask NULL,
 * of data[data > 0] (utils.py:57,66; step4:318-320): out3 (device fp64) = {count, sum(x - center), sum((x - center)^2)},
 * minmax (device, 2 floats) = {min, max} (undefined when count == 0). */
int bsg_intensity_moments(const float* data, const uint8_t* mask, size_t n, double center, double* out3, float* minmax,
                          void* stream);
/* The selected values as order-preserving uint32 keys, unordered, into keys_out (capacity >= n); *count_dev = how many. */
int bsg_masked_compact_keys(const float* data, const uint8_t* mask, size_t n, uint32_t* keys_out,
                            unsigned long long* count_dev, void* stream);
/* Exact order statistics by 4-pass radix select: out_dev[r] (device floats) = the value of 0-based rank ranks_host[r]
 * (1 <= nranks <= 8) among the `count` keys — the two neighbours np.percentile / np.median interpolate between. */
size_t bsg_select_workspace_bytes(void);
int bsg_select_ranks(const uint32_t* keys, size_t count, const unsigned long long* ranks_host, int nranks,
                     float* out_dev, void* workspace, size_t workspace_bytes, void* stream);
/* *out_dev = number of voxels with mask != 0 and x1 < t1 and x2 > t2 and x3 < t3 (fp64 comparisons; a NULL x_k skips
 * its test) — the CSF-like signal count of analyze_cystic_vs_solid (step4:331-337). */
int bsg_masked_threshold_count(const float* x1, const float* x2, const float* x3, const uint8_t* mask, size_t n, double t1,
                               double t2, double t3, unsigned long long* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Sliding-window plumbing of predict_3D (nnU-Net v1 _internal_predict_3D_3Dconv_tiled /
 * _internal_maybe_mirror_and_pred_3D; call site run_brats2021_inference_singlethread.py:97-106)
 * ------------------------------------------------------------------------------------------------------------ */

/* Mirror codes: bit0 = flip x (tensor dim 4), bit1 = flip y (dim 3), bit2 = flip z (dim 2); the upstream order of the
 * 8 TTA passes is codes 0..7. */

/* out[m][d][h][w][c] (bf16, or fp16 when out_f16 = 1; cpad channels, zero padded) =
 * vol[c][z0+fz(d)][y0+fy(h)][x0+fx(w)]: the tile crop data[None, :, lb_x:ub_x, ...] plus torch.flip(x, axes) for every
 * mirror m, written as one channels-last batch. */
/* kwpack = 1 (needs 3 * C <= 16, cpad == 16): channel k*C + c of an output voxel = channel c of its w-neighbour k-1 in
 * the copy's orientation, zero outside the tile — the input layout of a first conv planned with kw_taps = 1.
 * kwpack = 2 (needs 3 * C <= cpad, fp16): the fp16x3 split of the fp32 input, channels [hi (C) | hi (C) | lo (C) | 0]. */
int bsg_gather_patch_tta(const float* vol, int C, int Z, int Y, int X, int z0, int y0, int x0, int P0, int P1, int P2,
                         const int* mirror_codes_host, int nmirrors, void* out16, int cpad, int out_f16, int kwpack,
                         void* stream);

/* InstanceNorm3d / GroupNorm (generic_UNet.py:62-65,72) from the statistics the conv epilogue accumulated:
 * stats (fp64) [N][C][2] = (sum, sum of squares) over `count` voxels -> scale_shift (fp32) [N][C][2] with
 * y = x*scale + shift == (x-mean)*rsqrt(var+eps)*gamma + beta.  groups = 0: per channel; > 0: GroupNorm. */
int bsg_norm_finalize(const double* stats, int N, int C, int groups, double count, float eps, const float* gamma,
                      const float* beta, float* scale_shift, void* stream);
/* Same statistics -> rows [coff, coff+C) of a consumer-side table [N][ctot][4] = (scale, shift, slope, 0): the input
 * transform of the conv that consumes the raw tensor (bsg_conv_desc.in_norm).  Channels of the table that belong to an
 * un-normalised producer (the transposed-conv half of a concat buffer) are preset by the caller to (1, 0, 1, 0). */
int bsg_norm_finalize_table(const double* stats, int N, int C, int groups, double count, float eps, const float* gamma,
                            const float* beta, float slope, float* table, int ctot, int coff, void* stream);
/* In place on channels [coff, coff+C) of a (N, voxels, ctot) 16-bit buffer: x <- LeakyReLU(x*scale + shift), read as
 * fp16 when in_f16 = 1 (the conv stored its raw output as fp16, bsg_conv_desc.out_f16) else bf16, written back as fp16
 * when out_f16 = 1 else bf16. */
int bsg_norm_apply_lrelu(void* x, size_t voxels_per_item, int N, int C, int ctot, int coff,
                         const float* scale_shift, float slope, int in_f16, int out_f16, void* stream);

/* bsg_norm_apply_lrelu for the fp16x3 split layout (bsg_conv_desc.out_split_stride): reads hi (block 0) + lo (block 2) of
 * the C-channel tensor at [coff, coff + 3C), normalises in fp32 and writes the three blocks [hi | hi | lo] back. */
int bsg_norm_apply_lrelu_split(void* x, size_t voxels_per_item, int N, int C, int ctot, int coff, const float* scale_shift,
                               float slope, void* stream);

/* Fused tail of one tile: 1x1x1 segmentation head (generic_UNet.py:389-391, weights [ncls][cfeat] + optional bias,
 * HOST pointers), inference_apply_nonlin (0 sigmoid / 1 softmax / 2 identity), un-flip of each mirror's prediction,
 * result += mirror_weight * pred (mirror_weight = 1/num_results of the whole TTA), result *= gaussian,
 * aggregated_results[:, tile] += result.
 * feat: bf16 (fp16 when feat_f16 = 1) (nmirrors, P0, P1, P2, ctot) with the cfeat head inputs in channels [0,cfeat) (cfeat % 8 == 0, <= 64);
 * acc: fp32 [ncls][Z][Y][X]; gauss: fp32 [P0][P1][P2] or NULL.
 * norm_scale_shift (device fp32 [nmirrors][cfeat][2], or NULL): when given, `feat` holds the RAW output of the last conv
 * block and the head applies its deferred norm + LeakyReLU(norm_slope) on the fly (generic_UNet.py:72 of that block),
 * saving the block's separate bsg_norm_apply_lrelu pass. */
int bsg_head_tta_accumulate(const void* feat16, int feat_f16, int cfeat, int ctot, int P0, int P1, int P2,
                            const int* mirror_codes_host, int nmirrors, float mirror_weight,
                            const float* head_w_host, const float* head_b_host, int ncls, int nonlin,
                            const float* gauss, float* acc, int Z, int Y, int X, int z0, int y0, int x0,
                            const float* norm_scale_shift, float norm_slope, void* stream);
/* ... with the features in the fp16x3 split layout: head input = channels [0, cfeat) + channels [2*cfeat, 3*cfeat). */
int bsg_head_tta_accumulate_split(const void* feat16, int cfeat, int ctot, int P0, int P1, int P2,
                                  const int* mirror_codes_host, int nmirrors, float mirror_weight,
                                  const float* head_w_host, const float* head_b_host, int ncls, int nonlin,
                                  const float* gauss, float* acc, int Z, int Y, int X, int z0, int y0, int x0, void* stream);

/* class_probabilities = aggregated_results / aggregated_nb_of_predictions; mean over K accumulators (np.mean over
 * folds, run_brats2021_inference_singlethread.py:128); decision: mode 0 argmax(0) (main_files/run_inference.py:150),
 * mode 1 ordered threshold `for i,c in enumerate(order): seg[p[i] > 0.5] = c`
 * (save_segmentation_nifti_from_softmax with region_class_order, run_brats...py:144-156).
 * acc_list_host: K device pointers (host array); wsum: fp32 [nvox]; probs (fp32 [ncls][nvox]) and seg (u8 [nvox]) may
 * each be NULL. */
int bsg_finalize(const float* const* acc_list_host, int K, const float* wsum, int ncls, size_t nvox, int mode,
                 const int* order_host, float* probs, uint8_t* seg, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Multi-GPU: the (tile, mirror) work items of ONE case sharded over R GPUs (BASELINE configs[2]; the per-model predict
 * calls run_brats2021_inference_singlethread.py:97-106,113-124 are what is split).  Every rank accumulates its share
 * into a private fp32 accumulator [ncls][nvox]; the exchange step is one of:
 * ------------------------------------------------------------------------------------------------------------ */

/* Peer-memory route — reduce over ranks + bsg_finalize in ONE kernel.  The calling rank owns the voxel range
 * [v0, v0 + nv): it reads that range of all K*R accumulators (acc_table_dev: DEVICE array [K][R] of device pointers,
 * R-1 of every R of them peer memory mapped over NVLink), sums over ranks in rank order, divides by wsum (local,
 * geometry-only), averages the K folds, decides (mode / order_host as bsg_finalize) and stores the uint8 labels of the
 * range into each of the nseg label volumes of seg_table_dev (DEVICE array of device pointers: every rank's label
 * volume, so that all ranks end up with the whole volume, or just the local one).  nvox, v0, nv multiples of 4.
 * The caller orders the ranks around the call (a stream-ordered barrier before: all accumulators complete; after: all
 * slabs written) — e.g. two tiny NCCL all-reduces. */
int bsg_finalize_peer(const float* const* acc_table_dev, int K, int R, const float* wsum, int ncls, size_t nvox, size_t v0,
                      size_t nv, int mode, const int* order_host, uint8_t* const* seg_table_dev, int nseg, void* stream);
/* bsg_finalize_peer with the rank ordering INSIDE the kernel (no collective around the launch): flag_table_dev is a DEVICE
 * array [R] of pointers to every rank's flag block (2R + 2 uint32, zero-initialised once, in peer-mapped memory) as mapped
 * into this rank's address space; `epoch` is the number of this call (1, 2, ... — the same on every rank, which call in
 * lockstep).  The kernel announces "rank's accumulators complete" to every rank (release store at system scope), waits
 * until all ranks announced the epoch, reduces + finalizes + stores its slab into every label volume, announces "slab
 * landed, done with your accumulators" and waits for the same word from all ranks: once the launch has completed on the
 * stream, the local label volumes are whole and the local accumulators may be reused.  Every rank needs nv > 0.
 * The waits are bounded (10 s): a rank that never shows up makes the kernel give up and write a non-zero code into word
 * 2R + 1 of the local flag block (1 + r: rank r never announced; 0x100 + r: rank r's slab never landed) — the caller reads
 * that word after synchronising; the label volumes are then invalid. */
int bsg_finalize_peer_signal(const float* const* acc_table_dev, int K, int R, const float* wsum, int ncls, size_t nvox,
                             size_t v0, size_t nv, int mode, const int* order_host, uint8_t* const* seg_table_dev, int nseg,
                             uint32_t* const* flag_table_dev, int rank, uint32_t epoch, void* stream);
/* cudaDeviceEnablePeerAccess(current -> peer_device), idempotent. */
int bsg_enable_peer_access(int peer_device);
/* CUDA IPC plumbing for the peer route (ranks are processes).  bsg_ipc_export: the 64-byte handle of the ALLOCATION that
 * contains the device pointer `ptr` and ptr's offset inside it (a caching allocator may sub-allocate).  bsg_ipc_open: maps
 * such an allocation into the CURRENT device's address space (lazy peer access) and returns its base — peer pointer =
 * base + offset; open each handle once per process.  bsg_ipc_close unmaps it. */
int bsg_ipc_export(const void* ptr, void* handle64_host, size_t* offset_out);
int bsg_ipc_open(const void* handle64_host, void** base_out);
int bsg_ipc_close(void* base);

/* NCCL route — `ncclAllReduce(sum)` (root < 0) or `ncclReduce(sum)` to `root` of the fp32 accumulator in place, over
 * a communicator the library owns: rank 0 calls bsg_nccl_unique_id (128 bytes, host), ships the id to the other ranks
 * by any means (torch.distributed broadcast), every rank calls bsg_nccl_comm_create.  libnccl.so.2 is resolved with
 * dlopen at first use (inside a PyTorch process: the copy torch loaded). */
int bsg_nccl_unique_id(void* id128_host);
int bsg_nccl_comm_create(const void* id128_host, int nranks, int rank, void** comm_out);
int bsg_nccl_reduce_accumulator(void* comm, float* acc, size_t count, int root, void* stream);
int bsg_nccl_comm_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif /* BRAINSEG_B200_H */
