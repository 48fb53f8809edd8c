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
    float* stats;        /* fp32 [N][cout][2] += (sum, sum of squares) of the pre-activation output, or NULL;
                            feeds InstanceNorm / GroupNorm (generic_UNet.py:62-65) */
    int use_khshift;     /* -1 auto, 0 off, 1 on: halo reuse of the h taps inside shared memory */
    int max_ctas;        /* 0 = one CTA per SM */
} bsg_conv_desc;

typedef struct bsg_conv_plan bsg_conv_plan;

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

#ifdef __cplusplus
}
#endif
#endif /* BRAINSEG_B200_H */
