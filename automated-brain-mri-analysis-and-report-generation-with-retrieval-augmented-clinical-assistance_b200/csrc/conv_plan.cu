// Host planner for the tcgen05 conv kernel: picks the tile box / N tile / K chunk / pipeline depth for a layer,
// encodes the TMA tensor maps once, and launches.  Exposed through the C ABI as bsg_conv_plan_*.
#include <cudaTypedefs.h>
#include <string.h>
#include <new>
#include "bsg_common.cuh"
#include "conv_brick.cuh"
#include "conv_tc.cuh"

namespace bsg {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

namespace {

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, int cc, bool oob_nan = false) {
    auto fn = get_encode_fn();
    if (fn == nullptr) return set_error(BSG_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    uint32_t estr[5] = {1, 1, 1, 1, 1};
    CUtensorMapSwizzle sw = (cc == 64)   ? CU_TENSOR_MAP_SWIZZLE_128B
                            : (cc == 32) ? CU_TENSOR_MAP_SWIZZLE_64B
                                         : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                    reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_bytes),
                    reinterpret_cast<const cuuint32_t*>(box), reinterpret_cast<const cuuint32_t*>(estr),
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    oob_nan ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(BSG_ECUDA,
                         "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu box %u,%u,%u cc %d",
                         static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                         (unsigned long long)dims[2], box[0], box[1], box[2], cc);
    return BSG_OK;
}

uint32_t round_up(uint32_t x, uint32_t m) { return (x + m - 1) / m * m; }

}  // namespace
}  // namespace bsg

struct bsg_conv_plan {
    bsg::ConvArgs args;    // tile kernel (conv_tc.cu)
    bsg::BrickArgs bargs;  // brick kernel (conv_brick.cu), used when brick != 0
    int brick, brick_cc, brick_nt;
    int grid;
    size_t smem_bytes;
    double flops;
};

namespace bsg {
namespace {

// Brick kernel eligibility + geometry (see conv_brick.cuh).  Returns 1 when the plan was filled, 0 when the layer
// does not suit the brick kernel, < 0 on error.
int plan_brick(const bsg_conv_desc* d, bsg_conv_plan* p) {
    if (d->kind != BSG_CONV_K3 || d->stride != 1 || d->algo == 0 || d->out_split_stride != 0) return 0;
    const int cout_pad = static_cast<int>(round_up(d->cout, 32));
    if (cout_pad != 32 && cout_pad != 64) return 0;
    if (d->in_norm != nullptr && d->cin % 32 != 0) return 0;  // the in-consumer transform needs K chunks of >= 32 channels
    const int P = 256 / cout_pad;
    if (d->W % 8 != 0 || d->H % 16 != 0 || d->D % P != 0) return 0;
    BrickArgs& a = p->bargs;
    memset(&a, 0, sizeof(a));
    int cc = (d->cin % 64 == 0) ? 64 : (d->cin % 32 == 0 ? 32 : 16);
    // in-consumer norm: 32-channel chunks (half the stage, twice as many in flight next to the same slabs) were measured
    // SLOWER than 64-channel ones (64->32 @128^3 x 4: 1.19 vs 0.97 ms) — twice the MMA instructions per byte; kept as a
    // measurement switch only
    if (d->in_norm != nullptr && cc == 64 && d->in_norm_cc == 32 && d->cin / 32 * 3 <= 6) cc = 32;
    a.P = P;
    a.D = d->D;
    a.tw = d->W / 8;
    a.th = d->H / 16;
    a.tb = d->D / P;
    a.tn = d->N;
    a.nchunks = d->cin / cc;
    a.kwn = d->kw_taps == 1 ? 1 : 3;
    a.nphases = a.kwn * a.nchunks;
    a.slab_bytes = 9u * cout_pad * cc * 2u;
    // leave ~8 KB of the SM's shared memory unclaimed: the HBM-bound elementwise kernels of the other stream lane
    // (norm apply, gather, head) need their 1 KB system reservation each to become co-resident with this CTA
    const uint32_t avail = 227 * 1024 - 1024 - 1280 - 8192;
    const uint32_t stage_kwf = round_up(10u * 18u * cc * 2u, 1024), stage_3x = round_up(8u * 18u * cc * 2u, 1024);
    if (a.kwn == 3 && a.nphases <= 6 && static_cast<uint64_t>(a.nphases) * a.slab_bytes + 3ull * stage_kwf <= avail) {
        a.nslabbuf = a.nphases;  // resident slabs -> kw-fused activation boxes
        a.kwf = 1;
    } else if (a.kwn == 1 && a.nphases <= 6 && static_cast<uint64_t>(a.nphases) * a.slab_bytes + 3ull * stage_3x <= avail) {
        a.nslabbuf = a.nphases;  // 3x3x1 kernel: resident slabs, plain 8-wide boxes (no w halo to share)
        a.kwf = 0;
    } else if (2ull * a.slab_bytes + 2ull * stage_3x <= avail) {
        a.nslabbuf = 2;
        a.kwf = 0;
    } else {
        return 0;
    }
    // the in-consumer norm transform rewrites every landed box once: with streamed slabs an element arrives in three
    // kw-shifted boxes and the 128 B/clk shared-memory port, which MMA operand reads already fill, pays for it three
    // times (measured: 64->64 @128^3 1.71 -> 2.73 ms, more than the 0.37 ms pass it replaces) -> kw-fused plans only
    if (d->in_norm != nullptr && !a.kwf) return 0;
    const uint32_t box_w = a.kwf ? 10u : 8u;
    a.a_tx_bytes = box_w * 18u * cc * 2u;
    a.a_stage_bytes = a.kwf ? stage_kwf : stage_3x;
    a.nstages = static_cast<int>((avail - static_cast<uint32_t>(a.nslabbuf) * a.slab_bytes) / a.a_stage_bytes);
    if (a.nstages > 12) a.nstages = 12;

    const uint64_t ct = static_cast<uint64_t>(d->in_ctot);
    uint64_t dims[5] = {static_cast<uint64_t>(d->cin), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                        static_cast<uint64_t>(d->D), static_cast<uint64_t>(d->N)};
    uint64_t str[4] = {ct * 2, ct * 2 * d->W, ct * 2 * d->W * d->H, ct * 2 * d->W * d->H * d->D};
    uint32_t box[5] = {static_cast<uint32_t>(cc), box_w, 18u, 1u, 1u};
    // XF: padding arrives as NaN and is zeroed by the transform (after the norm, as the reference pads)
    int rc = encode_map(&a.mapA, d->in, 5, dims, str, box, cc, d->in_norm != nullptr);
    if (rc != BSG_OK) return rc;
    // weights [27 taps (kd, kw, kh)][cout_pad][cin] seen as (cin, row, kh, kw, kd): a box of the 3 kd taps of one
    // (kh, kw) lands as [kd][cout_pad][cc] = the B operand of one N = 3*cout_pad MMA
    const uint64_t tap_bytes = static_cast<uint64_t>(d->cin) * 2 * cout_pad;
    const uint64_t kwn = static_cast<uint64_t>(a.kwn);
    uint64_t wdims[5] = {static_cast<uint64_t>(d->cin), static_cast<uint64_t>(cout_pad), 3ull, kwn, 3ull};
    uint64_t wstr[4] = {static_cast<uint64_t>(d->cin) * 2, tap_bytes, tap_bytes * 3, tap_bytes * 3 * kwn};
    uint32_t wbox[5] = {static_cast<uint32_t>(cc), static_cast<uint32_t>(cout_pad), 1u, 1u, 3u};
    rc = encode_map(&a.mapW, d->weights, 5, wdims, wstr, wbox, cc);
    if (rc != BSG_OK) return rc;

    a.out = static_cast<__nv_bfloat16*>(d->out);
    a.os_w = d->out_ctot;
    a.os_h = static_cast<long long>(d->out_ctot) * d->W;
    a.os_d = a.os_h * d->H;
    a.os_n = a.os_d * d->D;
    a.out_c_off = d->out_coff;
    a.cout = d->cout;
    a.cout_pad = cout_pad;
    a.bias = d->bias;
    a.slope = d->slope;
    a.act = d->act;
    a.stats = d->stats;
    a.out_f16 = d->out_f16;
    a.in_f16 = d->in_f16;
    a.overflow = d->overflow;
    a.in_norm = d->in_norm;
    a.in_norm_c = d->in_norm_c;

    p->brick = 1;
    p->brick_cc = cc;
    p->brick_nt = cout_pad;
    const int units = a.tn * a.th * a.tw * a.tb;
    const int max_ctas = d->max_ctas > 0 ? d->max_ctas : sm_count_cached();
    p->grid = units < max_ctas ? units : max_ctas;
    p->smem_bytes = conv_brick_smem_bytes(a);
    p->flops = 2.0 * 9 * a.kwn * static_cast<double>(d->cin) * d->cout * (static_cast<double>(d->W) * d->H * d->D * d->N);
    return 1;
}

}  // namespace
}  // namespace bsg

using namespace bsg;

extern "C" {

int bsg_version(void) { return 100; }

size_t bsg_last_error(char* buf, size_t cap) {
    size_t n = strlen(g_err);
    if (buf != nullptr && cap > 0) {
        size_t m = n < cap - 1 ? n : cap - 1;
        memcpy(buf, g_err, m);
        buf[m] = 0;
    }
    return n;
}

int bsg_check_device(void) {
    int dev = 0, major = 0;
    BSG_CUDA_OK(cudaGetDevice(&dev));
    BSG_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return set_error(BSG_EARCH, "device compute capability %d.x is not sm_100", major);
    return BSG_OK;
}

int bsg_sm_count(void) { return sm_count_cached(); }

size_t bsg_conv_desc_size(void) { return sizeof(bsg_conv_desc); }

int bsg_conv_plan_create(const bsg_conv_desc* d, bsg_conv_plan** out_plan) {
    BSG_REQUIRE(d != nullptr && out_plan != nullptr, "null argument");
    BSG_REQUIRE(d->kind == BSG_CONV_K3 || d->kind == BSG_CONVT_K2S2 || d->kind == BSG_CONV_K1, "bad kind %d", d->kind);
    BSG_REQUIRE(d->cin > 0 && d->cin % 16 == 0, "cin %d must be a positive multiple of 16", d->cin);
    BSG_REQUIRE(d->in_ctot % 8 == 0 && d->out_ctot % 8 == 0 && d->out_coff % 8 == 0,
                "channel strides/offsets must be multiples of 8 (16-byte alignment)");
    BSG_REQUIRE(d->N > 0 && d->D > 0 && d->H > 0 && d->W > 0 && d->cout > 0, "empty tensor");
    BSG_REQUIRE((reinterpret_cast<uintptr_t>(d->in) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(d->weights) & 15) == 0,
                "pointers must be 16-byte aligned");
    const int stride = (d->kind == BSG_CONV_K3) ? d->stride : 1;
    BSG_REQUIRE(stride == 1 || stride == 2, "stride %d", stride);
    if (stride == 2) BSG_REQUIRE(d->D % 2 == 0 && d->H % 2 == 0 && d->W % 2 == 0, "stride 2 needs even extents");

    bsg_conv_plan* p = new (std::nothrow) bsg_conv_plan();
    if (p == nullptr) return set_error(BSG_ENOMEM, "host allocation failed");
    p->brick = 0;
    {
        const int br = plan_brick(d, p);
        if (br < 0) {
            delete p;
            return br;
        }
        if (br == 1) {
            *out_plan = p;
            return BSG_OK;
        }
    }
    if (d->kw_taps == 1) {
        delete p;
        return set_error(BSG_EINVAL, "kw_taps 1 (3x3x1 kernel) exists for the brick kernel only (stride-1, Cout <= 64, "
                                     "W %% 8 == 0, H %% 16 == 0, D %% (256 / Cout_pad) == 0)");
    }
    if (d->in_norm != nullptr) {
        delete p;
        return set_error(BSG_EINVAL, "in_norm: the in-consumer norm transform exists for the brick kernel only (stride-1 k3, "
                                     "Cout <= 64, Cin %% 32 == 0, W %% 8 == 0, H %% 16 == 0)");
    }
    ConvArgs& a = p->args;
    memset(&a, 0, sizeof(a));

    // tile coordinate space = output voxels (K3/K1) or input voxels (transposed conv)
    a.Wo = d->W / stride;
    a.Ho = d->H / stride;
    a.Do = d->D / stride;
    a.No = d->N;
    a.stride = stride;
    a.ntaps = (d->kind == BSG_CONV_K3) ? 27 : 1;
    a.cc = (d->cin % 64 == 0) ? 64 : (d->cin % 32 == 0 ? 32 : 16);
    a.nchunks = d->cin / a.cc;

    // tile box: 128 voxels, w fastest
    auto pow2_le = [](int x) {
        int r = 1;
        while (r * 2 <= x) r *= 2;
        return r;
    };
    a.bw = a.Wo >= 8 ? 8 : pow2_le(a.Wo);
    int rem = 128 / a.bw;
    a.bh = a.Ho >= rem ? rem : pow2_le(a.Ho);
    rem /= a.bh;
    a.bd = a.Do >= rem ? rem : pow2_le(a.Do);
    rem /= a.bd;
    a.bn = rem;
    a.tw = ceil_div(a.Wo, a.bw);
    a.th = ceil_div(a.Ho, a.bh);
    a.td = ceil_div(a.Do, a.bd);
    a.tn = ceil_div(a.No, a.bn);

    // N tiling
    a.cout = d->cout;
    a.cout_pad = static_cast<int>(round_up(d->cout, 32));
    if (a.cout_pad > 512) {
        delete p;
        return set_error(BSG_EINVAL, "cout %d > 512 not supported", d->cout);
    }
    int split = 1;
    while (a.cout_pad / split > 256 || a.cout_pad % split != 0 || (a.cout_pad / split) % 32 != 0) {
        ++split;
        if (split > 64) {
            delete p;
            return set_error(BSG_EINVAL, "cannot tile cout %d", d->cout);
        }
    }
    if (d->kind != BSG_CONVT_K2S2) {
        // small levels (<= 8^3 voxels per item): a handful of 128-voxel tiles cannot fill 148 SMs, and each CTA then walks
        // the whole K = 27 * Cin loop at N = 256.  Narrower N tiles (>= 64 columns, still 67 % tensor-efficient) multiply
        // the CTA count and divide every CTA's MMA time; the activation tile they re-read is tiny.
        const long long mtiles = static_cast<long long>(a.tw) * a.th * a.td * a.tn;
        const int sms = sm_count_cached();
        while (mtiles * split < sms) {
            int next = split + 1;
            while (next <= a.cout_pad / 32 && (a.cout_pad % next != 0 || (a.cout_pad / next) % 32 != 0)) ++next;
            if (next > a.cout_pad / 32 || a.cout_pad / next < 64 || mtiles * next > sms) break;  // stay within one wave
            split = next;
        }
    }
    a.ntile = a.cout_pad / split;
    a.out_mul = (d->kind == BSG_CONVT_K2S2) ? 2 : 1;
    a.n_ntiles = split * (a.out_mul == 2 ? 8 : 1);
    if (a.out_mul == 2 && split == 1) {
        // transposed conv: the 8 output parities are extra GEMM columns; let one N tile span as many whole parities
        // as fit 256 columns so the activation tile is read once instead of once per parity
        int m = 1;
        while (m < 8 && a.cout_pad * m * 2 <= 256) m *= 2;
        a.ntile = a.cout_pad * m;
        a.n_ntiles = 8 / m;
    }
    // M blocking: a work item = two M tiles (planes d0, d0 + 1 of one activation box) against ONE weight stage — two
    // accumulators x two buffers fill the 512 TMEM columns.  Halves the weight bytes per MMA cycle through L2 and shared
    // memory.  Evidence (profiles/r02_conv_tile_stride2_full.md): the stride-2 convs with >= 64 input channels run one-tap
    // stages of 16 KB activations + 16 KB weights per 256 MMA cycles; six of them in flight against ~2500 cycles of TMA
    // latency under load supply 75 B/clk where full rate needs 128 — tensor pipe 53-55 %.
    a.mb = 1;
    {
        const long long mtiles = static_cast<long long>(a.tw) * a.th * a.td * a.tn;
        const bool can = d->kind == BSG_CONV_K3 && a.bw == 8 && a.bd == 1 && a.bn == 1 && a.ntile <= 128 && a.Do % 2 == 0 &&
                         d->out_split_stride == 0 && d->tma_store != 1;  // direct 16-bit epilogue only
        // Measured per 4 forwards (gpurun_out/r02_layers24_*): stride 2 32->64 @128^3 0.250 -> 0.214 ms, 64->128 @128^3
        // 0.577 -> 0.533; stride 1 (haloed kh box, two planes per stage, against the 2-CTA weight multicast it replaces)
        // 128->128 @64^3 0.766 -> 0.717, 256->128 @64^3 1.369 -> 1.302 — but 0.087 -> 0.091 and 0.158 -> 0.165 on the
        // @32^3 layers, whose 1024 tiles no longer balance over 148 CTAs in pairs: stride 1 only from 16 waves on.
        const long long items = mtiles * a.n_ntiles;
        const bool want = d->mblock == 1 || (d->mblock <= 0 && items >= (stride == 2 ? 4ll : 16ll) * sm_count_cached());
        if (can && want) {
            a.mb = 2;
            a.td = ceil_div(a.Do, 2);
        }
    }
    a.tmem_cols = 32;
    while (a.tmem_cols < static_cast<uint32_t>(2 * a.ntile * a.mb)) a.tmem_cols *= 2;

    // pair mode: 2-CTA clusters share every weight stage (each CTA fetches half of the N rows and multicasts them), which
    // halves the L2 -> SM weight traffic of the layers that re-read their weights once per 128-voxel tile
    a.pair = 0;
    {
        const long long mtiles = static_cast<long long>(a.tw) * a.th * a.td * a.tn;
        // measured (gpurun_out/bringup16.log): +3 % on the stride-1 layers with N >= 128, nothing on the stride-2 layers
        // (they are bound by the per-stage issue overhead of their 1-tap stages, not by weight traffic)
        const bool wanted = d->pair == 1 || (d->pair != 0 && a.ntile >= 128 && stride == 1 && a.mb == 1);
        if (wanted && d->kind != BSG_CONVT_K2S2 && a.ntile % 16 == 0 && mtiles % 2 == 0 &&
            (d->pair == 1 || mtiles * a.n_ntiles >= 2ll * sm_count_cached()))
            a.pair = 1;
    }
    // Epilogue through shared memory + TMA tensor stores.  A transposed conv's output voxels are two apart, so the direct
    // epilogue's per-thread rows are 32 separate 32-byte writes per store instruction; staged, a warp's 32 voxels leave as
    // one tensor store of 64-channel rows = whole 128-byte lines.  Measured per 4 forwards: 64->64 @64^3 0.348 -> 0.321 ms,
    // 128->128 @32^3 0.107 -> 0.094, 256->256 @16^3 0.043 -> 0.038; with 32-channel (64-byte, half-line) rows the staged
    // route LOSES (64->32 @64^3: 0.160 -> 0.184 ms), as it does on the stride-1 / stride-2 convs whose direct rows are
    // already contiguous (33 KB of staging taken from their pipeline: -7 % on model 2's forward).  So: transposed convs
    // with >= 64-channel rows only; tma_store = 1 forces it anywhere, 2 forbids it.
    a.store_cols = (a.cout_pad % 64 == 0 && a.ntile % 64 == 0) ? 64 : 32;
    a.tma_out = (d->tma_store == 1 || (d->tma_store <= 0 && d->kind == BSG_CONVT_K2S2 && a.store_cols == 64)) &&
                d->out_split_stride == 0 && d->cout % 8 == 0;
    // kh halo reuse: needs the canonical 8 x 16 x 1 x 1 box, stride 1, 27 taps and >= 3 pipeline stages
    const uint32_t budget = 227 * 1024 - 4096 - 6144 - (a.tma_out ? kTmaOutSmemBytes : 0);  // barriers + bias + alignment slack + room for co-resident CTAs
    auto stage_bytes = [&](int khs, int taps3, uint32_t* ab, uint32_t* bb) {
        const uint32_t rows = (khs ? static_cast<uint32_t>((a.bh + 2) * 8) : 128u) * static_cast<uint32_t>(a.mb);
        *ab = round_up(rows * a.cc * 2, 1024) * (taps3 ? 3 : 1);
        *bb = round_up(static_cast<uint32_t>(a.ntile) * a.cc * 2 * ((khs || taps3) ? 3 : 1), 1024);
        return *ab + *bb;
    };
    int khs = 0;
    if (a.ntaps == 27 && stride == 1 && a.bw == 8 && a.bd == 1 && a.bn == 1 && d->use_khshift != 0) {
        uint32_t ab, bb;
        const uint32_t sb = stage_bytes(1, 0, &ab, &bb);
        if (budget / sb >= 3 || d->use_khshift == 1 || (a.mb == 2 && budget / sb >= 2)) khs = 1;
        if (budget / sb < 2) khs = 0;
    }
    a.khshift = khs;
    // three kh taps per stage as three separate boxes (any box shape, either stride) where the haloed box is not
    // available: a third of the pipeline steps.  Only with >= 3 stages in flight, and never in pair mode.
    int taps3 = 0;
    if (!khs && !a.pair && a.ntaps == 27 && d->use_khshift != 0) {
        uint32_t ab, bb;
        if (budget / stage_bytes(0, 1, &ab, &bb) >= 3) taps3 = 1;
    }
    a.taps3 = taps3;
    const uint32_t sb = stage_bytes(khs, taps3, &a.a_stage_bytes, &a.b_stage_bytes);
    a.stage_tx_bytes = (khs ? static_cast<uint32_t>((a.bh + 2) * 8) : 128u) * a.mb * a.cc * 2 * (taps3 ? 3 : 1) +
                       static_cast<uint32_t>(a.ntile) * a.cc * 2 * ((khs || taps3) ? 3 : 1);
    a.nstages = static_cast<int>(budget / sb);
    if (a.nstages > 12) a.nstages = 12;
    if (a.nstages < 2) {
        delete p;
        return set_error(BSG_EINVAL, "layer does not fit shared memory (stage %u bytes)", sb);
    }

    // tensor maps
    const uint64_t ct = static_cast<uint64_t>(d->in_ctot);
    const __nv_bfloat16* in = static_cast<const __nv_bfloat16*>(d->in);
    int rc = BSG_OK;
    if (stride == 1) {
        uint64_t dims[5] = {static_cast<uint64_t>(d->cin), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                            static_cast<uint64_t>(d->D), static_cast<uint64_t>(d->N)};
        uint64_t str[4] = {ct * 2, ct * 2 * d->W, ct * 2 * d->W * d->H, ct * 2 * d->W * d->H * d->D};
        uint32_t box[5] = {static_cast<uint32_t>(a.cc), static_cast<uint32_t>(a.bw),
                           static_cast<uint32_t>(a.bh + (khs ? 2 : 0)), static_cast<uint32_t>(a.bd * a.mb),
                           static_cast<uint32_t>(a.bn)};
        rc = encode_map(&a.mapA[0], in, 5, dims, str, box, a.cc);
    } else {
        for (int par = 0; par < 8 && rc == BSG_OK; ++par) {
            const int pw = par & 1, ph = (par >> 1) & 1, pd = (par >> 2) & 1;
            const __nv_bfloat16* base =
                in + (static_cast<uint64_t>(pw) + static_cast<uint64_t>(ph) * d->W +
                      static_cast<uint64_t>(pd) * d->W * d->H) * ct;
            uint64_t dims[5] = {static_cast<uint64_t>(d->cin), static_cast<uint64_t>(d->W / 2),
                                static_cast<uint64_t>(d->H / 2), static_cast<uint64_t>(d->D / 2),
                                static_cast<uint64_t>(d->N)};
            uint64_t str[4] = {ct * 4, ct * 4 * d->W, ct * 4 * d->W * d->H, ct * 2 * d->W * d->H * d->D};
            uint32_t box[5] = {static_cast<uint32_t>(a.cc), static_cast<uint32_t>(a.bw), static_cast<uint32_t>(a.bh),
                               static_cast<uint32_t>(a.bd * a.mb), static_cast<uint32_t>(a.bn)};
            rc = encode_map(&a.mapA[par], base, 5, dims, str, box, a.cc);
        }
    }
    if (rc == BSG_OK) {
        const uint64_t rows = static_cast<uint64_t>(a.cout_pad) * (a.out_mul == 2 ? 8 : 1);
        uint64_t dims[3] = {static_cast<uint64_t>(d->cin), rows, static_cast<uint64_t>(a.ntaps)};
        uint64_t str[2] = {static_cast<uint64_t>(d->cin) * 2, static_cast<uint64_t>(d->cin) * 2 * rows};
        uint32_t box[3] = {static_cast<uint32_t>(a.cc), static_cast<uint32_t>(a.ntile), (khs || taps3) ? 3u : 1u};
        rc = encode_map(&a.mapW, d->weights, 3, dims, str, box, a.cc);
    }
    if (rc == BSG_OK && a.pair) {
        const uint64_t rows = static_cast<uint64_t>(a.cout_pad);
        uint64_t dims[3] = {static_cast<uint64_t>(d->cin), rows, static_cast<uint64_t>(a.ntaps)};
        uint64_t str[2] = {static_cast<uint64_t>(d->cin) * 2, static_cast<uint64_t>(d->cin) * 2 * rows};
        uint32_t box[3] = {static_cast<uint32_t>(a.cc), static_cast<uint32_t>(a.ntile / 2), 1u};
        rc = encode_map(&a.mapWh, d->weights, 3, dims, str, box, a.cc);
    }
    if (rc == BSG_OK && a.tma_out) {
        // one epilogue warp = 32 consecutive rows of the 128-row tile = the sub-box (bw, sh, sd, sn) of the tile box
        const int sh = a.bh < 32 / a.bw ? a.bh : 32 / a.bw;
        const int sd = a.bd < 32 / (a.bw * sh) ? a.bd : 32 / (a.bw * sh);
        const int sn = 32 / (a.bw * sh * sd);
        const int om = a.out_mul;
        const uint64_t oct = static_cast<uint64_t>(d->out_ctot);
        const uint64_t Wo_ = static_cast<uint64_t>(a.Wo) * om, Ho_ = static_cast<uint64_t>(a.Ho) * om, Do_ = static_cast<uint64_t>(a.Do) * om;
        uint64_t dims[5] = {static_cast<uint64_t>(d->cout), static_cast<uint64_t>(a.Wo), static_cast<uint64_t>(a.Ho),
                            static_cast<uint64_t>(a.Do), static_cast<uint64_t>(a.No)};
        uint64_t str[4] = {oct * 2 * om, oct * 2 * Wo_ * om, oct * 2 * Wo_ * Ho_ * om, oct * 2 * Wo_ * Ho_ * Do_};
        uint32_t box[5] = {static_cast<uint32_t>(a.store_cols), static_cast<uint32_t>(a.bw), static_cast<uint32_t>(sh),
                           static_cast<uint32_t>(sd), static_cast<uint32_t>(sn)};
        const __nv_bfloat16* obase = static_cast<const __nv_bfloat16*>(d->out) + d->out_coff;
        for (int par = 0; par < (om == 2 ? 8 : 1) && rc == BSG_OK; ++par) {
            const uint64_t pw = par & 1, ph = (par >> 1) & 1, pd = (par >> 2) & 1;
            rc = encode_map(&a.mapO[par], obase + ((pd * Ho_ + ph) * Wo_ + pw) * oct, 5, dims, str, box, a.store_cols);
        }
    }
    if (rc != BSG_OK) {
        delete p;
        return rc;
    }

    // epilogue
    const int Wout = a.Wo * a.out_mul, Hout = a.Ho * a.out_mul, Dout = a.Do * a.out_mul;
    a.out = static_cast<__nv_bfloat16*>(d->out);
    a.os_w = d->out_ctot;
    a.os_h = static_cast<long long>(d->out_ctot) * Wout;
    a.os_d = a.os_h * Hout;
    a.os_n = a.os_d * Dout;
    a.out_c_off = d->out_coff;
    a.bias = d->bias;
    a.slope = d->slope;
    a.act = d->act;
    a.stats = d->stats;
    a.out_f16 = d->out_f16;
    a.in_f16 = d->in_f16;
    a.overflow = d->overflow;
    a.split_stride = d->out_split_stride;
    if (a.split_stride != 0 && !(d->out_f16 && d->in_f16 && a.split_stride % 8 == 0)) {
        delete p;
        return set_error(BSG_EINVAL, "out_split_stride needs fp16 operands and a block stride that is a multiple of 8");
    }

    const int total_tiles = a.tn * a.td * a.th * a.tw * a.n_ntiles;
    int max_ctas = d->max_ctas > 0 ? d->max_ctas : sm_count_cached();
    p->grid = total_tiles < max_ctas ? total_tiles : max_ctas;
    if (a.pair) {
        const int clusters = (total_tiles / 2) < (max_ctas / 2) ? (total_tiles / 2) : (max_ctas / 2);
        p->grid = 2 * (clusters > 0 ? clusters : 1);
    }
    p->smem_bytes = conv_tc_smem_bytes(a);
    p->flops = 2.0 * a.ntaps * (a.out_mul == 2 ? 8 : 1) * static_cast<double>(d->cin) * d->cout *
               (static_cast<double>(a.Wo) * a.Ho * a.Do * a.No);
    *out_plan = p;
    return BSG_OK;
}

int bsg_conv_plan_run(const bsg_conv_plan* plan, void* stream) {
    BSG_REQUIRE(plan != nullptr, "null plan");
    if (plan->brick)
        BSG_CUDA_OK(launch_conv_brick(plan->bargs, plan->brick_cc, plan->brick_nt, plan->grid, plan->smem_bytes,
                                      static_cast<cudaStream_t>(stream)));
    else
        BSG_CUDA_OK(launch_conv_tc(plan->args, plan->grid, plan->smem_bytes, static_cast<cudaStream_t>(stream)));
    return BSG_OK;
}

void bsg_conv_plan_destroy(bsg_conv_plan* plan) { delete plan; }

int bsg_conv_plan_info(const bsg_conv_plan* plan, bsg_conv_info* info) {
    BSG_REQUIRE(plan != nullptr && info != nullptr, "null argument");
    if (plan->brick) {
        const BrickArgs& b = plan->bargs;
        info->bw = 8;
        info->bh = 16;
        info->bd = b.P;
        info->bn = 1;
        info->ntile = plan->brick_nt;
        info->n_ntiles = 1;
        info->cc = plan->brick_cc;
        info->nstages = b.nstages;
        info->khshift = 2 + b.nslabbuf + (b.kwf ? 10 : 0) + (b.in_norm ? 1000 : 0);  /* brick marker: 2 + slab buffers (+10: kw-fused, +1000: in-consumer norm) */
        info->grid = plan->grid;
        info->smem_bytes = plan->smem_bytes;
        info->flops = plan->flops;
        return BSG_OK;
    }
    const ConvArgs& a = plan->args;
    info->bw = a.bw;
    info->bh = a.bh;
    info->bd = a.bd;
    info->bn = a.bn;
    info->ntile = a.ntile;
    info->n_ntiles = a.n_ntiles;
    info->cc = a.cc;
    info->nstages = a.nstages;
    info->khshift = a.khshift + (a.taps3 ? 3 : 0) + (a.mb == 2 ? 10 : 0) + (a.pair ? 100 : 0);  /* 1: kh halo reuse, 3: three kh taps per stage, +10: M blocking, +100: 2-CTA pair mode */
    info->grid = plan->grid;
    info->smem_bytes = plan->smem_bytes;
    info->flops = plan->flops;
    return BSG_OK;
}

}  // extern "C"
