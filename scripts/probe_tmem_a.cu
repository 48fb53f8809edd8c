// Hardware probe (B200): tcgen05.mma issue cost per instruction for small N with both operands in shared memory (SS)
// versus the A operand staged into TMEM with tcgen05.cp and re-used by several MMAs (TS).  Decides whether the
// Cout = 32 conv layers (shared-memory-read bound in SS form) should stage their activation slabs through TMEM.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/probe_tmem_a scripts/probe_tmem_a.cu -lcuda
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../automated-brain-mri-analysis-and-report-generation-with-retrieval-augmented-clinical-assistance_b200/csrc/bsg_ptx.cuh"

using namespace bsg;

#define CK(x)                                                                                \
    do {                                                                                     \
        cudaError_t e_ = (x);                                                                \
        if (e_ != cudaSuccess) {                                                             \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
            exit(1);                                                                         \
        }                                                                                    \
    } while (0)

__device__ __forceinline__ void utccp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;\n" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

struct Params {
    CUtensorMap mapA, mapB;
    int mode;   // 0: SS, 1: TS (cp + MMA), 2: cp only, 3: SS with one shared accumulator
    int N;      // MMA N
    int nacc;   // accumulators an A slab is re-used for
    int iters;  // timing iterations (0 = correctness pass: one K=64 product into accumulator 0)
    float* out;          // [128][256] (block 0)
    long long* cycles;   // [gridDim.x]
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;               // 128 x 64 bf16, SW128: 16 KB
    uint8_t* sB = smem + 16 * 1024;   // 256 x 64 bf16, SW128: 32 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t tmemA = tmem + 448;  // 4 K-slabs x 8 columns
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar[0], 16 * 1024 + 32 * 1024);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                         smem_u32(sA)),
                     "l"(reinterpret_cast<uint64_t>(&p.mapA)), "r"(smem_u32(&bar[0])), "r"(0), "r"(0)
                     : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                         smem_u32(sB)),
                     "l"(reinterpret_cast<uint64_t>(&p.mapB)), "r"(smem_u32(&bar[0])), "r"(0), "r"(0)
                     : "memory");
        mbar_wait(&bar[0], 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, p.N);
        const uint64_t dbase = make_smem_desc(0, 1024, kLayoutSW128);
        const uint32_t a16 = smem_u32(sA) >> 4, b16 = smem_u32(sB) >> 4;
        const int iters = p.iters > 0 ? p.iters : 1;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int k = 0; k < 4; ++k) {
                const uint64_t ad = dbase | static_cast<uint64_t>(a16 + 2 * k);
                const uint64_t bd = dbase | static_cast<uint64_t>(b16 + 2 * k);
                const uint32_t accf = (p.iters > 0) ? 1u : (k != 0);
                if (p.mode == 1 || p.mode == 2) utccp_128x256b(tmemA + 8 * k, ad);
                if (p.mode == 2) continue;
                for (int j = 0; j < p.nacc; ++j) {
                    const uint32_t d = tmem + (p.mode == 3 ? 0 : j * p.N);
                    if (p.mode == 1)
                        umma_bf16_ts(d, tmemA + 8 * k, bd, idesc, accf);
                    else
                        umma_bf16(d, ad, bd, idesc, accf);
                }
            }
        }
        umma_commit(&bar[1]);
        mbar_wait(&bar[1], 0);
        const long long t1 = clock64();
        p.cycles[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    tc_fence_after();
    if (p.iters == 0 && blockIdx.x == 0) {
        for (int cb = 0; cb < p.N; cb += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + cb, v);
            tmem_ld_wait();
            for (int i = 0; i < 32; ++i) p.out[(warp * 32 + lane) * 256 + cb + i] = __uint_as_float(v[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}


// Straight-line issue: GROUP MMAs (or cp + MMAs) per iteration, fully unrolled, one thread, descriptors derived from
// two bases by compile-time offsets.  MODE 0: SS; 1: TS (one cp feeds NACC MMAs); 2: SS, commit + mbarrier wait after
// every group (exposes the drain / refill cost of the issue queue).
template <int N, int MODE, int NACC, int GROUP>
__global__ void __launch_bounds__(128, 1) unrolled_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + 16 * 1024;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t tmemA = tmem + 448;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar[0], 16 * 1024 + 32 * 1024);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                         smem_u32(sA)),
                     "l"(reinterpret_cast<uint64_t>(&p.mapA)), "r"(smem_u32(&bar[0])), "r"(0), "r"(0)
                     : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                         smem_u32(sB)),
                     "l"(reinterpret_cast<uint64_t>(&p.mapB)), "r"(smem_u32(&bar[0])), "r"(0), "r"(0)
                     : "memory");
        mbar_wait(&bar[0], 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint64_t dbase = make_smem_desc(0, 1024, kLayoutSW128);
        const uint64_t abase = dbase | static_cast<uint64_t>(smem_u32(sA) >> 4);
        const uint64_t bbase = dbase | static_cast<uint64_t>(smem_u32(sB) >> 4);
        uint32_t parity = 0;
        const long long t0 = clock64();
        for (int it = 0; it < p.iters; ++it) {
#pragma unroll
            for (int g = 0; g < GROUP; ++g) {
                const int k = g % 4;
                const uint64_t ad = abase + 2 * k;
                const uint64_t bd = bbase + 2 * k;
                if (MODE == 1) {
                    utccp_128x256b(tmemA + 8 * k, ad);
#pragma unroll
                    for (int j = 0; j < NACC; ++j) umma_bf16_ts(tmem + j * N, tmemA + 8 * k, bd, idesc, 1u);
                } else {
                    umma_bf16(tmem + (g % NACC) * N, ad, bd, idesc, 1u);
                }
            }
            if (MODE == 2) {
                umma_commit(&bar[1]);
                mbar_wait(&bar[1], parity);
                parity ^= 1u;
            }
        }
        if (MODE != 2) {
            umma_commit(&bar[1]);
            mbar_wait(&bar[1], 0);
        }
        const long long t1 = clock64();
        p.cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

template <int N, int MODE, int NACC, int GROUP>
static void run_unrolled(Params p, const char* name, long long* dcyc) {
    CK(cudaFuncSetAttribute(unrolled_kernel<N, MODE, NACC, GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int grid : {1, 148}) {
        p.iters = 300;
        unrolled_kernel<N, MODE, NACC, GROUP><<<grid, 128, 64 * 1024>>>(p);
        CK(cudaDeviceSynchronize());
        std::vector<long long> cyc(grid);
        CK(cudaMemcpy(cyc.data(), dcyc, grid * 8, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (long long v : cyc) mx = v > mx ? v : mx;
        const int mmas = MODE == 1 ? GROUP * NACC : GROUP;
        printf("unrolled grid %3d %-28s: %7.1f cycles per MMA (ideal tensor %5.1f)%s\n", grid, name,
               (double)mx / p.iters / mmas, N / 2.0, MODE == 1 ? " [+1 cp per NACC MMAs]" : "");
    }
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
}
static void make_map(CUtensorMap* m, void* base, int rows) {
    cuuint64_t dims[2] = {64, static_cast<cuuint64_t>(rows)};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(rows)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("encode failed %d\n", (int)r);
        exit(1);
    }
}

int main() {
    CK(cudaSetDevice(0));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    std::vector<__nv_bfloat16> ha(128 * 64), hb(256 * 64);
    std::vector<float> fa(128 * 64), fb(256 * 64);
    srand(1);
    for (size_t i = 0; i < ha.size(); ++i) {
        ha[i] = __float2bfloat16((rand() % 17 - 8) / 8.0f);
        fa[i] = __bfloat162float(ha[i]);
    }
    for (size_t i = 0; i < hb.size(); ++i) {
        hb[i] = __float2bfloat16((rand() % 13 - 6) / 4.0f);
        fb[i] = __bfloat162float(hb[i]);
    }
    __nv_bfloat16 *da, *db;
    float* dout;
    long long* dcyc;
    CK(cudaMalloc(&da, ha.size() * 2));
    CK(cudaMalloc(&db, hb.size() * 2));
    CK(cudaMalloc(&dout, 128 * 256 * 4));
    CK(cudaMalloc(&dcyc, 148 * 8));
    CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    Params p;
    make_map(&p.mapA, da, 128);
    make_map(&p.mapB, db, 256);
    p.out = dout;
    p.cycles = dcyc;
    std::vector<float> h(128 * 256);
    // ---- correctness: SS and TS against the host product
    for (int mode = 0; mode < 2; ++mode)
        for (int N : {32, 64}) {
            p.mode = mode;
            p.N = N;
            p.nacc = 1;
            p.iters = 0;
            CK(cudaMemset(dout, 0, 128 * 256 * 4));
            probe_kernel<<<1, 128, 64 * 1024>>>(p);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost));
            double maxerr = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int k = 0; k < 64; ++k) ref += (double)fa[m * 64 + k] * fb[n * 64 + k];
                    maxerr = fmax(maxerr, fabs(ref - h[m * 256 + n]));
                }
            printf("correctness %s N=%d: max err %g %s\n", mode ? "TS(cp)" : "SS", N, maxerr, maxerr < 1e-3 ? "ok" : "MISMATCH");
        }

    // ---- straight-line issue
    run_unrolled<16, 0, 1, 24>(p, "SS N=16 1acc x24", dcyc);
    run_unrolled<32, 0, 1, 24>(p, "SS N=32 1acc x24", dcyc);
    run_unrolled<32, 0, 3, 24>(p, "SS N=32 3acc x24", dcyc);
    run_unrolled<64, 0, 1, 24>(p, "SS N=64 1acc x24", dcyc);
    run_unrolled<96, 0, 1, 24>(p, "SS N=96 1acc x24", dcyc);
    run_unrolled<128, 0, 1, 24>(p, "SS N=128 1acc x24", dcyc);
    run_unrolled<256, 0, 1, 24>(p, "SS N=256 1acc x24", dcyc);
    run_unrolled<32, 1, 3, 8>(p, "TS N=32 cp+3 x8", dcyc);
    run_unrolled<32, 1, 6, 4>(p, "TS N=32 cp+6 x4", dcyc);
    run_unrolled<32, 1, 1, 24>(p, "TS N=32 cp+1 x24", dcyc);
    run_unrolled<64, 1, 3, 8>(p, "TS N=64 cp+3 x8", dcyc);
    run_unrolled<128, 1, 3, 8>(p, "TS N=128 cp+3 x8", dcyc);
    run_unrolled<32, 2, 1, 6>(p, "SS N=32 commit+wait every 6", dcyc);
    run_unrolled<32, 2, 1, 18>(p, "SS N=32 commit+wait every 18", dcyc);
    run_unrolled<64, 2, 1, 12>(p, "SS N=64 commit+wait every 12", dcyc);
    run_unrolled<256, 2, 1, 4>(p, "SS N=256 commit+wait every 4", dcyc);
    // ---- timing (rolled loop)
    struct Cfg {
        const char* name;
        int mode, N, nacc;
    } cfgs[] = {
        {"SS N=32 x3acc", 0, 32, 3},   {"SS N=32 1acc", 3, 32, 3},   {"SS N=64 x3acc", 0, 64, 3},
        {"SS N=128 x3acc", 0, 128, 3}, {"SS N=256 x1acc", 0, 256, 1}, {"SS N=16 x3acc", 0, 16, 3},
        {"TS N=32 x3acc", 1, 32, 3},   {"TS N=32 x6acc", 1, 32, 6},  {"TS N=64 x3acc", 1, 64, 3},
        {"TS N=32 x1acc", 1, 32, 1},   {"cp only", 2, 32, 1},        {"SS N=96 x3acc", 0, 96, 3},
        {"TS N=96 x3acc", 1, 96, 3},
    };
    for (int grid : {1, 148})
        for (const Cfg& c : cfgs) {
            p.mode = c.mode;
            p.N = c.N;
            p.nacc = c.nacc;
            p.iters = 400;
            probe_kernel<<<grid, 128, 64 * 1024>>>(p);
            CK(cudaDeviceSynchronize());
            std::vector<long long> cyc(grid);
            CK(cudaMemcpy(cyc.data(), dcyc, grid * 8, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (long long v : cyc) mx = v > mx ? v : mx;
            const double per_iter = (double)mx / p.iters;
            const int mmas = c.mode == 2 ? 4 : 4 * c.nacc;
            printf("grid %3d  %-16s: %8.1f cycles per K=64 step-group, %6.1f per %s, ideal tensor %5.1f\n", grid, c.name, per_iter,
                   per_iter / mmas, c.mode == 2 ? "cp" : "MMA", c.N / 2.0);
        }
    return 0;
}
