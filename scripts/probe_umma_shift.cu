// Hardware probe (B200): can a tcgen05 shared-memory matrix descriptor start at a row that is NOT aligned to the
// swizzle atom (8 rows)?  If yes, one haloed activation tile in shared memory can serve every (kd, kh, kw) tap of a
// 3x3x3 conv through descriptor offsets alone.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/probe_umma_shift scripts/probe_umma_shift.cu -lcuda
//   run  : build/probe_umma_shift
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../automated-brain-mri-analysis-and-report-generation-with-retrieval-augmented-clinical-assistance_b200/csrc/bsg_ptx.cuh"

using namespace bsg;

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

struct Params {
    CUtensorMap mapA, mapB;
    int rows;        // rows of A loaded into smem
    int row_bytes;   // 128 (SW128) / 64 (SW64) / 32 (SW32)
    int shift;       // A start row
    int base_off;    // descriptor base-offset field
    int sbo_rows;    // rows between consecutive 8-row groups
    int ksteps;      // row_bytes / 32
    int layout;      // descriptor layout code
    float* out;      // [128][64]
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + 64 * 1024;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(slot, 64);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        const uint32_t nk = p.row_bytes / 2;
        mbar_expect_tx(&bar[0], p.rows * p.row_bytes + 64 * p.row_bytes);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
                     ::"r"(smem_u32(sA)), "l"(reinterpret_cast<uint64_t>(&p.mapA)), "r"(smem_u32(&bar[0])), "r"(0), "r"(0)
                     : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
                     ::"r"(smem_u32(sB)), "l"(reinterpret_cast<uint64_t>(&p.mapB)), "r"(smem_u32(&bar[0])), "r"(0), "r"(0)
                     : "memory");
        mbar_wait(&bar[0], 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, 64);
        for (int k = 0; k < p.ksteps; ++k) {
            uint64_t ad = make_smem_desc(smem_u32(sA) + p.shift * p.row_bytes + k * 32, p.sbo_rows * p.row_bytes, p.layout);
            ad |= static_cast<uint64_t>(p.base_off & 7) << 49;
            const uint64_t bd = make_smem_desc(smem_u32(sB) + k * 32, 8 * p.row_bytes, p.layout);
            umma_bf16(tmem, ad, bd, idesc, k != 0);
        }
        (void)nk;
        umma_commit(&bar[1]);
    }
    mbar_wait(&bar[1], 0);
    tc_fence_after();
    for (int cb = 0; cb < 64; cb += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + cb, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) p.out[(warp * 32 + lane) * 64 + cb + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 64);
    }
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
}

static void make_map(CUtensorMap* m, void* base, int cols, int rows, int box_rows, CUtensorMapSwizzle sw) {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t str[1] = {static_cast<cuuint64_t>(cols) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(cols), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("encode failed %d\n", (int)r);
        exit(1);
    }
}

int main() {
    CK(cudaSetDevice(0));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    const int R = 256;
    float* dout;
    CK(cudaMalloc(&dout, 128 * 64 * 4));
    std::vector<float> h(128 * 64);
    struct Mode {
        const char* name;
        int row_bytes, layout;
        CUtensorMapSwizzle sw;
    } modes[] = {{"SW128", 128, 2, CU_TENSOR_MAP_SWIZZLE_128B}, {"SW64", 64, 4, CU_TENSOR_MAP_SWIZZLE_64B},
                 {"SW32", 32, 6, CU_TENSOR_MAP_SWIZZLE_32B}};
    for (const Mode& md : modes) {
        const int cols = md.row_bytes / 2;
        for (int pass = 0; pass < 2; ++pass) {
            std::vector<__nv_bfloat16> ha(R * cols), hb(64 * cols);
            for (int r = 0; r < R; ++r)
                for (int c = 0; c < cols; ++c) ha[r * cols + c] = __float2bfloat16(pass == 0 ? (float)r : (float)c);
            for (int n = 0; n < 64; ++n)
                for (int c = 0; c < cols; ++c) hb[n * cols + c] = __float2bfloat16(n == c ? 1.f : 0.f);
            __nv_bfloat16 *da, *db;
            CK(cudaMalloc(&da, ha.size() * 2));
            CK(cudaMalloc(&db, hb.size() * 2));
            CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
            Params p;
            make_map(&p.mapA, da, cols, R, R, md.sw);
            make_map(&p.mapB, db, cols, 64, 64, md.sw);
            p.rows = R;
            p.row_bytes = md.row_bytes;
            p.ksteps = md.row_bytes / 32;
            p.layout = md.layout;
            p.out = dout;
            const int shifts[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 17};
            const int sbos[] = {8, 10, 16};
            for (int sbo : sbos)
                for (int shift : shifts)
                    for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
                        // base offset candidates: 0, or the 128-byte-granular phase of the start address in its pattern
                        const int phase = ((shift * md.row_bytes) >> 7) & 7;
                        if (bo_mode == 1 && phase == 0) continue;
                        if (15 * sbo + 7 + shift >= R) continue;
                        p.shift = shift;
                        p.sbo_rows = sbo;
                        p.base_off = bo_mode ? phase : 0;
                        CK(cudaMemset(dout, 0xff, 128 * 64 * 4));
                        probe_kernel<<<1, 128, 100 * 1024>>>(p);
                        CK(cudaDeviceSynchronize());
                        CK(cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost));
                        int bad = 0, first_m = -1, first_n = -1;
                        float got = 0, want = 0;
                        for (int m = 0; m < 128; ++m)
                            for (int n = 0; n < cols; ++n) {
                                const int srow = (m / 8) * sbo + (m % 8) + shift;
                                const float w = pass == 0 ? (float)srow : (float)n;
                                if (h[m * 64 + n] != w) {
                                    if (!bad) first_m = m, first_n = n, got = h[m * 64 + n], want = w;
                                    ++bad;
                                }
                            }
                        printf("%s pass %s sbo %2d shift %2d base_off %d : %s", md.name, pass == 0 ? "rows" : "cols", sbo, shift,
                               p.base_off, bad ? "MISMATCH" : "ok");
                        if (bad) printf(" (%d bad; first at m %d n %d got %g want %g)", bad, first_m, first_n, got, want);
                        printf("\n");
                    }
            CK(cudaFree(da));
            CK(cudaFree(db));
        }
    }
    return 0;
}
