// Exchange step of the sharded sliding window (BASELINE configs[2]: the (tile, mirror) work items of ONE case are dealt
// to R GPUs, each accumulates its share into a private fp32 accumulator; reference call site being sharded:
// run_brats2021_inference_singlethread.py:97-106 per-fold predict, :113-128 fold mean, :144-156 regions decision).
//
//   peer_finalize_kernel   ONE kernel = reduce over ranks + finalize: every rank owns a slab of the voxel range, reads
//                          that slab of ALL ranks' accumulators through NVLink peer pointers (plain ld.global on
//                          mapped peer memory), sums them in rank order (deterministic), divides by the weight sum,
//                          averages the folds, takes the regions / argmax decision and writes the uint8 labels into
//                          EVERY rank's label volume (peer stores).  Replaces all-reduce (2 x 107 MB on the wire per
//                          rank) + finalize + broadcast by (R-1)/R x 107 MB of peer reads + 9 MB of peer writes, and no
//                          rank ever holds or finalizes more than its 1/R slab.
//   bsg_nccl_*             the plain-NCCL route (one ncclAllReduce / ncclReduce of the accumulator) over the library's
//                          own communicator; libnccl is resolved at run time (dlopen), so the library loads without it.
#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>
#include "bsg_common.cuh"

namespace bsg {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxClasses = 8;
constexpr int kMaxPtrs = 128;  // K folds x R ranks

struct PeerFinalizeParams {
    const float* const* acc_table;  // device [K][R]: accumulator of fold k on rank r, each [ncls][nvox]
    uint8_t* const* seg_table;      // device [nseg]: label volumes to write (every rank's, or just the local one)
    int K, R, nseg, ncls, mode;
    int order[kMaxClasses];
    // in-kernel rank ordering (bsg_finalize_peer_signal); null flag_table: the caller orders the ranks around the launch
    uint32_t* const* flag_table;    // device [R]: every rank's flag block {arrive[R], done[R], blocks_done} as mapped here
    int rank;
    uint32_t epoch;
};

// Flag block of one rank (2R + 2 uint32): [0, R) arrive[src] = last epoch for which rank `src` announced "my accumulators are
// complete", [R, 2R) done[src] = last epoch for which rank `src` finished reading this rank's accumulators and writing
// its slab into this rank's label volumes, [2R] = local count of finished blocks, [2R + 1] = 0 or the code of a wait
// that timed out.
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounded wait for a flag to reach `epoch`: a rank that never arrives (crashed process, mismatched call sequence) must
// not wedge this GPU.  After kFlagTimeoutNs the waiter gives up, records the failure in the rank's flag block (word
// 2R + 1: the host reads it) and lets the kernel run to its end.
constexpr unsigned long long kFlagTimeoutNs = 10ull * 1000 * 1000 * 1000;
// epochs are compared modulo 2^32 (a flag can only lag or equal the epoch being waited for, never lead it by 2^31)
__device__ __forceinline__ bool reached(uint32_t flag, uint32_t epoch) { return static_cast<int32_t>(flag - epoch) >= 0; }

__device__ __forceinline__ bool exceeds_half(float a, float w) {  // fl32(a / w) > 0.5, exactly (see tail.cu)
    return static_cast<double>(a) > static_cast<double>(w) * (0.5 + 0x1p-25);
}

__global__ void __launch_bounds__(kThreads) peer_finalize_kernel(const PeerFinalizeParams fp, const float* __restrict__ wsum,
                                                                 size_t nvox, size_t v0, size_t nv) {
    __shared__ const float* s_acc[kMaxPtrs];
    __shared__ uint8_t* s_seg[16];
    for (int i = threadIdx.x; i < fp.K * fp.R; i += blockDim.x) s_acc[i] = fp.acc_table[i];
    for (int i = threadIdx.x; i < fp.nseg; i += blockDim.x) s_seg[i] = fp.seg_table[i];
    if (fp.flag_table != nullptr) {
        // (1) announce: this launch is stream-ordered behind this rank's accumulation, so its accumulators are complete —
        //     block 0 tells every rank; (2) every block waits until all ranks have announced the same epoch before it
        //     touches their accumulators.  The ranks are independent processes: a waiting block only ever waits for
        //     kernels that depend on nothing of this rank's, so the wait is bounded by the peers' own queues.
        if (blockIdx.x == 0 && threadIdx.x < fp.R) st_flag(fp.flag_table[threadIdx.x] + fp.rank, fp.epoch);
        if (threadIdx.x < fp.R) {
            uint32_t* local = fp.flag_table[fp.rank];
            const unsigned long long t0 = now_ns();
            while (!reached(ld_flag(local + threadIdx.x), fp.epoch)) {
                __nanosleep(200);
                if (now_ns() - t0 > kFlagTimeoutNs) {
                    atomicExch(local + 2 * fp.R + 1, 1u + threadIdx.x);  // 1 + the rank that never announced
                    break;
                }
            }
        }
    }
    __syncthreads();
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    const size_t ngroups = nv / 4;
    const bool by_compare = fp.K == 1 && fp.mode == 1;
    for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
        const size_t i = v0 + g * 4;
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wsum + i));
        float p[4][kMaxClasses];
#pragma unroll
        for (int k = 0; k < kMaxClasses; ++k) {
            if (k < fp.ncls) {
                float s[4] = {0.f, 0.f, 0.f, 0.f};
                for (int j = 0; j < fp.K; ++j) {
                    // sum over ranks in rank order: the same fp32 additions on every rank and in every run.  All (up
                    // to 8 per batch) peer loads of a class are in flight before the first addition.
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int r0 = 0; r0 < fp.R; r0 += 8) {
                        float4 b[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r)
                            if (r0 + r < fp.R)
                                b[r] = __ldg(reinterpret_cast<const float4*>(s_acc[j * fp.R + r0 + r] + k * nvox + i));
#pragma unroll
                        for (int r = 0; r < 8; ++r)
                            if (r0 + r < fp.R) {
                                if (r0 + r == 0)
                                    a = b[r];
                                else
                                    a.x += b[r].x, a.y += b[r].y, a.z += b[r].z, a.w += b[r].w;
                            }
                    }
                    // one fold + regions decision: fl(a / w) > 0.5 decided exactly without the division (tail.cu)
                    const float q[4] = {by_compare ? (exceeds_half(a.x, wv.x) ? 1.f : 0.f) : __fdiv_rn(a.x, wv.x),
                                        by_compare ? (exceeds_half(a.y, wv.y) ? 1.f : 0.f) : __fdiv_rn(a.y, wv.y),
                                        by_compare ? (exceeds_half(a.z, wv.z) ? 1.f : 0.f) : __fdiv_rn(a.z, wv.z),
                                        by_compare ? (exceeds_half(a.w, wv.w) ? 1.f : 0.f) : __fdiv_rn(a.w, wv.w)};
#pragma unroll
                    for (int v = 0; v < 4; ++v) s[v] = j == 0 ? q[v] : s[v] + q[v];
                }
#pragma unroll
                for (int v = 0; v < 4; ++v) p[v][k] = fp.K > 1 ? __fdiv_rn(s[v], static_cast<float>(fp.K)) : s[v];
            }
        }
        uint32_t packed = 0;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            int lab = 0;
            if (fp.mode == 0) {
                float best = p[v][0];
#pragma unroll
                for (int k = 1; k < kMaxClasses; ++k)
                    if (k < fp.ncls && p[v][k] > best) {
                        best = p[v][k];
                        lab = k;
                    }
            } else {
#pragma unroll
                for (int k = 0; k < kMaxClasses; ++k)
                    if (k < fp.ncls && p[v][k] > 0.5f) lab = fp.order[k];
            }
            packed |= static_cast<uint32_t>(lab & 255) << (8 * v);
        }
        for (int t = 0; t < fp.nseg; ++t) *reinterpret_cast<uint32_t*>(s_seg[t] + i) = packed;
    }
    if (fp.flag_table != nullptr) {
        // (3) the last block of this rank to finish tells every rank "my slab has landed in your label volumes and I am
        //     done with your accumulators", then (4) waits for the same word from all ranks: when this kernel completes,
        //     the local label volumes are whole and the local accumulators may be overwritten.
        __threadfence_system();  // this thread's peer stores are performed before the block's count below
        __syncthreads();
        __shared__ int s_last;
        uint32_t* local = fp.flag_table[fp.rank];
        if (threadIdx.x == 0) {
            const uint32_t n = atomicAdd(local + 2 * fp.R, 1u) + 1u;
            s_last = (n == gridDim.x) ? 1 : 0;
            if (s_last) local[2 * fp.R] = 0u;  // ready for the next launch (same stream: ordered)
        }
        __syncthreads();
        if (s_last && threadIdx.x < fp.R) {
            __threadfence_system();
            st_flag(fp.flag_table[threadIdx.x] + fp.R + fp.rank, fp.epoch);
            const unsigned long long t0 = now_ns();
            while (!reached(ld_flag(local + fp.R + threadIdx.x), fp.epoch)) {
                __nanosleep(200);
                if (now_ns() - t0 > kFlagTimeoutNs) {
                    atomicExch(local + 2 * fp.R + 1, 0x100u + threadIdx.x);  // 0x100 + the rank whose slab never landed
                    break;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- NCCL (dlopen)
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
    bool ok;
};

const NcclApi* nccl_api() {
    static NcclApi api = {};
    static bool tried = false;
    if (!tried) {
        tried = true;
        // by soname: inside a PyTorch process this resolves to the libnccl torch already loaded
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h != nullptr) {
            api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
            api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
            api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
            api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
            api.Reduce = reinterpret_cast<decltype(api.Reduce)>(dlsym(h, "ncclReduce"));
            api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
            api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.Reduce &&
                     api.GetErrorString;
        }
    }
    return api.ok ? &api : nullptr;
}

#define BSG_NCCL_OK(api, expr)                                                                     \
    do {                                                                                           \
        ncclResult_t _r = (expr);                                                                  \
        if (_r != ncclSuccess)                                                                     \
            return ::bsg::set_error(BSG_ECUDA, "%s failed: %s", #expr, (api)->GetErrorString(_r)); \
    } while (0)

}  // namespace
}  // namespace bsg

using namespace bsg;

extern "C" {

static int finalize_peer_launch(const float* const* acc_table_dev, int K, int R, const float* wsum, int ncls, size_t nvox,
                                size_t v0, size_t nv, int mode, const int* order_host, uint8_t* const* seg_table_dev, int nseg,
                                uint32_t* const* flag_table_dev, int rank, uint32_t epoch, void* stream) {
    BSG_REQUIRE(acc_table_dev != nullptr && wsum != nullptr && seg_table_dev != nullptr, "null argument");
    BSG_REQUIRE(K >= 1 && R >= 1 && K * R <= kMaxPtrs, "K %d x R %d accumulators (<= %d)", K, R, kMaxPtrs);
    BSG_REQUIRE(nseg >= 1 && nseg <= 16, "nseg %d (1..16)", nseg);
    BSG_REQUIRE(ncls >= 1 && ncls <= kMaxClasses, "ncls %d", ncls);
    BSG_REQUIRE(mode == 0 || (mode == 1 && order_host != nullptr), "mode %d", mode);
    BSG_REQUIRE(nvox % 4 == 0 && v0 % 4 == 0 && nv % 4 == 0 && v0 + nv <= nvox,
                "voxel range [%zu, %zu) of %zu must be 4-aligned", v0, v0 + nv, nvox);
    BSG_REQUIRE((reinterpret_cast<uintptr_t>(wsum) & 15) == 0, "wsum must be 16-byte aligned");
    BSG_REQUIRE(flag_table_dev == nullptr || (rank >= 0 && rank < R && nv > 0 && R <= kThreads),
                "in-kernel ordering needs 0 <= rank < R and a non-empty voxel range on every rank");
    if (nv == 0) return BSG_OK;
    PeerFinalizeParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.acc_table = acc_table_dev;
    fp.seg_table = seg_table_dev;
    fp.K = K;
    fp.R = R;
    fp.nseg = nseg;
    fp.ncls = ncls;
    fp.mode = mode;
    for (int k = 0; k < ncls; ++k) fp.order[k] = order_host ? order_host[k] : k;
    fp.flag_table = flag_table_dev;
    fp.rank = rank;
    fp.epoch = epoch;
    int grid = grid_for(nv / 4, kThreads, 8);
    // in-kernel ordering: blocks may sit waiting for the peers' announcements — one block per SM is plenty for a 1/R slab
    // and keeps the waiting footprint next to the other stream's conv kernels small
    if (flag_table_dev != nullptr && grid > sm_count_cached()) grid = sm_count_cached();
    peer_finalize_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(fp, wsum, nvox, v0, nv);
    BSG_CUDA_OK(cudaGetLastError());
    return BSG_OK;
}

int bsg_finalize_peer(const float* const* acc_table_dev, int K, int R, const float* wsum, int ncls, size_t nvox, size_t v0,
                      size_t nv, int mode, const int* order_host, uint8_t* const* seg_table_dev, int nseg, void* stream) {
    return finalize_peer_launch(acc_table_dev, K, R, wsum, ncls, nvox, v0, nv, mode, order_host, seg_table_dev, nseg, nullptr,
                                0, 0u, stream);
}

int bsg_finalize_peer_signal(const float* const* acc_table_dev, int K, int R, const float* wsum, int ncls, size_t nvox,
                             size_t v0, size_t nv, int mode, const int* order_host, uint8_t* const* seg_table_dev, int nseg,
                             uint32_t* const* flag_table_dev, int rank, uint32_t epoch, void* stream) {
    BSG_REQUIRE(flag_table_dev != nullptr, "null flag table");
    return finalize_peer_launch(acc_table_dev, K, R, wsum, ncls, nvox, v0, nv, mode, order_host, seg_table_dev, nseg,
                                flag_table_dev, rank, epoch, stream);
}

// ---- CUDA IPC: peers map a rank's buffer into THEIR device's address space.  The handle is opened on the consuming
// device (cudaIpcMemLazyEnablePeerAccess), which is what makes plain loads / stores from that device's kernels legal;
// a mapping created under the producing device's context (what torch's tensor sharing does) is not reachable from
// kernels running on another device.
int bsg_ipc_export(const void* ptr, void* handle64_host, size_t* offset_out) {
    BSG_REQUIRE(ptr != nullptr && handle64_host != nullptr && offset_out != nullptr, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    // the handle names the whole allocation the pointer lies in (a caching allocator may have carved the buffer out of a
    // larger cudaMalloc block): report the offset of ptr inside it
    typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    static GetRangeFn get_range = nullptr;
    if (get_range == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        BSG_CUDA_OK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || p == nullptr)
            return set_error(BSG_ECUDA, "cuMemGetAddressRange entry point not available");
        get_range = reinterpret_cast<GetRangeFn>(p);
    }
    CUdeviceptr base = 0;
    size_t size = 0;
    if (get_range(&base, &size, reinterpret_cast<CUdeviceptr>(ptr)) != CUDA_SUCCESS)
        return set_error(BSG_ECUDA, "cuMemGetAddressRange failed for %p", ptr);
    cudaIpcMemHandle_t h;
    BSG_CUDA_OK(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
    memcpy(handle64_host, &h, sizeof(h));
    *offset_out = static_cast<size_t>(reinterpret_cast<CUdeviceptr>(ptr) - base);
    return BSG_OK;
}

int bsg_ipc_open(const void* handle64_host, void** base_out) {
    BSG_REQUIRE(handle64_host != nullptr && base_out != nullptr, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, sizeof(h));
    BSG_CUDA_OK(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess));
    return BSG_OK;
}

int bsg_ipc_close(void* base) {
    if (base == nullptr) return BSG_OK;
    BSG_CUDA_OK(cudaIpcCloseMemHandle(base));
    return BSG_OK;
}

int bsg_enable_peer_access(int peer_device) {
    int dev = 0, can = 0;
    BSG_CUDA_OK(cudaGetDevice(&dev));
    if (peer_device == dev) return BSG_OK;
    BSG_CUDA_OK(cudaDeviceCanAccessPeer(&can, dev, peer_device));
    if (!can) return set_error(BSG_ECUDA, "device %d cannot access peer %d", dev, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();  // clear the sticky-free error state
        return BSG_OK;
    }
    BSG_CUDA_OK(e);
    return BSG_OK;
}

int bsg_nccl_unique_id(void* id128_host) {
    BSG_REQUIRE(id128_host != nullptr, "null argument");
    const NcclApi* api = nccl_api();
    if (api == nullptr) return set_error(BSG_ECUDA, "libnccl.so.2 could not be loaded");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    BSG_NCCL_OK(api, api->GetUniqueId(static_cast<ncclUniqueId*>(id128_host)));
    return BSG_OK;
}

int bsg_nccl_comm_create(const void* id128_host, int nranks, int rank, void** comm_out) {
    BSG_REQUIRE(id128_host != nullptr && comm_out != nullptr && nranks >= 1 && rank >= 0 && rank < nranks, "bad argument");
    const NcclApi* api = nccl_api();
    if (api == nullptr) return set_error(BSG_ECUDA, "libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    memcpy(&id, id128_host, sizeof(id));
    ncclComm_t comm = nullptr;
    BSG_NCCL_OK(api, api->CommInitRank(&comm, nranks, id, rank));
    *comm_out = comm;
    return BSG_OK;
}

int bsg_nccl_reduce_accumulator(void* comm, float* acc, size_t count, int root, void* stream) {
    BSG_REQUIRE(comm != nullptr && acc != nullptr, "null argument");
    const NcclApi* api = nccl_api();
    if (api == nullptr) return set_error(BSG_ECUDA, "libnccl.so.2 could not be loaded");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (root < 0)
        BSG_NCCL_OK(api, api->AllReduce(acc, acc, count, ncclFloat, ncclSum, static_cast<ncclComm_t>(comm), s));
    else
        BSG_NCCL_OK(api, api->Reduce(acc, acc, count, ncclFloat, ncclSum, root, static_cast<ncclComm_t>(comm), s));
    return BSG_OK;
}

int bsg_nccl_comm_destroy(void* comm) {
    if (comm == nullptr) return BSG_OK;
    const NcclApi* api = nccl_api();
    if (api == nullptr) return set_error(BSG_ECUDA, "libnccl.so.2 could not be loaded");
    BSG_NCCL_OK(api, api->CommDestroy(static_cast<ncclComm_t>(comm)));
    return BSG_OK;
}

}  // extern "C"
