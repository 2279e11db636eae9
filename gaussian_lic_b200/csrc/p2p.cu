// p2p.cu -- in-place mean all-reduce of the packed gradient buffer over NVLink peer memory (our own collective).
//
// The view-sharded mapping iteration has exactly one exchange step: every rank holds the
// per-Gaussian gradients of its own view and needs their mean, plus the union of the visibility bytes.
// Instead of handing that to a library, each rank maps every peer's gradient buffer (CUDA IPC) and runs a
// two-shot all-reduce itself: rank r owns slice r of the buffer, loads that slice from all peers straight
// over NVLink (128-bit loads), reduces in registers and stores the result back into every peer's buffer
// (128-bit stores).  Slice r of any buffer is read and written by rank r only, so the exchange is in place.
// Two tiny signalling kernels (system-scope release/acquire flags in the mapped buffers) order it against
// the peers' backward kernels and against their readers.  Bytes over NVLink per rank: 2*(N-1)/N of the
// payload, the minimum for an all-reduce.
#include "common.cuh"
#include <algorithm>
#include <cstring>
#include <cstdlib>

namespace glic {

namespace {

constexpr int P2P_MAX_RANKS = 8;

struct PeerBufs { char* p[P2P_MAX_RANKS]; };

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Thread q tells peer q "rank `rank` reached epoch e", then waits until peer q has told us the same.
// The epoch lives in the rank's own flag block (word P2P_MAX_RANKS) and advances by one per barrier, so the
// launch carries no per-call argument and the whole exchange can sit inside a replayed CUDA graph.
__global__ void p2p_barrier_kernel(PeerBufs bufs, int rank, int world, size_t flag_off) {
    const int q = threadIdx.x;
    uint32_t epoch = 0;
    if (q == 0) {
        uint32_t* ctr = reinterpret_cast<uint32_t*>(bufs.p[rank] + flag_off) + P2P_MAX_RANKS;
        epoch = *ctr + 1;
        *ctr = epoch;
    }
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    if (q >= world) return;
    __threadfence_system();
    uint32_t* remote = reinterpret_cast<uint32_t*>(bufs.p[q] + flag_off) + rank;
    st_release_sys(remote, epoch);
    const uint32_t* local = reinterpret_cast<const uint32_t*>(bufs.p[rank] + flag_off) + q;
    unsigned long long spins = 0;
    while ((int)(ld_acquire_sys(local) - epoch) < 0) {
        if (++spins > (1ull << 28)) __trap();       // a lost peer must not hang this GPU forever
    }
}

// slice [lo4, hi4) in float4 units: mean over ranks; [vlo4, vhi4) in uint4 units: bitwise OR (visibility bytes are 0/1).
// UNROLL independent 16-byte loads per peer are issued before the first use so that a thread keeps
// world*UNROLL NVLink requests in flight (remote-load latency is a few microseconds).
template <int WORLD, int UNROLL>
__global__ void __launch_bounds__(512)
p2p_reduce_kernel(PeerBufs bufs, size_t lo4, size_t hi4, size_t vis_off, size_t vlo4, size_t vhi4, float inv) {
    const size_t tile = (size_t)blockDim.x * UNROLL;
    const size_t stride = (size_t)gridDim.x * tile;
    for (size_t base = lo4 + (size_t)blockIdx.x * tile + threadIdx.x; base < hi4; base += stride) {
        float4 v[WORLD][UNROLL];
#pragma unroll
        for (int q = 0; q < WORLD; ++q)
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const size_t i = base + (size_t)u * blockDim.x;
                v[q][u] = i < hi4 ? __ldcv(reinterpret_cast<const float4*>(bufs.p[q]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const size_t i = base + (size_t)u * blockDim.x;
            float4 acc = v[0][u];
#pragma unroll
            for (int q = 1; q < WORLD; ++q) { acc.x += v[q][u].x; acc.y += v[q][u].y; acc.z += v[q][u].z; acc.w += v[q][u].w; }
            acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
            if (i < hi4) {
#pragma unroll
                for (int q = 0; q < WORLD; ++q) reinterpret_cast<float4*>(bufs.p[q])[i] = acc;
            }
        }
    }
    const size_t vstride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = vlo4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < vhi4; i += vstride) {
        uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int q = 0; q < WORLD; ++q) {
            const uint4 w = __ldcv(reinterpret_cast<const uint4*>(bufs.p[q] + vis_off) + i);
            acc.x |= w.x; acc.y |= w.y; acc.z |= w.z; acc.w |= w.w;
        }
#pragma unroll
        for (int q = 0; q < WORLD; ++q) reinterpret_cast<uint4*>(bufs.p[q] + vis_off)[i] = acc;
    }
}

template <int WORLD, int UNROLL>
void launch_reduce(const PeerBufs& pb, size_t lo4, size_t hi4, size_t vis_off, size_t vlo4, size_t vhi4, int ctas_per_sm, cudaStream_t s) {
    const size_t work = std::max((hi4 - lo4 + UNROLL - 1) / UNROLL, vhi4 - vlo4);
    const unsigned blocks = (unsigned)std::max<size_t>(1, std::min<size_t>((work + 511) / 512, (size_t)148 * ctas_per_sm));
    p2p_reduce_kernel<WORLD, UNROLL><<<blocks, 512, 0, s>>>(pb, lo4, hi4, vis_off, vlo4, vhi4, 1.0f / (float)WORLD);
}

}  // namespace
}  // namespace glic

using namespace glic;

extern "C" {

size_t glic_p2p_buffer_bytes(size_t n_floats, size_t n_vis_bytes) {
    const size_t f = (n_floats * 4 + 255) & ~size_t(255);
    const size_t v = (n_vis_bytes + 255) & ~size_t(255);
    return f + v + 256;                     // + flag words (one per rank)
}

int glic_p2p_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64) {
    if (!dev_ptr || !handle64) { set_error("p2p_alloc: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    GLIC_CUDA_TRY(cudaMalloc(dev_ptr, bytes));
    GLIC_CUDA_TRY(cudaMemset(*dev_ptr, 0, bytes));
    cudaIpcMemHandle_t h;
    GLIC_CUDA_TRY(cudaIpcGetMemHandle(&h, *dev_ptr));
    static_assert(sizeof(h) == 64, "ipc handle size");
    memcpy(handle64, &h, 64);
    return GLIC_OK;
}

int glic_p2p_open(const unsigned char* handle64, void** peer_ptr) {
    if (!handle64 || !peer_ptr) { set_error("p2p_open: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    GLIC_CUDA_TRY(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GLIC_OK;
}

int glic_p2p_close(void* peer_ptr) { GLIC_CUDA_TRY(cudaIpcCloseMemHandle(peer_ptr)); return GLIC_OK; }
int glic_p2p_free(void* dev_ptr) { GLIC_CUDA_TRY(cudaFree(dev_ptr)); return GLIC_OK; }

// Rank r's share of the two-shot exchange, in 16-byte units: floats [lo4, hi4), visibility [vlo4, vhi4) (relative to the
// visibility block at byte f_bytes); also the byte offsets of the visibility block and of the flag words.  The slices of the
// `world` ranks partition both blocks exactly (tests/test_cabi.py), so every 16-byte unit is reduced by exactly one rank.
int glic_p2p_slice(int rank, int world, size_t n_floats, size_t n_vis_bytes, size_t* out6) {
    if (world < 1 || world > P2P_MAX_RANKS || rank < 0 || rank >= world || !out6) { set_error("p2p_slice: bad rank / world"); return GLIC_ERR_INVALID_ARGUMENT; }
    const size_t f_bytes = (n_floats * 4 + 255) & ~size_t(255);
    const size_t v_bytes = (n_vis_bytes + 255) & ~size_t(255);
    const size_t f4 = f_bytes / 16, v4 = v_bytes / 16;
    const size_t fper = (f4 + world - 1) / world, vper = (v4 + world - 1) / world;
    out6[0] = std::min(f4, fper * rank); out6[1] = std::min(f4, fper * (rank + 1));
    out6[2] = std::min(v4, vper * rank); out6[3] = std::min(v4, vper * (rank + 1));
    out6[4] = f_bytes; out6[5] = f_bytes + v_bytes;
    return GLIC_OK;
}

int glic_p2p_allreduce_mean(int rank, int world, void* const* bufs_host, size_t n_floats, size_t n_vis_bytes, void* stream) {
    if (world < 1 || world > P2P_MAX_RANKS || rank < 0 || rank >= world || !bufs_host) { set_error("p2p_allreduce: bad rank / world"); return GLIC_ERR_INVALID_ARGUMENT; }
    cudaStream_t s = (cudaStream_t)stream;
    PeerBufs pb;
    for (int q = 0; q < P2P_MAX_RANKS; ++q) pb.p[q] = q < world ? static_cast<char*>(bufs_host[q]) : nullptr;
    size_t sl[6];
    glic_p2p_slice(rank, world, n_floats, n_vis_bytes, sl);
    const size_t lo4 = sl[0], hi4 = sl[1], vlo4 = sl[2], vhi4 = sl[3], f_bytes = sl[4], flag_off = sl[5];
    { StageTimer _t(GLIC_STAGE_ALLREDUCE, s);
      p2p_barrier_kernel<<<1, 32, 0, s>>>(pb, rank, world, flag_off);         // every peer's gradients are complete
      GLIC_LAUNCH_CHECK();
      static const int ctas = getenv("GLIC_P2P_CTAS") ? atoi(getenv("GLIC_P2P_CTAS")) : 2;   // tuning knobs, see profiles/
      static const int unr = getenv("GLIC_P2P_UNROLL") ? atoi(getenv("GLIC_P2P_UNROLL")) : 4;
#define GLIC_P2P_CASE(WD) case WD: if (unr >= 4 && WD <= 4) launch_reduce<WD, 4>(pb, lo4, hi4, f_bytes, vlo4, vhi4, ctas, s); \
                                   else if (unr >= 2) launch_reduce<WD, 2>(pb, lo4, hi4, f_bytes, vlo4, vhi4, ctas, s); \
                                   else launch_reduce<WD, 1>(pb, lo4, hi4, f_bytes, vlo4, vhi4, ctas, s); break;
      switch (world) {
          GLIC_P2P_CASE(1) GLIC_P2P_CASE(2) GLIC_P2P_CASE(3) GLIC_P2P_CASE(4) GLIC_P2P_CASE(5) GLIC_P2P_CASE(6) GLIC_P2P_CASE(7) GLIC_P2P_CASE(8)
      }
#undef GLIC_P2P_CASE
      GLIC_LAUNCH_CHECK();
      p2p_barrier_kernel<<<1, 32, 0, s>>>(pb, rank, world, flag_off);      // every slice has been written everywhere
      GLIC_LAUNCH_CHECK(); }
    return GLIC_OK;
}

}  // extern "C"
