// warp_rows.cuh -- warp-cooperative movement of per-Gaussian rows between HBM and shared memory.
//
// A warp owns 32 consecutive Gaussians; their K-float rows (K = 3*M = 45 SH-rest floats at degree 3) form
// ONE contiguous 32*K-float run in HBM.  Instead of 32 threads each walking its own 180-byte row (every
// load/store instruction touching 32 different sectors), the warp moves the whole run with coalesced
// 128-bit accesses into / out of a shared-memory slab; threads then use their row from shared memory
// (row stride K is odd, so lane-strided row access is bank-conflict free).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace glic {

constexpr int SH_ROW_MAX = 45;   // 15 SH-rest coefficients x 3 channels

// global rows [first, first+cnt) x K floats  ->  slab[cnt*K]
__device__ __forceinline__ void warp_load_rows(const float* __restrict__ g, size_t first, int cnt, int K, float* slab, int lane) {
    const float* src = g + first * (size_t)K;
    const int nfl = cnt * K;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const int n4 = nfl >> 2;
        for (int i = lane; i < n4; i += 32) reinterpret_cast<float4*>(slab)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        for (int i = (n4 << 2) + lane; i < nfl; i += 32) slab[i] = __ldg(src + i);
    } else {
        for (int i = lane; i < nfl; i += 32) slab[i] = __ldg(src + i);
    }
    __syncwarp();
}

// slab[cnt*K]  ->  global rows [first, first+cnt) x K floats
__device__ __forceinline__ void warp_store_rows(float* __restrict__ g, size_t first, int cnt, int K, const float* slab, int lane) {
    __syncwarp();
    float* dst = g + first * (size_t)K;
    const int nfl = cnt * K;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        const int n4 = nfl >> 2;
        for (int i = lane; i < n4; i += 32) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(slab)[i];
        for (int i = (n4 << 2) + lane; i < nfl; i += 32) dst[i] = slab[i];
    } else {
        for (int i = lane; i < nfl; i += 32) dst[i] = slab[i];
    }
}

}  // namespace glic
