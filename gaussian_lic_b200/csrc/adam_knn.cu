// adam_knn.cu -- visibility-masked Adam step and the simple-knn density initialiser.
//
// Replaces adamUpdateCUDA / ADAM::adamUpdate (reference cuda_rasterizer/adam.cu:9-67) and
// SimpleKNN::knn (simple-knn/simple_knn.cu:45-221).  The knn path reuses this library's own
// onesweep sort (no CUB / Thrust, no cudaMalloc, no host round trips for min/max).
#include "common.cuh"
#include "adam_math.cuh"
#include <algorithm>
#include <cfloat>

namespace glic {

namespace {

// One thread per scalar; element j belongs to Gaussian j / M (adam.cu:24-26).  No bias correction.
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ exp_avg,
            float* __restrict__ exp_avg_sq, const uint8_t* __restrict__ visible, float lr, float b1, float b2, float eps,
            uint32_t N, uint32_t M) {
    const size_t total = (size_t)N * M;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (size_t)gridDim.x * blockDim.x) {
        if (!visible[j / M]) continue;
        float p = param[j], m = exp_avg[j], v = exp_avg_sq[j];
        adam_element(p, m, v, grad[j], lr, b1, b2, eps);
        param[j] = p;
        exp_avg[j] = m;
        exp_avg_sq[j] = v;
    }
}

// ---- knn -----------------------------------------------------------------------------------
struct Box { float mn[3], mx[3]; };
constexpr int KNN_BOX = 1024;   // simple_knn.cu:12

__global__ void __launch_bounds__(1024) knn_minmax_kernel(int P, const float* __restrict__ pts, float* __restrict__ mm) {
    __shared__ float red[6][32];
    // seeded with the origin like the reference's DeviceReduce init = {0,0,0} (simple_knn.cu:191-200)
    float mn[3] = {0.f, 0.f, 0.f}, mx[3] = {0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < P; i += blockDim.x)
#pragma unroll
        for (int k = 0; k < 3; ++k) { const float v = pts[3 * i + k]; mn[k] = fminf(mn[k], v); mx[k] = fmaxf(mx[k], v); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 3; ++k) { red[k][warp] = mn[k]; red[3 + k][warp] = mx[k]; }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float a = red[k][lane], b = red[3 + k][lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
                b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
            }
            if (lane == 0) { mm[k] = a; mm[3 + k] = b; }
        }
    }
}

__device__ __forceinline__ uint32_t prep_morton(uint32_t x) {   // simple_knn.cu:45-52
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}

__global__ void __launch_bounds__(256)
knn_morton_kernel(int P, const float* __restrict__ pts, const float* __restrict__ mm, uint64_t* __restrict__ keys,
                  uint32_t* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    uint32_t code = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float t = ((pts[3 * i + k] - mm[k]) / (mm[3 + k] - mm[k])) * (float)((1 << 10) - 1);
        code |= prep_morton((uint32_t)t) << k;
    }
    keys[i] = code;
    vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(KNN_BOX)
knn_box_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ order, Box* __restrict__ boxes) {
    __shared__ float red[6][32];
    const int i = blockIdx.x * KNN_BOX + threadIdx.x;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < P) {
        const uint32_t id = order[i];
#pragma unroll
        for (int k = 0; k < 3; ++k) mn[k] = mx[k] = pts[3 * (size_t)id + k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 3; ++k) { red[k][warp] = mn[k]; red[3 + k][warp] = mx[k]; }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float a = red[k][lane], b = red[3 + k][lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
                b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
            }
            if (lane == 0) { boxes[blockIdx.x].mn[k] = a; boxes[blockIdx.x].mx[k] = b; }
        }
    }
}

__device__ __forceinline__ void update3(float px, float py, float pz, float qx, float qy, float qz, float* best) {
    const float dx = qx - px, dy = qy - py, dz = qz - pz;
    float d = dx * dx + dy * dy + dz * dz;
#pragma unroll
    for (int j = 0; j < 3; ++j)
        if (best[j] > d) { const float t = best[j]; best[j] = d; d = t; }
}

__device__ __forceinline__ float box_dist2(const Box& b, float x, float y, float z) {   // simple_knn.cu:119-129
    float dx = 0.f, dy = 0.f, dz = 0.f;
    if (x < b.mn[0] || x > b.mx[0]) dx = fminf(fabsf(x - b.mn[0]), fabsf(x - b.mx[0]));
    if (y < b.mn[1] || y > b.mx[1]) dy = fminf(fabsf(y - b.mn[1]), fabsf(y - b.mx[1]));
    if (z < b.mn[2] || z > b.mx[2]) dz = fminf(fabsf(z - b.mn[2]), fabsf(z - b.mx[2]));
    return dx * dx + dy * dy + dz * dz;
}

// Each point: bound from its +-3 Morton neighbours, then exact scan of every box that can still
// hold one of the three nearest (simple_knn.cu:147-183).  Sorted coordinates are staged per box in
// shared memory so the inner loop reads smem broadcasts instead of dependent global gathers.
__global__ void __launch_bounds__(KNN_BOX)
knn_dist_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ order, const Box* __restrict__ boxes,
                int nboxes, float* __restrict__ out) {
    __shared__ float sx[KNN_BOX], sy[KNN_BOX], sz[KNN_BOX];
    const int i = blockIdx.x * KNN_BOX + threadIdx.x;
    const bool valid = i < P;
    float px = 0.f, py = 0.f, pz = 0.f;
    uint32_t id = 0;
    if (valid) { id = order[i]; px = pts[3 * (size_t)id]; py = pts[3 * (size_t)id + 1]; pz = pts[3 * (size_t)id + 2]; }
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    if (valid) {
        for (int j = max(0, i - 3); j <= min(P - 1, i + 3); ++j) {
            if (j == i) continue;
            const uint32_t o = order[j];
            update3(px, py, pz, pts[3 * (size_t)o], pts[3 * (size_t)o + 1], pts[3 * (size_t)o + 2], best);
        }
    }
    const float reject = best[2];
    best[0] = best[1] = best[2] = FLT_MAX;
    for (int b = 0; b < nboxes; ++b) {
        const Box bx = boxes[b];
        const float d = box_dist2(bx, px, py, pz);
        const bool need = valid && !(d > reject || d > best[2]);
        if (__syncthreads_or(need)) {
            const int j = b * KNN_BOX + threadIdx.x;
            if (j < P) {
                const uint32_t o = order[j];
                sx[threadIdx.x] = pts[3 * (size_t)o]; sy[threadIdx.x] = pts[3 * (size_t)o + 1]; sz[threadIdx.x] = pts[3 * (size_t)o + 2];
            }
            __syncthreads();
            if (need) {
                const int cnt = min(KNN_BOX, P - b * KNN_BOX);
                for (int k = 0; k < cnt; ++k) {
                    if (b * KNN_BOX + k == i) continue;
                    update3(px, py, pz, sx[k], sy[k], sz[k], best);
                }
            }
            __syncthreads();
        }
    }
    if (valid) out[id] = (best[0] + best[1] + best[2]) / 3.0f;
}

struct KnnTemp {
    float* mm;
    Box* boxes;
    uint64_t* keys[2];
    uint32_t* vals[2];
    void* sort_temp;
    size_t sort_temp_size;
    static KnnTemp carve(void* ws, int P, size_t* bytes) {
        const size_t n = P > 0 ? P : 1;
        Carver c(ws);
        KnnTemp t;
        t.mm = c.take<float>(8);
        t.boxes = c.take<Box>((n + KNN_BOX - 1) / KNN_BOX);
        t.keys[0] = c.take<uint64_t>(n); t.keys[1] = c.take<uint64_t>(n);
        t.vals[0] = c.take<uint32_t>(n); t.vals[1] = c.take<uint32_t>(n);
        t.sort_temp_size = sort_temp_bytes(P);
        t.sort_temp = c.take<char>(t.sort_temp_size);
        if (bytes) *bytes = c.total();
        return t;
    }
};

}  // namespace

}  // namespace glic

using namespace glic;

extern "C" int glic_adam_update(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const uint8_t* visible,
                                float lr, float b1, float b2, float eps, uint32_t N, uint32_t M, void* stream) {
    const size_t total = (size_t)N * M;
    if (total == 0) return GLIC_OK;       // an empty group (sh-rest at degree 0: [P,0,3]) has null data pointers; the reference launches nothing
    if (!param || !grad || !exp_avg || !exp_avg_sq || !visible) { set_error("adam: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148u * 32u);
    StageTimer _t(GLIC_STAGE_ADAM, (cudaStream_t)stream);
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, visible, lr, b1, b2, eps, N, M);
    GLIC_LAUNCH_CHECK();
    return GLIC_OK;
}

extern "C" size_t glic_knn_temp_bytes(int P) {
    size_t b = 0;
    KnnTemp::carve(nullptr, P, &b);
    return b;
}

extern "C" int glic_knn_mean_dist2(int P, const float* points, float* mean_dists, void* temp, size_t temp_bytes, void* stream) {
    if (P < 0 || (P > 0 && (!points || !mean_dists || !temp))) { set_error("knn: bad arguments"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (P == 0) return GLIC_OK;
    if (temp_bytes < glic_knn_temp_bytes(P)) { set_error("knn: temp too small"); return GLIC_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    KnnTemp t = KnnTemp::carve(temp, P, nullptr);
    knn_minmax_kernel<<<1, 1024, 0, s>>>(P, points, t.mm);
    GLIC_LAUNCH_CHECK();
    knn_morton_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, points, t.mm, t.keys[0], t.vals[0]);
    GLIC_LAUNCH_CHECK();
    const int cur = launch_sort_pairs(P, 30, t.keys, t.vals, t.sort_temp, t.sort_temp_size, s);
    if (cur < 0) return cur;
    const int nboxes = (P + KNN_BOX - 1) / KNN_BOX;
    knn_box_kernel<<<nboxes, KNN_BOX, 0, s>>>(P, points, t.vals[cur], t.boxes);
    GLIC_LAUNCH_CHECK();
    knn_dist_kernel<<<nboxes, KNN_BOX, 0, s>>>(P, points, t.vals[cur], t.boxes, nboxes, mean_dists);
    GLIC_LAUNCH_CHECK();
    return GLIC_OK;
}
