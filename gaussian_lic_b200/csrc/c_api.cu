// c_api.cu -- extern "C" entry points of libglic_b200.so (include/glic_b200.h): argument
// validation, workspace carving, stage sequencing.  No torch types, no allocation, no exceptions.
#include "common.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace glic {

static thread_local std::string g_error;
unsigned long long g_launches = 0;

void set_error(const std::string& s) { g_error = s; }

// ---- stage profiling ----------------------------------------------------------------------
namespace {
struct EvPair { cudaEvent_t a, b; };
bool g_prof = false;
std::vector<EvPair> g_pairs[GLIC_STAGE_COUNT];
std::vector<EvPair> g_pool;
}
bool profiling_on() { return g_prof; }
void profile_mark(int stage, bool begin, cudaStream_t s) {
    if (stage < 0 || stage >= GLIC_STAGE_COUNT) return;
    if (begin) {
        EvPair p;
        if (!g_pool.empty()) { p = g_pool.back(); g_pool.pop_back(); }
        else { cudaEventCreate(&p.a); cudaEventCreate(&p.b); }
        cudaEventRecord(p.a, s);
        g_pairs[stage].push_back(p);
    } else if (!g_pairs[stage].empty()) {
        cudaEventRecord(g_pairs[stage].back().b, s);
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int check_view(const glic_view* v) {
    if (!v || !v->viewmatrix || !v->projmatrix || !v->campos) { set_error("view: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (v->width <= 0 || v->height <= 0) { set_error("view: non-positive image size"); return GLIC_ERR_INVALID_ARGUMENT; }
    return GLIC_OK;
}

// ---- debug copy-out kernels ---------------------------------------------------------------
__global__ void debug_geom_kernel(int P, GeomState g, float* depth, float* xy, float* conic_opacity, float* rgb,
                                  uint32_t* tiles, uint32_t* offsets, uint8_t* clamped) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float4 r0 = g.rec[3 * i], r1 = g.rec[3 * i + 1], r2 = g.rec[3 * i + 2];
    if (depth) depth[i] = r2.y;
    if (xy) { xy[2 * i] = r0.x; xy[2 * i + 1] = r0.y; }
    if (conic_opacity) { conic_opacity[4 * i] = r0.z; conic_opacity[4 * i + 1] = r0.w; conic_opacity[4 * i + 2] = r1.x; conic_opacity[4 * i + 3] = r1.y; }
    if (rgb) { rgb[3 * i] = r1.z; rgb[3 * i + 1] = r1.w; rgb[3 * i + 2] = r2.x; }
    if (tiles) tiles[i] = g.tiles[i];
    if (offsets) offsets[i] = g.offsets[i];
    if (clamped) { const unsigned c = g.clamped[i]; clamped[3 * i] = c & 1u; clamped[3 * i + 1] = (c >> 1) & 1u; clamped[3 * i + 2] = (c >> 2) & 1u; }
}

__global__ void debug_keys_kernel(long long R, GeomState g, const uint32_t* tile_keys, const uint32_t* point_list, uint64_t* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t gid = point_list[i];
    out[i] = ((uint64_t)tile_keys[i] << 32) | (uint64_t)__float_as_uint(g.rec[3 * (size_t)gid + 2].y);
}

}  // namespace glic

using namespace glic;

extern "C" {

const char* glic_last_error(void) { return g_error.c_str(); }
int glic_abi_version(void) { return GLIC_ABI_VERSION; }
uint64_t glic_launch_count(void) { return g_launches; }

int glic_profile_enable(int on) {
    for (int i = 0; i < GLIC_STAGE_COUNT; ++i) { for (auto& p : g_pairs[i]) g_pool.push_back(p); g_pairs[i].clear(); }
    g_prof = on != 0;
    return GLIC_OK;
}
int glic_profile_read(float* ms_host, int* count_host) {
    if (!ms_host || !count_host) { set_error("profile_read: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    for (int i = 0; i < GLIC_STAGE_COUNT; ++i) {
        float tot = 0.f; int n = 0;
        for (auto& p : g_pairs[i]) {
            if (cudaEventSynchronize(p.b) != cudaSuccess) { cudaGetLastError(); continue; }
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { tot += ms; ++n; } else cudaGetLastError();
        }
        ms_host[i] = tot; count_host[i] = n;
    }
    return GLIC_OK;
}

size_t glic_geom_bytes(int P) { size_t b = 0; GeomState::carve(nullptr, P > 0 ? P : 0, &b); return b; }
size_t glic_image_bytes(int width, int height) { size_t b = 0; ImageState::carve(nullptr, width, height, &b); return b; }
size_t glic_binning_bytes(int64_t R) { size_t b = 0; BinningState::carve(nullptr, R, &b); return b; }
int64_t glic_max_buckets(int64_t R, int width, int height) {
    // B = sum_t ceil(n_t/32) <= floor(R/32) + min(T, R): at most one partial bucket per non-empty tile
    const int64_t T = (int64_t)((width + TILE - 1) / TILE) * ((height + TILE - 1) / TILE);
    return R / BUCKET + std::min<int64_t>(T, R);
}
size_t glic_sample_bytes(int64_t R, int width, int height) {
    size_t b = 0;
    SampleState::carve(nullptr, glic_max_buckets(R, width, height), &b);
    return b;
}
size_t glic_sort_temp_bytes(int64_t n) { return sort_temp_bytes(n); }

int glic_forward_preprocess(int P, int sh_degree, int M, const float* means3D, const float* scales, float scale_modifier,
                            const float* rotations, const float* opacities, const float* dc, const float* sh,
                            const glic_view* view, int no_color, int* radii, void* geom_ws, size_t geom_bytes,
                            void* image_ws, size_t image_bytes, int64_t* num_rendered_host, void* stream) {
    if (int e = check_view(view)) return e;
    if (P < 0 || sh_degree < 0 || sh_degree > 3 || M < 0) { set_error("forward_preprocess: bad P / sh_degree / M"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (!num_rendered_host) { set_error("forward_preprocess: num_rendered_host is NULL"); return GLIC_ERR_INVALID_ARGUMENT; }
    *num_rendered_host = 0;
    if (!no_color && (sh_degree + 1) * (sh_degree + 1) - 1 > M) { set_error("forward_preprocess: sh has fewer coefficients than sh_degree needs"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (!image_ws || image_bytes < glic_image_bytes(view->width, view->height)) { set_error("forward_preprocess: image workspace too small"); return GLIC_ERR_WORKSPACE; }
    if (P == 0) return GLIC_OK;
    if (!means3D || !scales || !rotations || !opacities || !radii || (!no_color && (!dc || (M > 0 && sh_degree > 0 && !sh)))) {
        set_error("forward_preprocess: null input pointer"); return GLIC_ERR_INVALID_ARGUMENT;
    }
    if (!aligned16(rotations)) { set_error("forward_preprocess: rotations must be 16-byte aligned"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (!geom_ws || geom_bytes < glic_geom_bytes(P)) { set_error("forward_preprocess: geometry workspace too small"); return GLIC_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    GeomState g = GeomState::carve(geom_ws, P);
    const ViewParams vp = make_view_params(view);
    { StageTimer _t(GLIC_STAGE_PREPROCESS, s);
    if (int e = launch_preprocess_forward(P, sh_degree, M, means3D, scales, scale_modifier, rotations, opacities, dc, sh, vp,
                                          no_color != 0, radii, g, s)) return e; }
    // R is complete as soon as the preprocess kernel is: ship it to pinned host memory NOW and keep the GPU busy with the
    // depth sort while the host waits for exactly that copy (the reference blocks here too, rasterizer_impl.cu:398, but
    // with an idle GPU).
    struct RChannel { unsigned int* pinned; cudaEvent_t ready, done; cudaStream_t stream; };
    static thread_local RChannel r_channels[64] = {};        // one per device: streams and events are device-bound
    int dev = 0;
    GLIC_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("forward_preprocess: device index out of range"); return GLIC_ERR_INVALID_ARGUMENT; }
    RChannel& rc = r_channels[dev];
    if (!rc.pinned) {
        GLIC_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&rc.pinned), sizeof(unsigned int), cudaHostAllocDefault));
        GLIC_CUDA_TRY(cudaEventCreateWithFlags(&rc.ready, cudaEventDisableTiming));
        GLIC_CUDA_TRY(cudaEventCreateWithFlags(&rc.done, cudaEventDisableTiming));
        GLIC_CUDA_TRY(cudaStreamCreateWithFlags(&rc.stream, cudaStreamNonBlocking));   // the 4-byte copy must not sit in front of the depth sort
    }
    unsigned int* r_pinned = rc.pinned;
    cudaEvent_t r_ready = rc.ready, r_event = rc.done;
    cudaStream_t r_stream = rc.stream;
    GLIC_CUDA_TRY(cudaEventRecord(r_ready, s));
    GLIC_CUDA_TRY(cudaStreamWaitEvent(r_stream, r_ready, 0));
    GLIC_CUDA_TRY(cudaMemcpyAsync(r_pinned, &g.hdr->total, sizeof(unsigned int), cudaMemcpyDeviceToHost, r_stream));
    GLIC_CUDA_TRY(cudaEventRecord(r_event, r_stream));
    {   // depth-first binning: order the Gaussians by (depth, index) once, then prefix-sum their tile counts in that order
        StageTimer _t(GLIC_STAGE_SORT, s);
        const int cur = launch_sort_pairs32(P, 32, g.depth_keys, g.order, g.sort_temp, g.sort_temp_size, s, nullptr, /*hist_ready=*/true);
        if (cur < 0) return cur;
        if (int e = launch_depth_scan(P, g, g.order[cur], 0xFFFFFFFFll, s)) return e;
    }
    GLIC_CUDA_TRY(cudaEventSynchronize(r_event));
    const unsigned int total = *r_pinned;
    *num_rendered_host = (int64_t)total;
    return GLIC_OK;
}

int glic_forward_render(int P, const glic_view* view, int no_color, int64_t R, void* geom_ws, void* image_ws,
                        void* binning_ws, size_t binning_bytes, void* sample_ws, size_t sample_bytes, float* out_color,
                        float* out_final_T, int64_t* num_buckets_host, void* stream) {
    if (int e = check_view(view)) return e;
    if (P < 0 || R < 0 || !image_ws || !out_final_T || (!no_color && !out_color)) { set_error("forward_render: bad arguments"); return GLIC_ERR_INVALID_ARGUMENT; }
    cudaStream_t s = (cudaStream_t)stream;
    const ViewParams vp = make_view_params(view);
    const int T = vp.grid_x * vp.grid_y;
    const size_t HW = (size_t)vp.W * vp.H;
    if (num_buckets_host) *num_buckets_host = 0;
    if (P == 0) {   // rasterize_points.cu:110: zero outputs
        GLIC_CUDA_TRY(cudaMemsetAsync(out_final_T, 0, sizeof(float) * HW, s));
        if (out_color) GLIC_CUDA_TRY(cudaMemsetAsync(out_color, 0, sizeof(float) * 3 * HW, s));
        return GLIC_OK;
    }
    if (!geom_ws) { set_error("forward_render: geometry workspace is NULL"); return GLIC_ERR_WORKSPACE; }
    if (!binning_ws || binning_bytes < glic_binning_bytes(R)) { set_error("forward_render: binning workspace too small"); return GLIC_ERR_WORKSPACE; }
    const int64_t max_buckets = no_color ? 0 : glic_max_buckets(R, vp.W, vp.H);
    if (!no_color && (!sample_ws || sample_bytes < glic_sample_bytes(R, vp.W, vp.H))) { set_error("forward_render: sample workspace too small"); return GLIC_ERR_WORKSPACE; }
    GeomState g = GeomState::carve(geom_ws, P);
    ImageState img = ImageState::carve(image_ws, vp.W, vp.H);
    BinningState bin = BinningState::carve(binning_ws, R);
    SampleState smp = SampleState::carve(no_color ? nullptr : sample_ws, max_buckets);

    int cur = 0;
    if (R > 0) {
        const int bit = (int)higher_msb((uint32_t)T);                       // rasterizer_impl.cu:417
        { StageTimer _t(GLIC_STAGE_EMIT, s);
          if (int e = sort_prepare(R, bit, bin.sort_temp, s)) return e;      // emit builds the tile sort's digit histograms
          if (int e = launch_emit_keys(P, vp, g, bin.keys[0], bin.vals[0], R, sort_hist(bin.sort_temp), bit, s)) return e; }
        // pairs arrive ordered by (depth, index): a stable sort on the tile bits alone finishes the job
        { StageTimer _t(GLIC_STAGE_SORT, s); cur = launch_sort_pairs32(R, bit, bin.keys, bin.vals, bin.sort_temp, bin.sort_temp_size, s, nullptr, true); }
        if (cur < 0) return cur;
    }
    GLIC_CUDA_TRY(cudaMemsetAsync(&bin.hdr->sorted_in_b, cur, 1, s));       // 0/1 in the low byte (header is zero-padded)
    GLIC_CUDA_TRY(cudaMemsetAsync(reinterpret_cast<char*>(&bin.hdr->sorted_in_b) + 1, 0, 3, s));
    { StageTimer _t(GLIC_STAGE_RANGES, s); if (int e = launch_tile_ranges(R, nullptr, bin.keys[cur], T, img, !no_color, s)) return e; }
    { StageTimer _t(GLIC_STAGE_RENDER_FWD, s); if (int e = launch_render_forward(vp, no_color != 0, bin.vals[cur], g, img, smp, out_color, out_final_T, s)) return e; }
    if (num_buckets_host) {
        unsigned int nb = 0;
        GLIC_CUDA_TRY(cudaMemcpyAsync(&nb, &img.hdr->num_buckets, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
        GLIC_CUDA_TRY(cudaStreamSynchronize(s));
        *num_buckets_host = nb;
    }
    return GLIC_OK;
}

int64_t glic_binning_capacity(size_t binning_bytes, size_t sample_bytes, int width, int height, int no_color) {
    // largest R whose workspaces fit: both size functions are monotone in R
    int64_t lo = 0, hi = (int64_t)1 << 31;
    while (lo < hi) {
        const int64_t mid = lo + (hi - lo + 1) / 2;
        const bool fits = glic_binning_bytes(mid) <= binning_bytes && (no_color || glic_sample_bytes(mid, width, height) <= sample_bytes);
        if (fits) lo = mid; else hi = mid - 1;
    }
    return lo;
}

int glic_forward(int P, int sh_degree, int M, const float* means3D, const float* scales, float scale_modifier,
                 const float* rotations, const float* opacities, const float* dc, const float* sh, const glic_view* view,
                 int no_color, int* radii, void* geom_ws, size_t geom_bytes, void* image_ws, size_t image_bytes, void* binning_ws,
                 size_t binning_bytes, void* sample_ws, size_t sample_bytes, float* out_color, float* out_final_T,
                 int64_t* counters_host, void* stream) {
    if (int e = check_view(view)) return e;
    if (P < 0 || sh_degree < 0 || sh_degree > 3 || M < 0) { set_error("forward: bad P / sh_degree / M"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (!no_color && (sh_degree + 1) * (sh_degree + 1) - 1 > M) { set_error("forward: sh has fewer coefficients than sh_degree needs"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (!image_ws || image_bytes < glic_image_bytes(view->width, view->height)) { set_error("forward: image workspace too small"); return GLIC_ERR_WORKSPACE; }
    if (!out_final_T || (!no_color && !out_color)) { set_error("forward: null output"); return GLIC_ERR_INVALID_ARGUMENT; }
    cudaStream_t s = (cudaStream_t)stream;
    const ViewParams vp = make_view_params(view);
    const int T = vp.grid_x * vp.grid_y;
    const size_t HW = (size_t)vp.W * vp.H;
    ImageState img = ImageState::carve(image_ws, vp.W, vp.H);
    if (P == 0) {
        GLIC_CUDA_TRY(cudaMemsetAsync(out_final_T, 0, sizeof(float) * HW, s));
        if (out_color) GLIC_CUDA_TRY(cudaMemsetAsync(out_color, 0, sizeof(float) * 3 * HW, s));
        GLIC_CUDA_TRY(cudaMemsetAsync(img.hdr, 0, sizeof(ImageHeader), s));
        if (counters_host) GLIC_CUDA_TRY(cudaMemcpyAsync(counters_host, img.hdr->counters, 3 * sizeof(long long), cudaMemcpyDeviceToHost, s));
        return GLIC_OK;
    }
    if (!means3D || !scales || !rotations || !opacities || !radii || (!no_color && (!dc || (M > 0 && sh_degree > 0 && !sh)))) {
        set_error("forward: null input pointer"); return GLIC_ERR_INVALID_ARGUMENT;
    }
    if (!aligned16(rotations)) { set_error("forward: rotations must be 16-byte aligned"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (!geom_ws || geom_bytes < glic_geom_bytes(P)) { set_error("forward: geometry workspace too small"); return GLIC_ERR_WORKSPACE; }
    const int64_t cap = glic_binning_capacity(binning_bytes, sample_bytes, vp.W, vp.H, no_color);
    if (!binning_ws || cap < 1 || (!no_color && !sample_ws)) { set_error("forward: binning / sample workspace missing or too small"); return GLIC_ERR_WORKSPACE; }
    GeomState g = GeomState::carve(geom_ws, P);
    BinningState bin = BinningState::carve(binning_ws, cap);
    const int64_t max_buckets = no_color ? 0 : glic_max_buckets(cap, vp.W, vp.H);
    SampleState smp = SampleState::carve(no_color ? nullptr : sample_ws, max_buckets);
    { StageTimer _t(GLIC_STAGE_PREPROCESS, s);
      if (int e = launch_preprocess_forward(P, sh_degree, M, means3D, scales, scale_modifier, rotations, opacities, dc, sh, vp,
                                            no_color != 0, radii, g, s)) return e; }
    { StageTimer _t(GLIC_STAGE_SORT, s);
      const int cur0 = launch_sort_pairs32(P, 32, g.depth_keys, g.order, g.sort_temp, g.sort_temp_size, s, nullptr, /*hist_ready=*/true);
      if (cur0 < 0) return cur0;
      if (int e = launch_depth_scan(P, g, g.order[cur0], cap, s)) return e; }
    const int bit = (int)higher_msb((uint32_t)T);
    { StageTimer _t(GLIC_STAGE_EMIT, s);
      if (int e = sort_prepare(cap, bit, bin.sort_temp, s)) return e;        // emit builds the tile sort's digit histograms
      if (int e = launch_emit_keys(P, vp, g, bin.keys[0], bin.vals[0], cap, sort_hist(bin.sort_temp), bit, s)) return e; }
    int cur;
    { StageTimer _t(GLIC_STAGE_SORT, s);
      cur = launch_sort_pairs32(cap, bit, bin.keys, bin.vals, bin.sort_temp, bin.sort_temp_size, s, &g.hdr->r_eff, true);
      if (cur < 0) return cur; }
    GLIC_CUDA_TRY(cudaMemsetAsync(&bin.hdr->sorted_in_b, cur, 1, s));       // 0/1 in the low byte (header is zero-padded)
    GLIC_CUDA_TRY(cudaMemsetAsync(reinterpret_cast<char*>(&bin.hdr->sorted_in_b) + 1, 0, 3, s));
    { StageTimer _t(GLIC_STAGE_RANGES, s);
      if (int e = launch_tile_ranges(cap, g.hdr, bin.keys[cur], T, img, !no_color, s)) return e; }
    { StageTimer _t(GLIC_STAGE_RENDER_FWD, s);
      if (int e = launch_render_forward(vp, no_color != 0, bin.vals[cur], g, img, smp, out_color, out_final_T, s)) return e; }
    if (counters_host) GLIC_CUDA_TRY(cudaMemcpyAsync(counters_host, img.hdr->counters, 3 * sizeof(long long), cudaMemcpyDeviceToHost, s));
    return GLIC_OK;
}

int glic_backward(int P, int sh_degree, int M, const float* means3D, const float* scales, float scale_modifier,
                  const float* rotations, const float* dc, const float* sh, const glic_view* view, const int* radii,
                  int64_t R, const void* geom_ws, const void* binning_ws, const void* image_ws, const void* sample_ws,
                  const float* dL_dpix, float lambda_erank, float* dL_dmeans2D, float* dL_dconic, float* dL_dopacity,
                  float* dL_dcolors, float* dL_dmeans3D, float* dL_dcov3D, float* dL_ddc, float* dL_dsh, float* dL_dscales,
                  float* dL_drotations, void* stream) {
    (void)dc;
    if (int e = check_view(view)) return e;
    if (P < 0 || R < 0 || sh_degree < 0 || sh_degree > 3 || M < 0) { set_error("backward: bad sizes"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (P == 0) return GLIC_OK;
    if (!means3D || !scales || !rotations || !radii || !dL_dpix || !dL_dmeans2D || !dL_dconic || !dL_dopacity || !dL_dcolors ||
        !dL_dmeans3D || !dL_dcov3D || !dL_ddc || !dL_dscales || !dL_drotations || (M > 0 && (!sh || !dL_dsh))) {
        set_error("backward: null pointer"); return GLIC_ERR_INVALID_ARGUMENT;
    }
    if (!aligned16(rotations) || !aligned16(dL_drotations)) { set_error("backward: rotations / dL_drotations must be 16-byte aligned"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (!geom_ws || !image_ws || (R > 0 && (!binning_ws || !sample_ws))) { set_error("backward: missing saved workspace"); return GLIC_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    const ViewParams vp = make_view_params(view);
    GeomState g = GeomState::carve(const_cast<void*>(geom_ws), P);
    ImageState img = ImageState::carve(const_cast<void*>(image_ws), vp.W, vp.H);
    // 2-D gradient accumulators are filled by atomics: zero them here (11 floats / Gaussian)
    { StageTimer _t(GLIC_STAGE_ZERO, s);
    GLIC_CUDA_TRY(cudaMemsetAsync(dL_dmeans2D, 0, sizeof(float) * 3 * (size_t)P, s));
    GLIC_CUDA_TRY(cudaMemsetAsync(dL_dconic, 0, sizeof(float) * 4 * (size_t)P, s));
    GLIC_CUDA_TRY(cudaMemsetAsync(dL_dopacity, 0, sizeof(float) * (size_t)P, s));
    GLIC_CUDA_TRY(cudaMemsetAsync(dL_dcolors, 0, sizeof(float) * 3 * (size_t)P, s)); }
    if (R > 0) {
        BinningState bin = BinningState::carve(const_cast<void*>(binning_ws), R);
        const int64_t max_buckets = glic_max_buckets(R, vp.W, vp.H);
        SampleState smp = SampleState::carve(const_cast<void*>(sample_ws), max_buckets);
        // the sort's ping-pong parity is a pure function of the pass count
        const int T = vp.grid_x * vp.grid_y;
        const int passes = ((int)higher_msb((uint32_t)T) + 7) / 8;
        const int cur = passes & 1;
        StageTimer _t(GLIC_STAGE_RENDER_BWD, s);
        if (int e = launch_render_backward(P, vp, max_buckets, bin.vals[cur], g, img, smp, dL_dpix, dL_dmeans2D, dL_dconic,
                                           dL_dopacity, dL_dcolors, s)) return e;
    }
    StageTimer _t(GLIC_STAGE_PREPROCESS_BWD, s);
    return launch_preprocess_backward(P, sh_degree, M, means3D, scales, scale_modifier, rotations, sh, vp, radii, g,
                                      lambda_erank, dL_dmeans2D, dL_dconic, dL_dcolors, dL_dmeans3D, dL_dcov3D, dL_ddc,
                                      dL_dsh, dL_dscales, dL_drotations, s);
}

// Internal (mapper.cu): the backward with the compact gradient form of preprocess_backward.cu.  dL_dcolors is zeroed and
// accumulated like the other 2-D gradients but may live anywhere (the mapper points it into its exchange block).
int glic_backward_compact_internal(int P, int sh_degree, int M, const float* means3D, const float* scales, const float* rotations,
                                   const float* sh, const glic_view* view, const int* radii, int64_t R, const void* geom_ws,
                                   const void* binning_ws, const void* image_ws, const void* sample_ws, const float* dL_dpix,
                                   float* dL_dmeans2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolors,
                                   const glic::CompactGrads* cg, void* after_render_event, void* stream) {
    if (int e = check_view(view)) return e;
    if (P <= 0) return GLIC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const ViewParams vp = make_view_params(view);
    GeomState g = GeomState::carve(const_cast<void*>(geom_ws), P);
    ImageState img = ImageState::carve(const_cast<void*>(image_ws), vp.W, vp.H);
    { StageTimer _t(GLIC_STAGE_ZERO, s);
    // conic | mean2D | opacity carved from ONE block (the mapper's capacity layout): one memset over the span when the slack
    // between the live prefixes is small, three otherwise
    const ptrdiff_t span = dL_dconic < dL_dmeans2D && dL_dmeans2D < dL_dopacity ? (dL_dopacity + P) - dL_dconic : -1;
    if (span > 0 && span <= 10 * (ptrdiff_t)P) {
        GLIC_CUDA_TRY(cudaMemsetAsync(dL_dconic, 0, sizeof(float) * (size_t)span, s));
    } else {
        GLIC_CUDA_TRY(cudaMemsetAsync(dL_dmeans2D, 0, sizeof(float) * 3 * (size_t)P, s));
        GLIC_CUDA_TRY(cudaMemsetAsync(dL_dconic, 0, sizeof(float) * 4 * (size_t)P, s));
        GLIC_CUDA_TRY(cudaMemsetAsync(dL_dopacity, 0, sizeof(float) * (size_t)P, s));
    }
    GLIC_CUDA_TRY(cudaMemsetAsync(dL_dcolors, 0, sizeof(float) * 3 * (size_t)P, s)); }
    if (R > 0) {
        BinningState bin = BinningState::carve(const_cast<void*>(binning_ws), R);
        const int64_t max_buckets = glic_max_buckets(R, vp.W, vp.H);
        SampleState smp = SampleState::carve(const_cast<void*>(sample_ws), max_buckets);
        const int T = vp.grid_x * vp.grid_y;
        const int passes = ((int)higher_msb((uint32_t)T) + 7) / 8;
        const int cur = passes & 1;
        StageTimer _t(GLIC_STAGE_RENDER_BWD, s);
        if (int e = launch_render_backward(P, vp, max_buckets, bin.vals[cur], g, img, smp, dL_dpix, dL_dmeans2D, dL_dconic,
                                           dL_dopacity, dL_dcolors, s)) return e;
    }
    // the colour gradients are final here: the multi-GPU push may start while the per-Gaussian backward still runs
    if (after_render_event) GLIC_CUDA_TRY(cudaEventRecord((cudaEvent_t)after_render_event, s));
    StageTimer _t(GLIC_STAGE_PREPROCESS_BWD, s);
    return launch_preprocess_backward_compact(P, sh_degree, M, means3D, scales, 1.0f, rotations, sh, vp, radii, g, dL_dmeans2D,
                                              dL_dconic, dL_dcolors, *cg, s);
}

int glic_sort_pairs_u64_u32(int64_t n, int end_bit, uint64_t* keys_in, uint32_t* vals_in, uint64_t* keys_out,
                            uint32_t* vals_out, void* temp, size_t temp_bytes, void* stream) {
    if (n < 0 || (n > 0 && (!keys_in || !vals_in || !keys_out || !vals_out || !temp))) { set_error("sort_pairs: bad arguments"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (n == 0) return GLIC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    uint64_t* k[2] = {keys_in, keys_out};
    uint32_t* v[2] = {vals_in, vals_out};
    const int cur = launch_sort_pairs(n, end_bit, k, v, temp, temp_bytes, s);
    if (cur < 0) return cur;
    if (cur == 0) {   // even number of passes: result sits in the input buffers
        GLIC_CUDA_TRY(cudaMemcpyAsync(keys_out, keys_in, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToDevice, s));
        GLIC_CUDA_TRY(cudaMemcpyAsync(vals_out, vals_in, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    }
    return GLIC_OK;
}

int glic_debug_geom(int P, const void* geom_ws, float* depth, float* xy, float* conic_opacity, float* rgb,
                    uint32_t* tiles_touched, uint32_t* offsets, uint8_t* clamped, void* stream) {
    if (P <= 0) return GLIC_OK;
    if (!geom_ws) { set_error("debug_geom: null workspace"); return GLIC_ERR_WORKSPACE; }
    GeomState g = GeomState::carve(const_cast<void*>(geom_ws), P);
    debug_geom_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, g, depth, xy, conic_opacity, rgb, tiles_touched, offsets, clamped);
    GLIC_LAUNCH_CHECK();
    return GLIC_OK;
}

int glic_debug_binning(int P, const void* geom_ws, int64_t R, int64_t count, const void* binning_ws, uint32_t* point_list,
                       uint64_t* keys_sorted, void* stream) {
    if (R <= 0 || count <= 0) return GLIC_OK;
    if (count > R) { set_error("debug_binning: count exceeds the carve size"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (!binning_ws || !geom_ws) { set_error("debug_binning: null workspace"); return GLIC_ERR_WORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    BinningState bin = BinningState::carve(const_cast<void*>(binning_ws), R);
    GeomState g = GeomState::carve(const_cast<void*>(geom_ws), P);
    unsigned int cur = 0;
    GLIC_CUDA_TRY(cudaMemcpyAsync(&cur, &bin.hdr->sorted_in_b, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    GLIC_CUDA_TRY(cudaStreamSynchronize(s));
    cur &= 1u;
    if (point_list) GLIC_CUDA_TRY(cudaMemcpyAsync(point_list, bin.vals[cur], sizeof(uint32_t) * (size_t)count, cudaMemcpyDeviceToDevice, s));
    if (keys_sorted) {   // the 64-bit (tile|depth) key of the reference, rebuilt from the tile key and the record's depth
        debug_keys_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>((long long)count, g, bin.keys[cur], bin.vals[cur], keys_sorted);
        GLIC_LAUNCH_CHECK();
    }
    return GLIC_OK;
}

int glic_debug_image(int width, int height, const void* image_ws, uint32_t* ranges, uint32_t* bucket_offsets,
                     uint32_t* n_contrib, uint32_t* max_contrib, int64_t* counters2_host, void* stream) {
    if (!image_ws || width <= 0 || height <= 0) { set_error("debug_image: bad arguments"); return GLIC_ERR_INVALID_ARGUMENT; }
    cudaStream_t s = (cudaStream_t)stream;
    ImageState img = ImageState::carve(const_cast<void*>(image_ws), width, height);
    const size_t T = (size_t)((width + TILE - 1) / TILE) * ((height + TILE - 1) / TILE);
    if (ranges) GLIC_CUDA_TRY(cudaMemcpyAsync(ranges, img.ranges, sizeof(uint2) * T, cudaMemcpyDeviceToDevice, s));
    if (bucket_offsets) GLIC_CUDA_TRY(cudaMemcpyAsync(bucket_offsets, img.bucket_offsets, sizeof(uint32_t) * T, cudaMemcpyDeviceToDevice, s));
    if (n_contrib) GLIC_CUDA_TRY(cudaMemcpyAsync(n_contrib, img.n_contrib, sizeof(uint32_t) * (size_t)width * height, cudaMemcpyDeviceToDevice, s));
    if (max_contrib) GLIC_CUDA_TRY(cudaMemcpyAsync(max_contrib, img.max_contrib, sizeof(uint32_t) * T, cudaMemcpyDeviceToDevice, s));
    if (counters2_host) {
        ImageHeader h;
        GLIC_CUDA_TRY(cudaMemcpyAsync(&h, img.hdr, sizeof(ImageHeader), cudaMemcpyDeviceToHost, s));
        GLIC_CUDA_TRY(cudaStreamSynchronize(s));
        counters2_host[0] = h.num_rendered;
        counters2_host[1] = h.num_buckets;
    }
    return GLIC_OK;
}

}  // extern "C"
