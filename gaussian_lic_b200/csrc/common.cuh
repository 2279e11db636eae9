// common.cuh -- shared declarations of the glic_b200 CUDA library (sm_100a only).
//
// Workspace layouts (all opaque to callers; see DESIGN.md "HBM layout"):
//   geom_ws   : GeomHeader | rec[3P] float4 (48 B AoS splat record) | depth_keys[2][P] u32 + order[2][P] u32
//               (depth sort ping-pong) | offsets[P] u32 (inclusive scan of tiles_touched in DEPTH order) |
//               clamped[P] u8 | look-back status u64[blocks] | depth-sort temp
//   image_ws  : ImageHeader | ranges[T] uint2 | bucket_offsets[T] u32 | max_contrib[T] u32 |
//               n_contrib[HW] u32 | pixel_colors[3HW] f32
//   binning_ws: BinHeader | tile keys[2][R] u32 | vals[2][R] u32 | sort temp
//   sample_ws : bucket_to_tile[Bmax] u32 | ckpt[Bmax*256] float4 (T, C.r, C.g, C.b)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>

#include "../../include/glic_b200.h"

namespace glic {

constexpr int TILE = GLIC_TILE;
constexpr int TILE_PIX = TILE * TILE;   // 256 pixels, one CTA
constexpr int BUCKET = GLIC_BUCKET;     // 32 splats per checkpoint bucket
constexpr int PRE_THREADS = 256;        // Gaussians per preprocess CTA

// ---- error plumbing --------------------------------------------------------------------
void set_error(const std::string& s);
extern unsigned long long g_launches;

#define GLIC_CUDA_TRY(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            glic::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));             \
            return GLIC_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

#define GLIC_LAUNCH_CHECK()                                                                   \
    do {                                                                                      \
        ++glic::g_launches;                                                                   \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            glic::set_error(std::string("kernel launch: ") + cudaGetErrorString(_e));       \
            return GLIC_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

// ---- optional per-stage cudaEvent timing (glic_profile_*) ---------------------------------
bool profiling_on();
void profile_mark(int stage, bool begin, cudaStream_t s);
struct StageTimer {
    int stage; cudaStream_t s; bool on;
    StageTimer(int st, cudaStream_t stream) : stage(st), s(stream), on(profiling_on()) { if (on) profile_mark(stage, true, s); }
    ~StageTimer() { if (on) profile_mark(stage, false, s); }
};

// ---- arena carving (128-byte aligned sub-arrays) ----------------------------------------
struct Carver {
    char* base;
    size_t off;
    __host__ explicit Carver(void* p) : base(static_cast<char*>(p)), off(0) {}
    template <typename T>
    __host__ T* take(size_t count) {
        off = (off + 127) & ~size_t(127);
        T* p = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return p;
    }
    __host__ size_t total() const { return (off + 127) & ~size_t(127); }
};

// Splat record, 48 bytes: three float4 per Gaussian (AoS so one gather = 1.5 sectors).
//   r0 = (x, y, conic.x, conic.y)   r1 = (conic.z, opacity, red, green)
//   r2 = (blue, depth, radius as int bits, hy = conservative vertical half-extent of the alpha >= 1/255 ellipse)
struct GeomHeader {
    unsigned int ticket;        // dynamic CTA id of the emit kernel's fused scan
    unsigned int total;         // R = sum of tiles_touched
    unsigned int visible;       // #Gaussians with radius > 0
    unsigned int order_cur;     // which ping-pong half holds the depth-sorted order
    unsigned int r_eff;         // R when it fits the binning workspace, else 0: what the sort / ranges / render use
    unsigned int overflow;      // 1 when R exceeded the capacity (frame invalid, caller re-runs with more room)
    unsigned int span_count;    // bump allocator of the row-span arena (preprocess -> emit)
    unsigned int pad[25];
};

size_t sort_temp_bytes(int64_t n);

struct GeomState {
    GeomHeader* hdr;
    float4* rec;
    uint32_t* depth_keys[2];    // float bits of depth (0xFFFFFFFF for culled Gaussians), sort ping-pong
    uint32_t* order[2];         // Gaussian indices, sorted by (depth, index) after the depth sort
    uint32_t* tiles;            // tiles_touched per Gaussian (index order)
    uint32_t* offsets;          // per Gaussian: END of its slot range = inclusive scan of tiles over the depth order
    uint32_t* tmask;            // per Gaussian: accept mask of a <= 32-tile rect, or first row-span entry of a bigger one
    uint32_t* spans;            // row spans (xa | xb << 16) of the big rects, bump-allocated per frame
    uint32_t span_cap;
    uint8_t* clamped;
    unsigned long long* scan_status;
    void* sort_temp;
    size_t sort_temp_size;
    __host__ static GeomState carve(void* ws, int P, size_t* bytes = nullptr) {
        const size_t n = P > 0 ? size_t(P) : 1;
        Carver c(ws);
        GeomState g;
        g.hdr = c.take<GeomHeader>(1);
        g.rec = c.take<float4>(size_t(3) * n);
        g.depth_keys[0] = c.take<uint32_t>(n);
        g.depth_keys[1] = c.take<uint32_t>(n);
        g.order[0] = c.take<uint32_t>(n);
        g.order[1] = c.take<uint32_t>(n);
        g.tiles = c.take<uint32_t>(n);
        g.offsets = c.take<uint32_t>(n);
        g.tmask = c.take<uint32_t>(n);
        g.span_cap = (uint32_t)(4 * n);
        g.spans = c.take<uint32_t>(4 * n);
        g.clamped = c.take<uint8_t>(n);
        g.scan_status = c.take<unsigned long long>((n + PRE_THREADS - 1) / PRE_THREADS + 1);
        g.sort_temp_size = sort_temp_bytes((int64_t)n);
        g.sort_temp = c.take<char>(g.sort_temp_size);
        if (bytes) *bytes = c.total();
        return g;
    }
};

// Function attributes (opt-in dynamic shared memory) are per device: run `fn` once per (call site, device).
// `flags` is the call site's own static array.
inline bool first_use_on_device(bool (&flags)[64]) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (flags[dev]) return false;
    flags[dev] = true;
    return true;
}

struct ImageHeader {
    long long num_rendered;     // R
    unsigned int num_buckets;   // B
    unsigned int num_live_buckets;   // buckets some pixel reached (backward work list)
    unsigned int bwd_ticket;    // dynamic work counter of the backward's persistent warps (reset by live_scan_kernel)
    long long counters[4];      // {R, B, overflow, 0}: copied to the host asynchronously by glic_forward
    unsigned int pad[20];
};

struct ImageState {
    ImageHeader* hdr;
    uint2* ranges;
    uint32_t* bucket_offsets;
    uint32_t* max_contrib;
    uint32_t* live_offsets;     // inclusive scan of ceil(max_contrib/32) (backward's live-bucket list)
    uint32_t* n_contrib;
    float* pixel_colors;
    __host__ static ImageState carve(void* ws, int W, int H, size_t* bytes = nullptr) {
        const size_t T = size_t((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
        const size_t HW = size_t(W) * H;
        Carver c(ws);
        ImageState s;
        s.hdr = c.take<ImageHeader>(1);
        s.ranges = c.take<uint2>(T);
        s.bucket_offsets = c.take<uint32_t>(T);
        s.max_contrib = c.take<uint32_t>(T);
        s.live_offsets = c.take<uint32_t>(T);
        s.n_contrib = c.take<uint32_t>(HW);
        s.pixel_colors = c.take<float>(3 * HW);
        if (bytes) *bytes = c.total();
        return s;
    }
};

struct BinHeader {
    unsigned int sorted_in_b;   // 1 when the sorted list lives in keys[1]/vals[1]
    unsigned int pad[31];
};

struct BinningState {
    BinHeader* hdr;
    uint32_t* keys[2];          // tile ids
    uint32_t* vals[2];
    void* sort_temp;
    size_t sort_temp_size;
    __host__ static BinningState carve(void* ws, int64_t R, size_t* bytes = nullptr) {
        const size_t n = R > 0 ? size_t(R) : 1;
        Carver c(ws);
        BinningState b;
        b.hdr = c.take<BinHeader>(1);
        b.keys[0] = c.take<uint32_t>(n);
        b.keys[1] = c.take<uint32_t>(n);
        b.vals[0] = c.take<uint32_t>(n);
        b.vals[1] = c.take<uint32_t>(n);
        b.sort_temp_size = sort_temp_bytes(R);
        b.sort_temp = c.take<char>(b.sort_temp_size);
        if (bytes) *bytes = c.total();
        return b;
    }
};

struct SampleState {
    uint32_t* bucket_to_tile;
    float4* ckpt;               // [Bmax][256] (T, C.r, C.g, C.b) per pixel per bucket
    __host__ static SampleState carve(void* ws, int64_t max_buckets, size_t* bytes = nullptr) {
        const size_t n = max_buckets > 0 ? size_t(max_buckets) : 1;
        Carver c(ws);
        SampleState s;
        s.bucket_to_tile = c.take<uint32_t>(n);
        s.ckpt = c.take<float4>(n * TILE_PIX);
        if (bytes) *bytes = c.total();
        return s;
    }
};

// getHigherMsb (rasterizer_impl.cu:42-57): number of tile-id bits included in the sort.
inline uint32_t higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

// View parameters passed by value to kernels (pointers stay device pointers).
struct ViewParams {
    const float* view;
    const float* proj;
    const float* campos;
    float tan_fovx, tan_fovy, focal_x, focal_y;
    float limx_neg, limx_pos, limy_neg, limy_pos;
    int W, H, grid_x, grid_y;
};

inline ViewParams make_view_params(const glic_view* v) {
    ViewParams p;
    p.view = v->viewmatrix; p.proj = v->projmatrix; p.campos = v->campos;
    p.tan_fovx = v->tan_fovx; p.tan_fovy = v->tan_fovy;
    p.focal_y = v->height / (2.0f * v->tan_fovy);     // rasterizer_impl.cu:348-349
    p.focal_x = v->width / (2.0f * v->tan_fovx);
    p.limx_neg = v->limx_neg; p.limx_pos = v->limx_pos; p.limy_neg = v->limy_neg; p.limy_pos = v->limy_pos;
    p.W = v->width; p.H = v->height;
    p.grid_x = (v->width + TILE - 1) / TILE; p.grid_y = (v->height + TILE - 1) / TILE;
    return p;
}

// Compact gradient outputs of the per-Gaussian backward (the mapper's form, see preprocess_backward.cu).
struct CompactGrads {
    float4* g_rot;            // [P] dL/d(raw rotation)
    float* g_xyz;             // [P,3]
    float* g_scale;           // [P,3] dL/d(log-scale)
    float* g_opacity;         // [P]   dL/d(opacity logit)
    const float* opacity;     // [P]   activated opacity (sigmoid output)
    const float* dL_dopacity; // [P]   dL/d(activated opacity) from the render backward
    const float4* rot_raw;    // [P]   raw (un-normalised) rotation
    int accumulate;           // 0: overwrite the geometric gradients, 1: add to them (second and later views of a rank)
};

// ---- stage launchers (defined in the .cu files) ------------------------------------------
int launch_preprocess_forward(int P, int D, int M, const float* means, const float* scales, float mod,
                              const float* rots, const float* opac, const float* dc, const float* sh,
                              const ViewParams& vp, bool no_color, int* radii, GeomState g, cudaStream_t s);
int launch_depth_scan(int P, GeomState g, const uint32_t* order, int64_t capacity, cudaStream_t s);
int launch_emit_keys(int P, const ViewParams& vp, GeomState g, uint32_t* tile_keys, uint32_t* vals, int64_t capacity, uint32_t* sort_hist_out,
                     int end_bit, cudaStream_t s);
// Sorts on bits [0,end_bit); returns 0/1 = index of the ping-pong buffer holding the result, <0 on error.
int launch_sort_pairs(int64_t n, int end_bit, uint64_t* keys[2], uint32_t* vals[2], void* temp, size_t temp_bytes,
                      cudaStream_t s);
// n_dev (optional): device-side true count <= n; kernels are launched for n (a capacity) and clip to *n_dev.
int launch_sort_pairs32(int64_t n, int end_bit, uint32_t* keys[2], uint32_t* vals[2], void* temp, size_t temp_bytes,
                        cudaStream_t s, const unsigned int* n_dev = nullptr, bool hist_ready = false);
// A producer kernel may build the digit histograms itself: sort_prepare() zeroes the temp block, the producer adds one count
// per key and pass into sort_hist(temp)[pass * 256 + digit], and the sort is launched with hist_ready = true.
int sort_prepare(int64_t n, int end_bit, void* temp, cudaStream_t s, int digit_bits = 8);
uint32_t* sort_hist(void* temp);
// R: host-side count or capacity; r_dev (optional): device-side true count
int launch_tile_ranges(int64_t R, const GeomHeader* ghdr, const uint32_t* tile_keys_sorted, int T, ImageState img, bool buckets,
                       cudaStream_t s);
int launch_render_forward(const ViewParams& vp, bool no_color, const uint32_t* point_list, GeomState g, ImageState img,
                          SampleState smp, float* out_color, float* out_final_T, cudaStream_t s);
int launch_render_backward(int P, const ViewParams& vp, int64_t max_buckets, const uint32_t* point_list, GeomState g,
                           ImageState img, SampleState smp, const float* dL_dpix, float* dL_dmean2D /*[P,3]*/,
                           float* dL_dconic /*[P,4]*/, float* dL_dopacity, float* dL_dcolors, cudaStream_t s);
int launch_preprocess_backward(int P, int D, int M, const float* means, const float* scales, float mod,
                               const float* rots, const float* sh, const ViewParams& vp, const int* radii, GeomState g,
                               float lambda_erank, const float* dL_dmean2D, const float* dL_dconic,
                               const float* dL_dcolors, float* dL_dmeans3D, float* dL_dcov3D, float* dL_ddc,
                               float* dL_dsh, float* dL_dscales, float* dL_drots, cudaStream_t s);

int launch_preprocess_backward_compact(int P, int D, int M, const float* means, const float* scales, float mod,
                                       const float* rots, const float* sh, const ViewParams& vp, const int* radii, GeomState g,
                                       const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolors,
                                       const CompactGrads& cg, cudaStream_t s);

}  // namespace glic

// internal cross-file entry points (not part of the C ABI)
extern "C" int glic_backward_compact_internal(int P, int sh_degree, int M, const float* means3D, const float* scales,
                                              const float* rotations, const float* sh, const glic_view* view, const int* radii,
                                              int64_t R, const void* geom_ws, const void* binning_ws, const void* image_ws,
                                              const void* sample_ws, const float* dL_dpix, float* dL_dmeans2D, float* dL_dconic,
                                              float* dL_dopacity, float* dL_dcolors, const glic::CompactGrads* cg,
                                              void* after_render_event, void* stream);
namespace glic {
int launch_adam_compact(uint32_t P, uint32_t Pcap, int D, int M, float* params, float* exp_avg, float* exp_avg_sq, const float* g_geo,
                        uint32_t Pcap_geo, const float* lr6, const float* g_color, const uint8_t* flags, const float* campos4,
                        int n_slots, float grad_scale, float color_scale, float b1, float b2, float eps, const unsigned int* skip_flag,
                        unsigned int* visible_count, cudaStream_t s);
int launch_view_flags(int P, const int* radii, const GeomState& g, uint8_t* flags, cudaStream_t s);
}  // namespace glic
