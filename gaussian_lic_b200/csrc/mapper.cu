// mapper.cu -- the native mapping host (include/glic_b200.h "Native mapping host").
//
// C++ counterpart of the reference's GaussianModel + extend() + optimize() + evaluateVisualQuality() + saveMap()
// (/root/reference/src/gaussian.cpp:113-830, camera.h:38-110) on top of this library's kernels: one object per GPU, plain
// CUDA runtime, no torch, no Python.  What is different from the reference loop, by design:
//   * the model, its Adam moments and every per-Gaussian buffer live in capacity-sized arenas: extend() appends in place
//     (densificationPostfix re-allocates and copies all six parameter tensors and both moment tensors per keyframe);
//   * an iteration enqueues ~25 asynchronous launches and never blocks the host (the reference synchronises 5x per
//     iteration, gaussian.cpp:679,692,698,704,713); the pinned keyframe image travels on a copy stream into a
//     double buffer; binning overflow is detected with a lag through pinned counters and skipped on the device;
//   * gradients never exist as [P,59]: see the header comment of preprocess_backward.cu / adam_compact_kernel;
//   * multi-GPU (view-sharded): per-view colour gradients + flag bytes are PUSHED to every peer's exchange block by the
//     copy engines over NVLink as soon as the render backward has produced them (overlapping the per-Gaussian backward),
//     the 11 geometric floats are mean-reduced in place by the two-shot kernel of p2p.cu; every replica then runs the same
//     Adam launch on the same bytes.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>
#include <vector>

using namespace glic;

namespace {

constexpr int MAX_SLOTS = 32;        // world * views_per_rank
constexpr int RING = 8;              // pinned result ring (loss, binning counters)

struct CamBlock {                    // what the rasterizer gets per view (camera.h:38-110, renderer.cpp:31-54)
    float view[16], proj[16], campos[4];
    float tanfovx, tanfovy, lim[4];
};

// camera.h:38-110 in closed form: FoV from the intrinsics (float members), world->view matrix Rt, projection P, both handed
// to the kernels transposed (column-major), camera centre = -R_cw^T t_cw, the asymmetric clamp limits of :63-66.
void camera_block(int W, int H, double fx, double fy, double cx, double cy, const float* R_wc, const float* t_wc, CamBlock& o) {
    double Rcw[9], tcw[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rcw[3 * i + j] = (double)R_wc[3 * j + i];
    for (int i = 0; i < 3; ++i) tcw[i] = -(Rcw[3 * i] * t_wc[0] + Rcw[3 * i + 1] * t_wc[1] + Rcw[3 * i + 2] * t_wc[2]);
    float Rt[16] = {0};
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) Rt[4 * i + j] = (float)Rcw[3 * i + j]; Rt[4 * i + 3] = (float)tcw[i]; }
    Rt[15] = 1.0f;
    const float fovx = (float)(2.0 * std::atan(W / (2.0 * fx))), fovy = (float)(2.0 * std::atan(H / (2.0 * fy)));
    const float znear = 0.01f, zfar = 100.0f;
    float Pm[16] = {0};
    Pm[0] = (float)(1.0 / std::tan((double)fovx / 2));
    Pm[5] = (float)(1.0 / std::tan((double)fovy / 2));
    Pm[2] = (2.0f * (float)cx - (float)W) / (float)W;
    Pm[6] = (2.0f * (float)cy - (float)H) / (float)H;
    Pm[14] = 1.0f;
    Pm[10] = zfar / (zfar - znear);
    Pm[11] = -(zfar * znear) / (zfar - znear);
    float full[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float acc = 0.f;
            for (int k = 0; k < 4; ++k) acc += Pm[4 * i + k] * Rt[4 * k + j];
            full[4 * i + j] = acc;
        }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) { o.view[4 * c + r] = Rt[4 * r + c]; o.proj[4 * c + r] = full[4 * r + c]; }
    for (int i = 0; i < 3; ++i) o.campos[i] = (float)(-(Rcw[i] * tcw[0] + Rcw[3 + i] * tcw[1] + Rcw[6 + i] * tcw[2]));
    o.campos[3] = 0.f;
    o.tanfovx = std::tan(fovx * 0.5f);
    o.tanfovy = std::tan(fovy * 0.5f);
    const float ffx = (float)fx, ffy = (float)fy, fcx = (float)cx, fcy = (float)cy;
    o.lim[0] = -0.15f * W / ffx - fcx / ffx; o.lim[1] = 1.15f * W / ffx - fcx / ffx;
    o.lim[2] = -0.15f * H / ffy - fcy / ffy; o.lim[3] = 1.15f * H / ffy - fcy / ffy;
}

struct Keyframe {
    float R_wc[9], t_wc[3];
    const float* image;
    CamBlock cam;
    float* cam_dev;                  // view16 | proj16 | campos4
    glic_view view;
};

inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

}  // namespace

struct glic_mapper {
    glic_mapper_config cfg{};
    int dev = 0, D = 0, M = 0, S = 1, k = 1;
    cudaStream_t stream = nullptr, copy_stream = nullptr, push_stream = nullptr;
    // ---- arenas (capacity Pcap) ----
    uint32_t P = 0, Pcap = 0;
    float *params = nullptr, *m1 = nullptr, *m2 = nullptr;
    float *act_opacity = nullptr, *act_scales = nullptr, *act_rots = nullptr;
    int* radii = nullptr;
    float *g_mean2D = nullptr, *g_conic = nullptr, *g_opacity = nullptr;
    void* geom_ws = nullptr; size_t geom_bytes = 0;
    // ---- exchange block: geo[11*Pcap + 4] | flags words | colour slots [2][S][Pcap][3] | flag bytes [2][S][Pcap] ----
    char* xblock = nullptr; size_t xbytes = 0, off_col = 0, off_flags = 0;
    size_t n_geo = 0;                // floats in the reduced region (11*Pcap + 4; the tail holds the collective overflow word)
    void* peers[8] = {nullptr};
    bool connected = false;
    std::vector<void*> opened;
    // ---- per-image buffers ----
    void *image_ws = nullptr, *binning_ws = nullptr, *sample_ws = nullptr, *loss_scratch = nullptr, *eval_scratch = nullptr, *extend_ws = nullptr;
    size_t image_bytes = 0, binning_bytes = 0, sample_bytes = 0, loss_bytes = 0, eval_bytes = 0, extend_bytes = 0;
    int64_t bin_cap = 0;
    float *color = nullptr, *final_T = nullptr, *dL_dpix = nullptr, *gt_dev[2] = {nullptr, nullptr}, *loss_dev = nullptr, *eval_out = nullptr;
    unsigned int* vis_acc = nullptr;
    cudaEvent_t gt_free[2]{}, gt_ready[2]{}, ev_flags{}, ev_col{}, ev_push{}, ring_ev[RING]{}, t0{}, t1{};
    long long* counters_host = nullptr;   // pinned [RING][4]
    float* loss_host = nullptr;           // pinned [RING]
    bool ring_used[RING] = {false};
    // ---- keyframes ----
    std::vector<Keyframe> train, test;
    std::mt19937_64 rng;
    uint64_t view_counter = 0;
    glic_mapper_stats st{};
    bool run_optimizer = true;            // GLIC_MAPPER_OPT_OPTIMIZER (benchmarks time the rasterization step alone with it off)
};

namespace {

#define MAP_TRY(expr) do { int _e = (expr); if (_e != GLIC_OK) return _e; } while (0)

int dev_alloc(void** p, size_t bytes, bool zero = true) {
    GLIC_CUDA_TRY(cudaMalloc(p, bytes ? bytes : 256));
    if (zero) GLIC_CUDA_TRY(cudaMemset(*p, 0, bytes ? bytes : 256));
    return GLIC_OK;
}

size_t arena_floats(const glic_mapper* m, uint32_t cap) { return glic_packed_floats(cap, (uint32_t)m->M); }

void arena_ptrs(const glic_mapper* m, float* base, uint32_t cap, float* out6[6]) {
    size_t off[6];
    glic_packed_offsets(cap, (uint32_t)m->M, off);
    for (int q = 0; q < 6; ++q) out6[q] = base + off[q];
}

float* geo_ptr(const glic_mapper* m) { return reinterpret_cast<float*>(m->xblock); }
float* col_slot(const glic_mapper* m, int parity, int slot) {
    return reinterpret_cast<float*>(m->xblock + m->off_col) + ((size_t)parity * m->S + slot) * (size_t)m->Pcap * 3;
}
uint8_t* flag_slot(const glic_mapper* m, int parity, int slot) {
    return reinterpret_cast<uint8_t*>(m->xblock + m->off_flags) + ((size_t)parity * m->S + slot) * (size_t)m->Pcap;
}

size_t xblock_layout(glic_mapper* m, uint32_t cap) {
    m->n_geo = (size_t)11 * cap + 4;
    size_t sl[6];
    glic_p2p_slice(0, 1, m->n_geo, 0, sl);
    const size_t flag_off = sl[5];
    m->off_col = flag_off + 256;
    m->off_flags = m->off_col + (((size_t)2 * m->S * cap * 3 * sizeof(float) + 255) & ~size_t(255));
    return m->off_flags + (((size_t)2 * m->S * cap + 255) & ~size_t(255));
}

// (re)allocates everything whose size depends on the Gaussian capacity; copies the live rows of the arenas
int set_capacity(glic_mapper* m, uint32_t cap) {
    cap = round_up(std::max<uint32_t>(cap, 256), 256);
    if (cap <= m->Pcap) return GLIC_OK;
    if (m->connected) { set_error("mapper: the arena capacity is fixed once the exchange blocks are connected (world > 1); create the mapper with a larger capacity"); return GLIC_ERR_WORKSPACE; }
    GLIC_CUDA_TRY(cudaStreamSynchronize(m->stream));
    const size_t nfl = arena_floats(m, cap);
    float *np = nullptr, *n1 = nullptr, *n2 = nullptr;
    MAP_TRY(dev_alloc((void**)&np, nfl * 4)); MAP_TRY(dev_alloc((void**)&n1, nfl * 4)); MAP_TRY(dev_alloc((void**)&n2, nfl * 4));
    if (m->P) {
        MAP_TRY(glic_arena_regrow(m->params, m->Pcap, np, cap, (uint32_t)m->M, m->P, m->stream));
        MAP_TRY(glic_arena_regrow(m->m1, m->Pcap, n1, cap, (uint32_t)m->M, m->P, m->stream));
        MAP_TRY(glic_arena_regrow(m->m2, m->Pcap, n2, cap, (uint32_t)m->M, m->P, m->stream));
        GLIC_CUDA_TRY(cudaStreamSynchronize(m->stream));
    }
    void* old[] = {m->params, m->m1, m->m2, m->act_opacity, m->act_scales, m->act_rots, m->radii, m->g_conic, m->geom_ws, m->xblock};
    for (void* p : old) if (p) cudaFree(p);
    m->params = np; m->m1 = n1; m->m2 = n2;
    MAP_TRY(dev_alloc((void**)&m->act_opacity, (size_t)cap * 4)); MAP_TRY(dev_alloc((void**)&m->act_scales, (size_t)cap * 12));
    MAP_TRY(dev_alloc((void**)&m->act_rots, (size_t)cap * 16)); MAP_TRY(dev_alloc((void**)&m->radii, (size_t)cap * 4));
    // the three 2-D gradient accumulators the render backward adds into: one block (conic first: float4-aligned)
    MAP_TRY(dev_alloc((void**)&m->g_conic, (size_t)cap * (16 + 12 + 4)));
    m->g_mean2D = m->g_conic + (size_t)cap * 4;
    m->g_opacity = m->g_mean2D + (size_t)cap * 3;
    m->geom_bytes = glic_geom_bytes((int)cap);
    MAP_TRY(dev_alloc(&m->geom_ws, m->geom_bytes));
    m->xbytes = xblock_layout(m, cap);
    MAP_TRY(dev_alloc((void**)&m->xblock, m->xbytes));
    m->peers[m->cfg.rank] = m->xblock;
    m->Pcap = cap;
    m->st.capacity = cap;
    return GLIC_OK;
}

int set_bin_capacity(glic_mapper* m, int64_t pairs, bool exact = false) {
    if (!exact && pairs <= m->bin_cap) return GLIC_OK;
    GLIC_CUDA_TRY(cudaStreamSynchronize(m->stream));
    if (m->binning_ws) cudaFree(m->binning_ws);
    if (m->sample_ws) cudaFree(m->sample_ws);
    m->binning_bytes = glic_binning_bytes(pairs);
    m->sample_bytes = glic_sample_bytes(pairs, m->cfg.width, m->cfg.height);
    MAP_TRY(dev_alloc(&m->binning_ws, m->binning_bytes, false));
    MAP_TRY(dev_alloc(&m->sample_ws, m->sample_bytes, false));
    m->bin_cap = glic_binning_capacity(m->binning_bytes, m->sample_bytes, m->cfg.width, m->cfg.height, 0);
    return GLIC_OK;
}

// lagged, non-blocking look at the pinned binning counters of finished iterations
int poll_ring(glic_mapper* m, bool wait) {
    for (int r = 0; r < RING; ++r) {
        if (!m->ring_used[r]) continue;
        cudaError_t q = wait ? cudaEventSynchronize(m->ring_ev[r]) : cudaEventQuery(m->ring_ev[r]);
        if (q == cudaErrorNotReady) { cudaGetLastError(); continue; }
        if (q != cudaSuccess) { set_error(std::string("mapper: ") + cudaGetErrorString(q)); return GLIC_ERR_CUDA; }
        m->ring_used[r] = false;
        const long long* c = m->counters_host + 4 * r;
        m->st.last_loss = m->loss_host[r];
        if (c[2]) {                                        // this frame overflowed the binning workspace: it was skipped on the device
            ++m->st.overflow_regrows;
            MAP_TRY(set_bin_capacity(m, std::max<int64_t>(c[0] + c[0] / 4, m->bin_cap * 3 / 2)));
        }
    }
    return GLIC_OK;
}

int activate(glic_mapper* m) {
    float* g[6];
    arena_ptrs(m, m->params, m->Pcap, g);
    return glic_activations_forward((int)m->P, g[3], g[2], g[0], m->act_opacity, m->act_scales, m->act_rots, m->stream);
}

int forward_view(glic_mapper* m, const Keyframe& kf, int no_color, long long* counters) {
    float* g[6];
    arena_ptrs(m, m->params, m->Pcap, g);
    return glic_forward((int)m->P, m->D, m->M, g[1], m->act_scales, 1.0f, m->act_rots, m->act_opacity, g[4], m->M ? g[5] : nullptr, &kf.view,
                        no_color, m->radii, m->geom_ws, m->geom_bytes, m->image_ws, m->image_bytes, m->binning_ws, m->binning_bytes,
                        m->sample_ws, m->sample_bytes, m->color, m->final_T, (int64_t*)counters, m->stream);
}

__global__ void overflow_word_kernel(const GeomHeader* ghdr, float* word) { if (ghdr->overflow) *word = 1.0f; }

// one optimiser step over the view batch `views` (S keyframe indices, slot-major); this rank renders slots [rank*k, rank*k + k)
int iteration(glic_mapper* m, const int* views) {
    if (m->P == 0) return GLIC_OK;
    if (m->cfg.world > 1 && !m->connected) { set_error("mapper: world > 1 needs glic_mapper_connect before the first iteration"); return GLIC_ERR_INVALID_ARGUMENT; }
    for (int sl = 0; sl < m->S; ++sl)
        if (views[sl] < 0 || views[sl] >= (int)m->train.size()) { set_error("mapper: view index out of range"); return GLIC_ERR_INVALID_ARGUMENT; }
    MAP_TRY(poll_ring(m, false));
    const int r = (int)(m->st.iterations % RING);
    if (m->ring_used[r]) { GLIC_CUDA_TRY(cudaEventSynchronize(m->ring_ev[r])); MAP_TRY(poll_ring(m, false)); }
    cudaStream_t s = m->stream;
    const int parity = (int)(m->st.iterations & 1);
    const int W = m->cfg.width, H = m->cfg.height;
    float* g[6];
    arena_ptrs(m, m->params, m->Pcap, g);
    float* geo = geo_ptr(m);
    float* geo_tail = geo + (size_t)11 * m->Pcap;
    GLIC_CUDA_TRY(cudaMemsetAsync(geo_tail, 0, 4 * sizeof(float), s));
    GLIC_CUDA_TRY(cudaMemsetAsync(m->loss_dev, 0, sizeof(float) * 2, s));
    MAP_TRY(activate(m));
    size_t goff[6];
    glic_packed_offsets(m->Pcap, 0, goff);
    for (int j = 0; j < m->k; ++j) {
        const int slot = m->cfg.rank * m->k + j;
        const Keyframe& kf = m->train[views[slot]];
        const int buf = (int)(m->view_counter++ & 1);
        // keyframe image H2D on the copy stream (gaussian.cpp:678), overlapping the previous view's backward / Adam
        GLIC_CUDA_TRY(cudaStreamWaitEvent(m->copy_stream, m->gt_free[buf], 0));
        GLIC_CUDA_TRY(cudaMemcpyAsync(m->gt_dev[buf], kf.image, sizeof(float) * 3 * (size_t)W * H, cudaMemcpyDefault, m->copy_stream));
        GLIC_CUDA_TRY(cudaEventRecord(m->gt_ready[buf], m->copy_stream));
        MAP_TRY(forward_view(m, kf, 0, j == m->k - 1 ? m->counters_host + 4 * r : nullptr));
        uint8_t* fl = flag_slot(m, parity, slot);
        float* col = col_slot(m, parity, slot);
        MAP_TRY(launch_view_flags((int)m->P, m->radii, GeomState::carve(m->geom_ws, (int)m->P), fl, s));
        overflow_word_kernel<<<1, 1, 0, s>>>(GeomState::carve(m->geom_ws, (int)m->P).hdr, geo_tail);
        GLIC_LAUNCH_CHECK();
        if (m->cfg.world > 1) {
            GLIC_CUDA_TRY(cudaEventRecord(m->ev_flags, s));
            GLIC_CUDA_TRY(cudaStreamWaitEvent(m->push_stream, m->ev_flags, 0));
            for (int q = 0; q < m->cfg.world; ++q) {
                if (q == m->cfg.rank) continue;
                uint8_t* dst = static_cast<uint8_t*>(m->peers[q]) + (fl - reinterpret_cast<uint8_t*>(m->xblock));
                GLIC_CUDA_TRY(cudaMemcpyAsync(dst, fl, m->P, cudaMemcpyDeviceToDevice, m->push_stream));
            }
        }
        GLIC_CUDA_TRY(cudaStreamWaitEvent(s, m->gt_ready[buf], 0));
        // loss of this view (the pinned ring keeps the last local view's value for the statistics)
        MAP_TRY(glic_l1_ssim_loss(3, H, W, m->cfg.lambda_dssim, m->color, m->gt_dev[buf], m->loss_dev + 1, m->dL_dpix, m->loss_scratch, m->loss_bytes, s));
        GLIC_CUDA_TRY(cudaEventRecord(m->gt_free[buf], s));
        CompactGrads cg;
        cg.g_rot = reinterpret_cast<float4*>(geo + goff[0]); cg.g_xyz = geo + goff[1]; cg.g_scale = geo + goff[2]; cg.g_opacity = geo + goff[3];
        cg.opacity = m->act_opacity; cg.dL_dopacity = m->g_opacity; cg.rot_raw = reinterpret_cast<const float4*>(g[0]);
        cg.accumulate = j > 0;
        MAP_TRY(glic_backward_compact_internal((int)m->P, m->D, m->M, g[1], m->act_scales, m->act_rots, m->M ? g[5] : nullptr, &kf.view, m->radii,
                                               m->bin_cap, m->geom_ws, m->binning_ws, m->image_ws, m->sample_ws, m->dL_dpix, m->g_mean2D,
                                               m->g_conic, m->g_opacity, col, &cg, m->cfg.world > 1 ? (void*)m->ev_col : nullptr, s));
        if (m->cfg.world > 1) {
            GLIC_CUDA_TRY(cudaStreamWaitEvent(m->push_stream, m->ev_col, 0));
            for (int q = 0; q < m->cfg.world; ++q) {
                if (q == m->cfg.rank) continue;
                char* dst = static_cast<char*>(m->peers[q]) + (reinterpret_cast<char*>(col) - m->xblock);
                GLIC_CUDA_TRY(cudaMemcpyAsync(dst, col, sizeof(float) * 3 * (size_t)m->P, cudaMemcpyDeviceToDevice, m->push_stream));
            }
        }
    }
    if (m->cfg.world > 1) {
        GLIC_CUDA_TRY(cudaEventRecord(m->ev_push, m->push_stream));
        GLIC_CUDA_TRY(cudaStreamWaitEvent(s, m->ev_push, 0));
        MAP_TRY(glic_p2p_allreduce_mean(m->cfg.rank, m->cfg.world, m->peers, m->n_geo, 0, s));
    }
    float lr6[6] = {m->cfg.rotation_lr, m->cfg.position_lr, m->cfg.scaling_lr, m->cfg.opacity_lr, m->cfg.feature_lr, m->cfg.feature_lr / 20.0f};
    // Reference quirk kept for parity (backward.cu:352 `if (shs)`): with no SH-rest tensor bound (degree 0, M = 0) the whole
    // colour backward is skipped, dL/ddc included, so dc never trains.
    const float color_scale = m->M ? 1.0f / (float)m->S : 0.0f;
    float campos[MAX_SLOTS * 4];
    for (int sl = 0; sl < m->S; ++sl) std::memcpy(campos + 4 * sl, m->train[views[sl]].cam.campos, 16);
    if (m->run_optimizer)
    MAP_TRY(launch_adam_compact(m->P, m->Pcap, m->D, m->M, m->params, m->m1, m->m2, geo, m->Pcap, lr6, col_slot(m, parity, 0), flag_slot(m, parity, 0),
                                campos, m->S, 1.0f / (float)m->k, color_scale, 0.9f, 0.999f, 1e-15f, reinterpret_cast<const unsigned int*>(geo_tail), m->vis_acc, s));
    GLIC_CUDA_TRY(cudaMemcpyAsync(m->loss_host + r, m->loss_dev + 1, sizeof(float), cudaMemcpyDeviceToHost, s));
    GLIC_CUDA_TRY(cudaEventRecord(m->ring_ev[r], s));
    m->ring_used[r] = true;
    ++m->st.iterations;
    return GLIC_OK;
}

}  // namespace

extern "C" {

int glic_camera_block(int width, int height, float fx, float fy, float cx, float cy, const float* R_wc, const float* t_wc, float* out41) {
    if (!R_wc || !t_wc || !out41 || width <= 0 || height <= 0) { set_error("camera_block: bad arguments"); return GLIC_ERR_INVALID_ARGUMENT; }
    CamBlock c;
    camera_block(width, height, fx, fy, cx, cy, R_wc, t_wc, c);
    std::memcpy(out41, c.view, 64); std::memcpy(out41 + 16, c.proj, 64); std::memcpy(out41 + 32, c.campos, 12);
    out41[35] = c.tanfovx; out41[36] = c.tanfovy; std::memcpy(out41 + 37, c.lim, 16);
    return GLIC_OK;
}

int glic_mapper_create(const glic_mapper_config* cfg, glic_mapper** out) {
    if (!cfg || !out) { set_error("mapper_create: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (cfg->width <= 0 || cfg->height <= 0 || cfg->sh_degree < 0 || cfg->sh_degree > 3 || cfg->world < 1 || cfg->world > 8 || cfg->rank < 0 ||
        cfg->rank >= cfg->world || cfg->views_per_rank < 1 || cfg->world * cfg->views_per_rank > MAX_SLOTS) {
        set_error("mapper_create: bad configuration"); return GLIC_ERR_INVALID_ARGUMENT;
    }
    glic_mapper* m = new glic_mapper();
    m->cfg = *cfg;
    if (m->cfg.max_iters <= 0) m->cfg.max_iters = 100;
    m->D = cfg->sh_degree; m->M = (m->D + 1) * (m->D + 1) - 1;
    m->k = cfg->views_per_rank; m->S = cfg->world * cfg->views_per_rank;
    m->rng.seed(cfg->seed);
    auto fail = [&](int e) { glic_mapper_destroy(m); return e; };
    if (cudaGetDevice(&m->dev) != cudaSuccess) { set_error("mapper_create: no CUDA device"); return fail(GLIC_ERR_NO_DEVICE); }
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&m->push_stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("mapper_create: stream creation failed"); return fail(GLIC_ERR_CUDA); }
    cudaEvent_t* evs[] = {&m->gt_free[0], &m->gt_free[1], &m->gt_ready[0], &m->gt_ready[1], &m->ev_flags, &m->ev_col, &m->ev_push};
    for (cudaEvent_t* e : evs) if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) { set_error("mapper_create: event creation failed"); return fail(GLIC_ERR_CUDA); }
    for (int r = 0; r < RING; ++r) if (cudaEventCreateWithFlags(&m->ring_ev[r], cudaEventDisableTiming) != cudaSuccess) return fail(GLIC_ERR_CUDA);
    if (cudaEventCreate(&m->t0) != cudaSuccess || cudaEventCreate(&m->t1) != cudaSuccess) return fail(GLIC_ERR_CUDA);
    const int W = cfg->width, H = cfg->height;
    const size_t HW = (size_t)W * H;
    if (int e = set_capacity(m, cfg->capacity ? cfg->capacity : 1u << 20)) return fail(e);
    m->image_bytes = glic_image_bytes(W, H);
    m->loss_bytes = glic_loss_scratch_bytes(3, H, W);
    m->eval_bytes = glic_eval_scratch_bytes(3, H, W);
    int e = GLIC_OK;
    if ((e = dev_alloc(&m->image_ws, m->image_bytes)) || (e = dev_alloc(&m->loss_scratch, m->loss_bytes, false)) || (e = dev_alloc(&m->eval_scratch, m->eval_bytes, false)) ||
        (e = dev_alloc((void**)&m->color, HW * 12)) || (e = dev_alloc((void**)&m->final_T, HW * 4)) || (e = dev_alloc((void**)&m->dL_dpix, HW * 12)) ||
        (e = dev_alloc((void**)&m->gt_dev[0], HW * 12)) || (e = dev_alloc((void**)&m->gt_dev[1], HW * 12)) || (e = dev_alloc((void**)&m->loss_dev, 64)) ||
        (e = dev_alloc((void**)&m->eval_out, 64)) || (e = dev_alloc((void**)&m->vis_acc, 64))) return fail(e);
    if (cudaHostAlloc((void**)&m->counters_host, sizeof(long long) * 4 * RING, cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void**)&m->loss_host, sizeof(float) * RING, cudaHostAllocDefault) != cudaSuccess) { set_error("mapper_create: pinned allocation failed"); return fail(GLIC_ERR_CUDA); }
    std::memset(m->counters_host, 0, sizeof(long long) * 4 * RING);
    std::memset(m->loss_host, 0, sizeof(float) * RING);
    if ((e = set_bin_capacity(m, std::max<int64_t>((int64_t)16 * m->Pcap, (int64_t)1 << 22)))) return fail(e);
    cudaEventRecord(m->gt_free[0], m->stream); cudaEventRecord(m->gt_free[1], m->stream);
    *out = m;
    return GLIC_OK;
}

int glic_mapper_destroy(glic_mapper* m) {
    if (!m) return GLIC_OK;
    cudaDeviceSynchronize();
    for (void* p : m->opened) cudaIpcCloseMemHandle(p);
    void* bufs[] = {m->params, m->m1, m->m2, m->act_opacity, m->act_scales, m->act_rots, m->radii, m->g_conic, m->geom_ws,
                    m->xblock, m->image_ws, m->binning_ws, m->sample_ws, m->loss_scratch, m->eval_scratch, m->extend_ws, m->color, m->final_T, m->dL_dpix,
                    m->gt_dev[0], m->gt_dev[1], m->loss_dev, m->eval_out, m->vis_acc};
    for (void* p : bufs) if (p) cudaFree(p);
    for (auto& kf : m->train) if (kf.cam_dev) cudaFree(kf.cam_dev);
    for (auto& kf : m->test) if (kf.cam_dev) cudaFree(kf.cam_dev);
    if (m->counters_host) cudaFreeHost(m->counters_host);
    if (m->loss_host) cudaFreeHost(m->loss_host);
    cudaEvent_t evs[] = {m->gt_free[0], m->gt_free[1], m->gt_ready[0], m->gt_ready[1], m->ev_flags, m->ev_col, m->ev_push, m->t0, m->t1};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    for (int r = 0; r < RING; ++r) if (m->ring_ev[r]) cudaEventDestroy(m->ring_ev[r]);
    if (m->stream) cudaStreamDestroy(m->stream);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    if (m->push_stream) cudaStreamDestroy(m->push_stream);
    cudaGetLastError();
    delete m;
    return GLIC_OK;
}

int glic_mapper_initialize(glic_mapper* m, uint32_t P, const float* xyz, const float* f_dc, const float* f_rest, const float* opacity_logit,
                           const float* log_scale, const float* rot) {
    if (!m || (P && (!xyz || !f_dc || !opacity_logit || !log_scale || !rot))) { set_error("mapper_initialize: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    MAP_TRY(set_capacity(m, P));
    cudaStream_t s = m->stream;
    const size_t nfl = arena_floats(m, m->Pcap);
    GLIC_CUDA_TRY(cudaMemsetAsync(m->params, 0, nfl * 4, s));
    GLIC_CUDA_TRY(cudaMemsetAsync(m->m1, 0, nfl * 4, s));
    GLIC_CUDA_TRY(cudaMemsetAsync(m->m2, 0, nfl * 4, s));
    float* g[6];
    arena_ptrs(m, m->params, m->Pcap, g);
    const float* src[6] = {rot, xyz, log_scale, opacity_logit, f_dc, f_rest};
    const size_t kk[6] = {4, 3, 3, 1, 3, (size_t)3 * m->M};
    for (int q = 0; q < 6; ++q)
        if (P && kk[q] && src[q]) GLIC_CUDA_TRY(cudaMemcpyAsync(g[q], src[q], sizeof(float) * kk[q] * P, cudaMemcpyHostToDevice, s));
    GLIC_CUDA_TRY(cudaStreamSynchronize(s));
    m->P = P;
    m->st.num_gaussians = P;
    if ((int64_t)16 * P > m->bin_cap) MAP_TRY(set_bin_capacity(m, (int64_t)16 * P));
    return GLIC_OK;
}

int glic_mapper_add_keyframe(glic_mapper* m, const glic_keyframe* kfi, int is_train) {
    if (!m || !kfi || !kfi->image) { set_error("mapper_add_keyframe: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    Keyframe kf;
    std::memcpy(kf.R_wc, kfi->R_wc, sizeof(kf.R_wc)); std::memcpy(kf.t_wc, kfi->t_wc, sizeof(kf.t_wc));
    kf.image = kfi->image;
    camera_block(m->cfg.width, m->cfg.height, m->cfg.fx, m->cfg.fy, m->cfg.cx, m->cfg.cy, kf.R_wc, kf.t_wc, kf.cam);
    MAP_TRY(dev_alloc((void**)&kf.cam_dev, 36 * sizeof(float), false));
    GLIC_CUDA_TRY(cudaMemcpy(kf.cam_dev, kf.cam.view, 36 * sizeof(float), cudaMemcpyHostToDevice));   // view | proj | campos are contiguous
    kf.view.viewmatrix = kf.cam_dev; kf.view.projmatrix = kf.cam_dev + 16; kf.view.campos = kf.cam_dev + 32;
    kf.view.tan_fovx = kf.cam.tanfovx; kf.view.tan_fovy = kf.cam.tanfovy;
    kf.view.limx_neg = kf.cam.lim[0]; kf.view.limx_pos = kf.cam.lim[1]; kf.view.limy_neg = kf.cam.lim[2]; kf.view.limy_pos = kf.cam.lim[3];
    kf.view.width = m->cfg.width; kf.view.height = m->cfg.height;
    (is_train ? m->train : m->test).push_back(kf);
    return GLIC_OK;
}

int glic_mapper_extend(glic_mapper* m, int n, const float* points, const float* colors, const float* depth_rsp) {
    if (!m || n < 0 || (n && (!points || !colors || !depth_rsp))) { set_error("mapper_extend: bad arguments"); return GLIC_ERR_INVALID_ARGUMENT; }
    m->st.last_inserted = 0;
    if (n == 0) return GLIC_OK;
    if (m->train.empty()) { set_error("mapper_extend: no training keyframe yet"); return GLIC_ERR_INVALID_ARGUMENT; }
    MAP_TRY(poll_ring(m, true));
    MAP_TRY(set_capacity(m, std::max<uint32_t>(m->P + (uint32_t)n, m->Pcap < m->P + (uint32_t)n ? 2 * m->Pcap : 0)));
    cudaStream_t s = m->stream;
    const int W = m->cfg.width, H = m->cfg.height;
    const Keyframe& kf = m->train.back();
    // staging: points | colors | depth | outputs (keep_idx, xyz, f_dc, log_scale, rot, opacity) | glic_extend's own workspace
    const size_t need = glic_extend_bytes(n, W, H) + (size_t)n * (3 + 3 + 1 + 1 + 3 + 3 + 3 + 4 + 1) * 4 + 4096;
    if (need > m->extend_bytes) {
        GLIC_CUDA_TRY(cudaStreamSynchronize(s));
        if (m->extend_ws) cudaFree(m->extend_ws);
        MAP_TRY(dev_alloc(&m->extend_ws, need, false));
        m->extend_bytes = need;
    }
    Carver c(m->extend_ws);
    float* d_pts = c.take<float>((size_t)3 * n); float* d_col = c.take<float>((size_t)3 * n); float* d_dep = c.take<float>(n);
    int* d_keep = c.take<int>(n); float* d_xyz = c.take<float>((size_t)3 * n); float* d_dc = c.take<float>((size_t)3 * n);
    float* d_ls = c.take<float>((size_t)3 * n); float* d_rot = c.take<float>((size_t)4 * n); float* d_op = c.take<float>(n);
    void* d_ws = c.take<char>(glic_extend_bytes(n, W, H));
    GLIC_CUDA_TRY(cudaEventRecord(m->t0, s));
    GLIC_CUDA_TRY(cudaMemcpyAsync(d_pts, points, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, s));
    GLIC_CUDA_TRY(cudaMemcpyAsync(d_col, colors, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, s));
    GLIC_CUDA_TRY(cudaMemcpyAsync(d_dep, depth_rsp, sizeof(float) * n, cudaMemcpyHostToDevice, s));
    // alpha-only render of the newest keyframe (gaussian.cpp:506-507); the binning capacity is settled synchronously here
    if (m->P) {
        MAP_TRY(activate(m));
        for (int attempt = 0; attempt < 4; ++attempt) {
            long long* cnt = m->counters_host;              // ring slot 0 is free: poll_ring(wait) drained the ring
            MAP_TRY(forward_view(m, kf, 1, cnt));
            GLIC_CUDA_TRY(cudaStreamSynchronize(s));
            if (!cnt[2]) break;
            ++m->st.overflow_regrows;
            MAP_TRY(set_bin_capacity(m, cnt[0] + cnt[0] / 4));
        }
    } else {
        std::vector<float> ones((size_t)W * H, 1.0f);                                       // empty model: alpha = 1 - T = 0
        GLIC_CUDA_TRY(cudaMemcpyAsync(m->final_T, ones.data(), sizeof(float) * ones.size(), cudaMemcpyHostToDevice, s));
        GLIC_CUDA_TRY(cudaStreamSynchronize(s));
    }
    float Rcw[9], tcw[3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rcw[3 * i + j] = kf.cam.view[4 * j + i];   // view is column-major Rt
    for (int i = 0; i < 3; ++i) tcw[i] = kf.cam.view[12 + i];
    int count = 0;
    MAP_TRY(glic_extend(n, d_pts, d_col, d_dep, Rcw, tcw, m->cfg.fx, m->cfg.fy, m->cfg.cx, m->cfg.cy, W, H, m->final_T, m->cfg.scaling_scale,
                        d_ws, glic_extend_bytes(n, W, H), d_keep, d_xyz, d_dc, d_ls, d_rot, d_op, &count, s));
    MAP_TRY(glic_arena_append(m->params, m->m1, m->m2, m->P, m->Pcap, (uint32_t)m->M, (uint32_t)count, d_xyz, d_dc, d_ls, d_rot, d_op, s));
    GLIC_CUDA_TRY(cudaEventRecord(m->t1, s));
    GLIC_CUDA_TRY(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&m->st.ms_extend, m->t0, m->t1);
    m->P += (uint32_t)count;
    m->st.num_gaussians = m->P;
    m->st.last_inserted = (uint32_t)count;
    return GLIC_OK;
}

int glic_mapper_sample_views(glic_mapper* m, int* views, int* count) {
    if (!m || !views || !count) { set_error("mapper_sample_views: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    // gaussian.cpp:645-662: every training camera when there are <= max_iters of them, otherwise a uniform subset of
    // max_iters without replacement; then a shuffle.  (Fisher-Yates on an explicit seed: identical on every rank.)
    const int n = (int)m->train.size();
    std::vector<int> all(n);
    for (int i = 0; i < n; ++i) all[i] = i;
    const int take = std::min(n, m->cfg.max_iters);
    for (int i = 0; i < take; ++i) {
        const int j = i + (int)(m->rng() % (uint64_t)(n - i));
        std::swap(all[i], all[j]);
    }
    for (int i = 0; i < take; ++i) views[i] = all[i];
    *count = take;
    return GLIC_OK;
}

int glic_mapper_optimize(glic_mapper* m, const int* views_host, int n_views) {
    if (!m) { set_error("mapper_optimize: null mapper"); return GLIC_ERR_INVALID_ARGUMENT; }
    std::vector<int> list;
    if (views_host && n_views > 0) list.assign(views_host, views_host + n_views);
    else {
        list.resize((size_t)std::max(1, m->cfg.max_iters));
        int cnt = 0;
        MAP_TRY(glic_mapper_sample_views(m, list.data(), &cnt));
        list.resize((size_t)cnt);
    }
    if (list.empty() || m->P == 0) return GLIC_OK;
    GLIC_CUDA_TRY(cudaMemsetAsync(m->vis_acc, 0, sizeof(unsigned int), m->stream));
    GLIC_CUDA_TRY(cudaEventRecord(m->t0, m->stream));
    const size_t iters = (list.size() + m->S - 1) / m->S;
    std::vector<int> batch((size_t)m->S);
    for (size_t it = 0; it < iters; ++it) {
        for (int sl = 0; sl < m->S; ++sl) batch[sl] = list[(it * m->S + sl) % list.size()];     // the last batch wraps around
        MAP_TRY(iteration(m, batch.data()));
    }
    GLIC_CUDA_TRY(cudaEventRecord(m->t1, m->stream));
    unsigned int vis = 0;
    GLIC_CUDA_TRY(cudaMemcpyAsync(&vis, m->vis_acc, sizeof(vis), cudaMemcpyDeviceToHost, m->stream));
    GLIC_CUDA_TRY(cudaStreamSynchronize(m->stream));
    MAP_TRY(poll_ring(m, true));
    cudaEventElapsedTime(&m->st.ms_optimize, m->t0, m->t1);
    m->st.mean_visible = (double)vis / (double)iters;
    if (m->connected) MAP_TRY(glic_p2p_check(m->xblock, m->n_geo, 0, m->stream));
    return GLIC_OK;
}

int glic_mapper_evaluate(glic_mapper* m, int is_train, int index, float* psnr, float* ssim) {
    if (!m || !psnr || !ssim) { set_error("mapper_evaluate: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    auto& list = is_train ? m->train : m->test;
    if (index < 0 || index >= (int)list.size()) { set_error("mapper_evaluate: keyframe index out of range"); return GLIC_ERR_INVALID_ARGUMENT; }
    MAP_TRY(poll_ring(m, true));
    const Keyframe& kf = list[index];
    cudaStream_t s = m->stream;
    const int W = m->cfg.width, H = m->cfg.height;
    MAP_TRY(activate(m));
    for (int attempt = 0; attempt < 4; ++attempt) {
        long long* cnt = m->counters_host;
        MAP_TRY(forward_view(m, kf, 0, cnt));
        GLIC_CUDA_TRY(cudaStreamSynchronize(s));
        if (!cnt[2]) break;
        ++m->st.overflow_regrows;
        MAP_TRY(set_bin_capacity(m, cnt[0] + cnt[0] / 4));
    }
    GLIC_CUDA_TRY(cudaMemcpyAsync(m->gt_dev[0], kf.image, sizeof(float) * 3 * (size_t)W * H, cudaMemcpyHostToDevice, s));
    MAP_TRY(glic_eval_psnr_ssim(3, H, W, m->color, m->gt_dev[0], m->eval_out, m->eval_scratch, m->eval_bytes, s));
    float out2[2] = {0.f, 0.f};
    GLIC_CUDA_TRY(cudaMemcpyAsync(out2, m->eval_out, sizeof(out2), cudaMemcpyDeviceToHost, s));
    GLIC_CUDA_TRY(cudaStreamSynchronize(s));
    *psnr = out2[0]; *ssim = out2[1];
    return GLIC_OK;
}

int glic_mapper_download(glic_mapper* m, float* xyz, float* f_dc, float* f_rest, float* opacity_logit, float* log_scale, float* rot,
                         float* exp_avg_packed, float* exp_avg_sq_packed) {
    if (!m) { set_error("mapper_download: null mapper"); return GLIC_ERR_INVALID_ARGUMENT; }
    GLIC_CUDA_TRY(cudaStreamSynchronize(m->stream));
    float* g[6];
    arena_ptrs(m, m->params, m->Pcap, g);
    float* dst[6] = {rot, xyz, log_scale, opacity_logit, f_dc, f_rest};
    const size_t kk[6] = {4, 3, 3, 1, 3, (size_t)3 * m->M};
    for (int q = 0; q < 6; ++q)
        if (dst[q] && kk[q] && m->P) GLIC_CUDA_TRY(cudaMemcpy(dst[q], g[q], sizeof(float) * kk[q] * m->P, cudaMemcpyDeviceToHost));
    float* mom[2] = {exp_avg_packed, exp_avg_sq_packed};
    float* srcm[2] = {m->m1, m->m2};
    for (int a = 0; a < 2; ++a) {
        if (!mom[a] || !m->P) continue;
        float* gm[6];
        arena_ptrs(m, srcm[a], m->Pcap, gm);
        size_t off = 0;                                     // packed (capacity = P) layout: glic_packed_offsets(P, M)
        for (int q = 0; q < 6; ++q) {
            if (kk[q]) GLIC_CUDA_TRY(cudaMemcpy(mom[a] + off, gm[q], sizeof(float) * kk[q] * m->P, cudaMemcpyDeviceToHost));
            off += kk[q] * m->P;
        }
    }
    return GLIC_OK;
}

int glic_mapper_save_map(glic_mapper* m, const char* path) {
    if (!m || !path) { set_error("mapper_save_map: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    const size_t P = m->P, M = (size_t)m->M;
    std::vector<float> xyz(3 * P), dc(3 * P), rest(3 * M * P), op(P), ls(3 * P), rot(4 * P);
    MAP_TRY(glic_mapper_download(m, xyz.data(), dc.data(), M ? rest.data() : nullptr, op.data(), ls.data(), rot.data(), nullptr, nullptr));
    return glic_ply_write(path, (uint32_t)P, (uint32_t)M, xyz.data(), dc.data(), rest.data(), op.data(), ls.data(), rot.data());
}

int glic_mapper_stats_get(glic_mapper* m, glic_mapper_stats* out) {
    if (!m || !out) { set_error("mapper_stats: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    *out = m->st;
    return GLIC_OK;
}

int glic_mapper_set_option(glic_mapper* m, int option, int value) {
    if (!m) { set_error("mapper_set_option: null mapper"); return GLIC_ERR_INVALID_ARGUMENT; }
    switch (option) {
        case GLIC_MAPPER_OPT_OPTIMIZER: m->run_optimizer = value != 0; return GLIC_OK;
        case GLIC_MAPPER_OPT_BINNING_PAIRS:            // start from a chosen (possibly too small) binning capacity: exercises the overflow path
            if (value < 1) { set_error("mapper_set_option: binning capacity must be positive"); return GLIC_ERR_INVALID_ARGUMENT; }
            MAP_TRY(poll_ring(m, true));
            return set_bin_capacity(m, value, /*exact=*/true);
        default: set_error("mapper_set_option: unknown option"); return GLIC_ERR_INVALID_ARGUMENT;
    }
}

int glic_mapper_synchronize(glic_mapper* m) {
    if (!m) return GLIC_OK;
    GLIC_CUDA_TRY(cudaStreamSynchronize(m->stream));
    return poll_ring(m, true);
}

int glic_mapper_export(glic_mapper* m, unsigned char* handle64) {
    if (!m || !handle64) { set_error("mapper_export: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    cudaIpcMemHandle_t h;
    GLIC_CUDA_TRY(cudaIpcGetMemHandle(&h, m->xblock));
    std::memcpy(handle64, &h, 64);
    return GLIC_OK;
}

int glic_mapper_connect(glic_mapper* m, const unsigned char* handles) {
    if (!m || !handles) { set_error("mapper_connect: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    for (int q = 0; q < m->cfg.world; ++q) {
        if (q == m->cfg.rank) { m->peers[q] = m->xblock; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + 64 * (size_t)q, 64);
        void* p = nullptr;
        GLIC_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        m->peers[q] = p;
        m->opened.push_back(p);
    }
    m->connected = true;
    return GLIC_OK;
}

}  // extern "C"
