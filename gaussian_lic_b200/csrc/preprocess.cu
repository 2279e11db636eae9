// preprocess.cu -- forward per-Gaussian stage, fused with the tile-count prefix sum, and the
// (tile|depth) key emission.
//
// Replaces, behind the C ABI: FORWARD::preprocess (reference forward.cu:232-319, 518-580),
// cub::DeviceScan::InclusiveSum (rasterizer_impl.cu:395) and duplicateWithKeys
// (rasterizer_impl.cu:59-193).  Design (not a translation):
//   * one kernel does cull + projection + EWA + SH colour + exact tile counting AND the global
//     inclusive scan of tiles_touched (single-pass decoupled look-back over dynamically ordered
//     CTAs), so tiles_touched never round-trips through HBM and no scan kernel / temp exists;
//   * one 48-byte AoS splat record per Gaussian feeds emit, render-forward and render-backward;
//   * every key-defining operation uses the fixed-order intrinsics of geom_math.cuh.
#include "geom_math.cuh"

namespace glic {

// ---- SH basis (auxiliary.h:22-39, forward.cu:29-77) ---------------------------------------
__device__ __constant__ float kSH_C1 = 0.4886025119029199f;
__device__ __constant__ float kSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                           -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float kSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                           0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                           -0.5900435899266435f};
constexpr float kSH_C0 = 0.28209479177387814f;

// ---- decoupled look-back status words ------------------------------------------------------
constexpr unsigned long long ST_AGG = 1ull << 32;
constexpr unsigned long long ST_PREFIX = 2ull << 32;

GLIC_DI void st_release(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
GLIC_DI unsigned long long ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Exclusive prefix of `aggregate` over all CTAs with a smaller ticket.  Called by warp 0 only.
GLIC_DI uint32_t lookback_exclusive(unsigned long long* status, int block, uint32_t aggregate, int lane) {
    if (block == 0) {
        if (lane == 0) st_release(&status[0], ST_PREFIX | aggregate);
        return 0;
    }
    if (lane == 0) st_release(&status[block], ST_AGG | aggregate);
    uint32_t exclusive = 0;
    int look = block - 1;
    unsigned spins = 0;
    while (true) {
        const int j = look - lane;
        unsigned long long st = ST_PREFIX;      // virtual predecessors before CTA 0 contribute 0
        if (j >= 0) st = ld_acquire(&status[j]);
        while (__any_sync(0xffffffffu, (st >> 32) == 0)) {
            if (j >= 0 && (st >> 32) == 0) st = ld_acquire(&status[j]);
            if (++spins > (1u << 24)) __trap();  // never hang the GPU: fail loudly instead
        }
        const unsigned pref = __ballot_sync(0xffffffffu, (st >> 32) == 2);
        const int first = pref ? (__ffs(pref) - 1) : 31;
        uint32_t v = (lane <= first) ? (uint32_t)st : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        exclusive += v;
        if (pref) break;
        look -= 32;
    }
    if (lane == 0) st_release(&status[block], ST_PREFIX | (unsigned long long)(exclusive + aggregate));
    return exclusive;
}


// ---- warp-balanced walk over every (Gaussian, candidate tile) pair of a warp ------------------
// The 32 Gaussians of a warp own rects of wildly different sizes (1 .. thousands of tiles).  Instead of
// one thread looping over its own rect, the warp flattens all its pairs into one list and tests 32 of
// them per iteration: iterations = ceil(sum n / 32) instead of max n.  Lane l of iteration `base` takes
// flattened item g = base + l, finds its owner by a 5-step binary search over the inclusive prefix held
// in the lanes (shuffles), fetches the owner's parameters with dynamic-source shuffles and runs the
// exact tile test.  Owners count their accepted tiles from the ballot restricted to their lane segment.
// EMIT: the item additionally writes key/value at offset(owner) + #accepted-before, i.e. row-major order
// inside each Gaussian's slot range, exactly like a sequential walk.
GLIC_DI uint32_t seg_mask(int s0, int s1) {   // bits [s0, s1), 0 <= s0 <= s1 <= 32
    const uint32_t hi = s1 >= 32 ? 0xffffffffu : ((1u << s1) - 1u);
    const uint32_t lo = s0 >= 32 ? 0xffffffffu : ((1u << s0) - 1u);
    return hi & ~lo;
}

template <bool EMIT>
GLIC_DI uint32_t warp_tile_walk(int n, float mx, float my, float cox, float coy, float coz, float thr, int x0, int y0, int rw,
                                int grid_x, uint32_t dbits, uint32_t idx, uint32_t off, uint64_t* __restrict__ keys,
                                uint32_t* __restrict__ vals) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    uint32_t incl = (uint32_t)n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    const uint32_t excl = incl - (uint32_t)n;
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t my_count = 0;
    for (uint32_t base = 0; base < total; base += 32) {
        const uint32_t g = base + lane;
        const bool valid = g < total;
        int lo = 0, hi = 31;                      // first lane j with incl_j > g
#pragma unroll
        for (int it = 0; it < 5; ++it) {
            const int mid = (lo + hi) >> 1;
            const uint32_t v = __shfl_sync(FULL, incl, mid);
            if (v > g) hi = mid; else lo = mid + 1;
        }
        const int owner = lo & 31;
        const uint32_t o_excl = __shfl_sync(FULL, excl, owner);
        const float o_mx = __shfl_sync(FULL, mx, owner), o_my = __shfl_sync(FULL, my, owner);
        const float o_cx = __shfl_sync(FULL, cox, owner), o_cy = __shfl_sync(FULL, coy, owner), o_cz = __shfl_sync(FULL, coz, owner);
        const float o_thr = __shfl_sync(FULL, thr, owner);
        const int o_x0 = __shfl_sync(FULL, x0, owner), o_y0 = __shfl_sync(FULL, y0, owner);
        const int o_rw = max(__shfl_sync(FULL, rw, owner), 1);
        const int t = valid ? (int)(g - o_excl) : 0;
        const int ty = t / o_rw + o_y0, tx = t % o_rw + o_x0;
        const bool ok = valid && tile_max_power(o_cx, o_cy, o_cz, o_mx, o_my, tx, ty) <= o_thr;
        const uint32_t acc = __ballot_sync(FULL, ok);
        if (EMIT) {
            const uint32_t o_n = __shfl_sync(FULL, (uint32_t)n, owner);
            const uint32_t o_cnt = __shfl_sync(FULL, my_count, owner);
            const uint32_t o_off = __shfl_sync(FULL, off, owner);
            const uint32_t o_idx = __shfl_sync(FULL, idx, owner);
            const uint32_t o_db = __shfl_sync(FULL, dbits, owner);
            if (ok) {
                const int s0 = (int)(max(o_excl, base) - base), s1 = (int)(min(o_excl + o_n, base + 32u) - base);
                const uint32_t pos = o_off + o_cnt + __popc(acc & seg_mask(s0, s1) & lt);
                keys[pos] = ((uint64_t)(uint32_t)(ty * grid_x + tx) << 32) | (uint64_t)o_db;
                vals[pos] = o_idx;
            }
        }
        if (n > 0 && incl > base && excl < base + 32u) {
            const int s0 = (int)(max(excl, base) - base), s1 = (int)(min(incl, base + 32u) - base);
            my_count += __popc(acc & seg_mask(s0, s1));
        }
    }
    return my_count;
}

__global__ void __launch_bounds__(PRE_THREADS)
preprocess_forward_kernel(int P, int D, int M, const float* __restrict__ means, const float* __restrict__ scales,
                          float mod, const float4* __restrict__ rots, const float* __restrict__ opac,
                          const float* __restrict__ dc, const float* __restrict__ sh, ViewParams vp, bool no_color,
                          int* __restrict__ radii, GeomState g) {
    __shared__ float s_view[16], s_proj[16], s_cam[3];
    __shared__ unsigned s_block;
    __shared__ uint32_t s_warp_sum[PRE_THREADS / 32];
    __shared__ uint32_t s_block_excl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_block = atomicAdd(&g.hdr->ticket, 1u);
    if (tid < 16) s_view[tid] = vp.view[tid];
    else if (tid < 32) s_proj[tid - 16] = vp.proj[tid - 16];
    else if (tid < 35) s_cam[tid - 32] = vp.campos[tid - 32];
    __syncthreads();
    const int block = (int)s_block;
    const int idx = block * PRE_THREADS + tid;

    uint32_t tiles = 0;
    int radius = 0;
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
    float blue = 0.f, depth = 0.f;
    unsigned clampbits = 0;
    // ---- phase 1: geometry of my Gaussian (fixed-order arithmetic) ----------------------------------
    float px = 0.f, py = 0.f, pz = 0.f, mx = 0.f, my = 0.f, cox = 0.f, coy = 0.f, coz = 0.f, o = 0.f, thr = 0.f, tz = 0.f;
    int irad = 0, n = 0, rx0 = 0, ry0 = 0, rw = 1;
    if (idx < P) {
        px = means[3 * idx]; py = means[3 * idx + 1]; pz = means[3 * idx + 2];
        Cov3 c3;
        cov3d_from_scale_rot(scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2], mod, rots[idx], c3);
        Cov2 c2;
        cov2d_project(px, py, pz, s_view, vp.focal_x, vp.focal_y, vp.limx_neg, vp.limx_pos, vp.limy_neg, vp.limy_pos,
                      c3.c, c2);
        tz = c2.tz;
        bool active = !(c2.tz <= 0.2f);                              // near cull, auxiliary.h:160
        const float hx = xform_row(s_proj, 0, px, py, pz);
        const float hy = xform_row(s_proj, 1, px, py, pz);
        const float hw = xform_row(s_proj, 3, px, py, pz);
        const float pw = __frcp_rn(fadd(hw, 0.0000001f));
        const float ndcx = fmul(hx, pw), ndcy = fmul(hy, pw);
        const float det = ffma(c2.a, c2.c, -fmul(c2.b, c2.b));        // a*c - b*b as fma(a, c, -(b*b))
        if (det == 0.0f) active = false;
        const float det_inv = __frcp_rn(det);
        cox = fmul(c2.c, det_inv); coy = fmul(det_inv, -c2.b); coz = fmul(c2.a, det_inv);
        o = opac[idx];
        if (o < (1.0f / 255.0f)) active = false;                     // forward.h:30, forward.cu:293
        if (active) {
            const float mid = fmul(fadd(c2.a, c2.c), 0.5f);
            const float lambda1 = fadd(mid, __fsqrt_rn(fmaxf(ffma(mid, mid, -det), 0.1f)));
            const float frad = ceilf(fmul(__fsqrt_rn(lambda1), 3.0f));
            mx = ndc_to_pix(ndcx, vp.W); my = ndc_to_pix(ndcy, vp.H);
            irad = (int)frad;
            const TileRect rc = tile_rect(mx, my, irad, vp.grid_x, vp.grid_y);
            thr = logf(__fdiv_rn(o, 1.0f / 255.0f));                 // forward.cu:302
            rw = rc.x1 - rc.x0; rx0 = rc.x0; ry0 = rc.y0;
            n = (rc.y1 - rc.y0) * rw;
        }
    }
    // ---- phase 2: exact tile counting, balanced across the warp -------------------------------------------
    const uint32_t cnt = warp_tile_walk<false>(n, mx, my, cox, coy, coz, thr, rx0, ry0, rw, vp.grid_x, 0u, 0u, 0u, nullptr, nullptr);
    // ---- phase 3: colour of the survivors -----------------------------------------------------------------
    if (cnt > 0) {
        {
            {
                tiles = cnt;
                radius = irad;
                depth = tz;
                float red = 0.f, green = 0.f;
                if (!no_color) {
                    // SH -> RGB (forward.cu:29-77); tolerance-pinned, natural arithmetic.
                    float dx = px - s_cam[0], dy = py - s_cam[1], dz = pz - s_cam[2];
                    const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
                    dx *= inv; dy *= inv; dz *= inv;
                    float res[3];
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) res[ch] = kSH_C0 * dc[3 * idx + ch];
                    if (D > 0) {
                        const float* s = sh + (size_t)idx * M * 3;
                        const float x = dx, y = dy, z = dz;
                        float b[15];
                        b[0] = -kSH_C1 * y; b[1] = kSH_C1 * z; b[2] = -kSH_C1 * x;
                        int nb = 3;
                        if (D > 1) {
                            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                            b[3] = kSH_C2[0] * xy; b[4] = kSH_C2[1] * yz; b[5] = kSH_C2[2] * (2.0f * zz - xx - yy);
                            b[6] = kSH_C2[3] * xz; b[7] = kSH_C2[4] * (xx - yy);
                            nb = 8;
                            if (D > 2) {
                                b[8] = kSH_C3[0] * y * (3.0f * xx - yy);
                                b[9] = kSH_C3[1] * xy * z;
                                b[10] = kSH_C3[2] * y * (4.0f * zz - xx - yy);
                                b[11] = kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                                b[12] = kSH_C3[4] * x * (4.0f * zz - xx - yy);
                                b[13] = kSH_C3[5] * z * (xx - yy);
                                b[14] = kSH_C3[6] * x * (xx - 3.0f * yy);
                                nb = 15;
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 15; ++k) {
                            if (k < nb) {
                                res[0] += b[k] * s[3 * k + 0];
                                res[1] += b[k] * s[3 * k + 1];
                                res[2] += b[k] * s[3 * k + 2];
                            }
                        }
                    }
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        res[ch] += 0.5f;
                        if (res[ch] < 0.0f) clampbits |= 1u << ch;
                        res[ch] = fmaxf(res[ch], 0.0f);
                    }
                    red = res[0]; green = res[1]; blue = res[2];
                }
                r0 = make_float4(mx, my, cox, coy);
                r1 = make_float4(coz, o, red, green);
            }
        }
    }

    // ---- CTA-wide inclusive scan of `tiles`, then the cross-CTA look-back -----------------
    uint32_t incl = tiles;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t ws = lane < PRE_THREADS / 32 ? s_warp_sum[lane] : 0u;
        uint32_t wi = ws;
#pragma unroll
        for (int o = 1; o < PRE_THREADS / 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += n;
        }
        const uint32_t aggregate = __shfl_sync(0xffffffffu, wi, PRE_THREADS / 32 - 1);
        if (lane < PRE_THREADS / 32) s_warp_sum[lane] = wi - ws;      // exclusive per-warp base
        const uint32_t excl = lookback_exclusive(g.scan_status, block, aggregate, lane);
        if (lane == 0) {
            s_block_excl = excl;
            if (block == (int)gridDim.x - 1) g.hdr->total = excl + aggregate;
        }
    }
    __syncthreads();
    if (idx < P) {
        g.offsets[idx] = s_block_excl + s_warp_sum[warp] + incl;      // inclusive, like the reference
        radii[idx] = radius;
        g.rec[3 * idx + 0] = r0;
        g.rec[3 * idx + 1] = r1;
        g.rec[3 * idx + 2] = make_float4(blue, depth, __int_as_float(radius), __uint_as_float(tiles));
        g.clamped[idx] = (uint8_t)clampbits;
    }
}

// ---- key emission ----------------------------------------------------------------------------
// One thread per Gaussian re-walks its rect with the same exact test and writes
// key = (tile << 32) | bits(depth), value = Gaussian index into its [offsets[i-1], offsets[i]) slots.
__global__ void __launch_bounds__(256)
emit_keys_kernel(int P, ViewParams vp, GeomState g, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int n = 0, x0 = 0, y0 = 0, rw = 1;
    float mx = 0.f, my = 0.f, cox = 0.f, coy = 0.f, coz = 0.f, thr = 0.f;
    uint32_t dbits = 0, off = 0;
    if (idx < P) {
        const float4 r2 = g.rec[3 * idx + 2];
        const uint32_t tiles = __float_as_uint(r2.w);
        if (tiles != 0) {
            const float4 r0 = g.rec[3 * idx + 0];
            const float4 r1 = g.rec[3 * idx + 1];
            off = g.offsets[idx] - tiles;
            mx = r0.x; my = r0.y; cox = r0.z; coy = r0.w; coz = r1.x;
            const TileRect rc = tile_rect(mx, my, __float_as_int(r2.z), vp.grid_x, vp.grid_y);
            thr = logf(__fdiv_rn(r1.y, 1.0f / 255.0f));
            rw = rc.x1 - rc.x0; x0 = rc.x0; y0 = rc.y0;
            n = (rc.y1 - rc.y0) * rw;
            dbits = __float_as_uint(r2.y);
        }
    }
    warp_tile_walk<true>(n, mx, my, cox, coy, coz, thr, x0, y0, rw, vp.grid_x, dbits, (uint32_t)idx, off, keys, vals);
}

int launch_preprocess_forward(int P, int D, int M, const float* means, const float* scales, float mod,
                              const float* rots, const float* opac, const float* dc, const float* sh,
                              const ViewParams& vp, bool no_color, int* radii, GeomState g, cudaStream_t s) {
    const int blocks = (P + PRE_THREADS - 1) / PRE_THREADS;
    // header (ticket/total) and the look-back status words start at zero
    GLIC_CUDA_TRY(cudaMemsetAsync(g.hdr, 0, sizeof(GeomHeader), s));
    GLIC_CUDA_TRY(cudaMemsetAsync(g.scan_status, 0, sizeof(unsigned long long) * (blocks + 1), s));
    preprocess_forward_kernel<<<blocks, PRE_THREADS, 0, s>>>(P, D, M, means, scales, mod,
                                                             reinterpret_cast<const float4*>(rots), opac, dc, sh, vp,
                                                             no_color, radii, g);
    GLIC_LAUNCH_CHECK();
    return GLIC_OK;
}

int launch_emit_keys(int P, const ViewParams& vp, GeomState g, uint64_t* keys, uint32_t* vals, cudaStream_t s) {
    emit_keys_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, vp, g, keys, vals);
    GLIC_LAUNCH_CHECK();
    return GLIC_OK;
}

}  // namespace glic
