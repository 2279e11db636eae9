// render.cu -- tile ranges, bucket offsets, per-tile front-to-back blend (forward) and the
// per-splat ("bucketed") backward.
//
// Replaces identifyTileRanges + perTileBucketCount + InclusiveSum (reference
// rasterizer_impl.cu:195-231,426-442), renderCUDA (forward.cu:321-481) incl. the D2D colour
// copy (rasterizer_impl.cu:471), and PerGaussianRenderCUDA (backward.cu:379-597).
//
// Design notes (ours):
//   * splat data of a batch is gathered ONCE per tile from the 48-byte records into shared
//     memory (xy, conic, opacity AND rgb -- the reference re-gathers rgb from global memory for
//     every contributing pixel-splat pair);
//   * each staged splat carries an 8-bit "row-pair" mask: warp w (rows 2w,2w+1 of the tile)
//     skips, warp-uniformly, splats whose alpha>=1/255 ellipse cannot reach its rows.  The mask
//     is conservative, so results are bit-identical to evaluating every pair;
//   * checkpoints are one float4 (T, C.r, C.g, C.b) per pixel per 32-splat bucket -> 16-byte
//     coalesced stores / loads;
//   * the final colour is written to the caller's image and to the saved-state copy by the same
//     kernel (no separate D2D copy);
//   * backward runs exactly one warp per non-empty bucket, 8 buckets per CTA, and feeds the
//     shuffle pipeline from a warp-private shared-memory slab filled with coalesced loads.
// The per-pixel arithmetic (power, alpha, T, C) keeps the reference's operation order through the
// fixed-order intrinsics of geom_math.cuh, so colours/transmittance agree bit-for-bit with the
// reference build on the same sorted list.
#include "geom_math.cuh"
#include "async_copy.cuh"
#include <algorithm>

namespace glic {

// ---- tile ranges -------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_ranges_kernel(int64_t R, const unsigned int* __restrict__ r_dev, const uint32_t* __restrict__ keys, uint32_t T,
                   uint2* __restrict__ ranges) {
    if (r_dev) R = min(R, (int64_t)*r_dev);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t cur = keys[i];
    if (cur >= T) return;
    if (i == 0) ranges[cur].x = 0;
    else {
        const uint32_t prev = keys[i - 1];
        if (cur != prev) {
            if (prev < T) ranges[prev].y = (uint32_t)i;
            ranges[cur].x = (uint32_t)i;
        }
    }
    if (i == R - 1) ranges[cur].y = (uint32_t)R;
}

// bucket_offsets = inclusive scan of ceil(n_t / 32); single CTA (T is a few thousand).
__global__ void __launch_bounds__(1024)
bucket_scan_kernel(int T, const uint2* __restrict__ ranges, uint32_t* __restrict__ bucket_offsets, ImageHeader* hdr,
                   long long R, const GeomHeader* __restrict__ ghdr, int buckets) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < T; base += 1024) {
        const int t = base + tid;
        uint32_t v = 0;
        if (t < T && buckets) { const uint2 r = ranges[t]; v = (r.y - r.x + BUCKET - 1) / BUCKET; }
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_tot[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += n;
            }
            warp_tot[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t out = carry + warp_tot[warp] + incl;
        if (t < T) bucket_offsets[t] = out;
        __syncthreads();
        if (tid == 1023) carry = out;
        __syncthreads();
    }
    if (tid == 0) {
        long long true_R = R, ovf = 0;
        if (ghdr) { R = ghdr->r_eff; true_R = ghdr->total; ovf = ghdr->overflow; }
        hdr->num_buckets = carry; hdr->num_rendered = R;
        hdr->counters[0] = true_R; hdr->counters[1] = carry; hdr->counters[2] = ovf; hdr->counters[3] = 0;
    }
}

// ---- forward blend -------------------------------------------------------------------------
// Splat batches are staged by the TMA engine: every thread issues one 48-byte cp.async.bulk (its splat's
// record, gathered by sorted index) into a double-buffered shared-memory tile buffer; an mbarrier per
// buffer counts the bytes.  Batch i+1 is in flight while batch i is blended, no register staging.
struct __align__(16) SplatRec { float4 a, b, c; };     // (x, y, conic.x, conic.y) (conic.z, opacity, r, g) (b, depth, radius, hy)

__global__ void __launch_bounds__(TILE_PIX)
render_forward_kernel(ViewParams vp, bool no_color, const uint2* __restrict__ ranges,
                      const uint32_t* __restrict__ point_list, const float4* __restrict__ rec,
                      const uint32_t* __restrict__ bucket_offsets, uint32_t* __restrict__ bucket_to_tile,
                      float4* __restrict__ ckpt, uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ max_contrib,
                      float* __restrict__ pixel_colors, float* __restrict__ out_color, float* __restrict__ out_T) {
    __shared__ SplatRec s_rec[2][TILE_PIX];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_red[TILE_PIX / 32];

    const int tid = threadIdx.y * TILE + threadIdx.x;
    const int warp = tid >> 5;
    const uint32_t tile = blockIdx.y * vp.grid_x + blockIdx.x;
    const int pix_min_x = blockIdx.x * TILE, pix_min_y = blockIdx.y * TILE;
    const int px = pix_min_x + threadIdx.x, py = pix_min_y + threadIdx.y;
    const bool inside = px < vp.W && py < vp.H;
    const float pfx = (float)px, pfy = (float)py;
    const size_t pid = (size_t)py * vp.W + px;
    const uint2 range = ranges[tile];
    const int n_splats = (int)(range.y - range.x);
    const int rounds = (n_splats + TILE_PIX - 1) / TILE_PIX;

    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
    uint32_t bbm = 0;
    if (!no_color) {
        bbm = tile == 0 ? 0u : bucket_offsets[tile - 1];
        const int nb = (n_splats + BUCKET - 1) / BUCKET;
        for (int b = tid; b < nb; b += TILE_PIX) bucket_to_tile[bbm + b] = tile;
    }
    __syncthreads();
    // issue the gather of batch `r` into buffer r & 1 (one 48-byte bulk copy per thread)
    auto prefetch = [&](int r) {
        const int cnt = min(TILE_PIX, n_splats - r * TILE_PIX);
        if (tid == 0) mbar_arrive_expect_tx(&s_bar[r & 1], (uint32_t)(cnt * sizeof(SplatRec)));
        if (tid < cnt) {
            const uint32_t id = point_list[range.x + r * TILE_PIX + tid];
            bulk_copy_g2s(&s_rec[r & 1][tid], rec + 3 * (size_t)id, (uint32_t)sizeof(SplatRec), &s_bar[r & 1]);
        }
    };
    if (rounds > 0) prefetch(0);

    // this warp covers tile rows 2w, 2w+1: splats whose alpha >= 1/255 ellipse cannot reach them are skipped
    // warp-uniformly (conservative half-extent `hy` precomputed per splat => bit-identical results)
    const float row_mid = (float)(pix_min_y + 2 * warp) + 0.5f;

    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t last_contributor = 0;
    int waited = 0;                         // batches whose barrier this thread has consumed

    for (int i = 0; i < rounds; ++i) {
        // everyone is past batch i-1 (its buffer may be refilled) -- and we may stop if all pixels are done
        if (__syncthreads_count(done) == TILE_PIX) break;
        if (i + 1 < rounds) prefetch(i + 1);
        mbar_wait(&s_bar[i & 1], (uint32_t)((i >> 1) & 1));
        waited = i + 1;
        const SplatRec* batch = s_rec[i & 1];
        const int nb = min(TILE_PIX, n_splats - i * TILE_PIX);
        // one checkpoint per 32-splat bucket, then the bucket's splats: the bucket loop keeps the checkpoint bookkeeping
        // out of the per-splat instruction stream
        for (int jb = 0; jb < nb && !done; jb += BUCKET) {
            if (!no_color) {
                ckpt[(size_t)bbm * TILE_PIX + tid] = make_float4(T, C0, C1, C2);
                ++bbm;
            }
            const int je = min(jb + BUCKET, nb);
            const SplatRec* sr = batch + jb;
            for (int j = jb; j < je; ++j, ++sr) {
                const float4 c = sr->c;
                const float4 a = sr->a;
                if (fabsf(row_mid - a.y) > c.w + 0.5f) continue;      // warp-uniform row-pair cull
                const float4 b = sr->b;
                const float dx = fsub(a.x, pfx), dy = fsub(a.y, pfy);
                const float power = splat_power(dx, dy, a.z, a.w, b.x);
                if (power > 0.0f) continue;
                const float alpha = fminf(0.99f, fmul(b.y, expf(power)));
                if (alpha < (1.0f / 255.0f)) continue;
                const float test_T = fmul(T, fsub(1.0f, alpha));
                if (test_T < 0.0001f) { done = true; break; }
                if (!no_color) {
                    C0 = ffma(T, fmul(alpha, b.z), C0);
                    C1 = ffma(T, fmul(alpha, b.w), C1);
                    C2 = ffma(T, fmul(alpha, c.x), C2);
                }
                T = test_T;
                last_contributor = (uint32_t)(i * TILE_PIX + j + 1);
            }
        }
    }
    // never leave the CTA with a bulk copy still in flight towards its shared memory
    {
        const int issued = min(rounds, waited + 1);
        if (issued > waited) mbar_wait(&s_bar[waited & 1], (uint32_t)((waited >> 1) & 1));
    }

    if (inside) {
        out_T[pid] = T;
        if (!no_color) {
            const size_t HW = (size_t)vp.W * vp.H;
            n_contrib[pid] = last_contributor;
            out_color[pid] = C0; out_color[HW + pid] = C1; out_color[2 * HW + pid] = C2;
            pixel_colors[pid] = C0; pixel_colors[HW + pid] = C1; pixel_colors[2 * HW + pid] = C2;
        }
    }
    if (no_color) return;
    uint32_t m = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) s_red[warp] = m;
    __syncthreads();
    if (tid == 0) {
        uint32_t mm = 0;
#pragma unroll
        for (int w = 0; w < TILE_PIX / 32; ++w) mm = max(mm, s_red[w]);
        max_contrib[tile] = mm;
    }
}

// ---- per-splat backward --------------------------------------------------------------------
// Only buckets that some pixel actually reached are worth a warp: tile t has ceil(max_contrib[t]/32) LIVE
// buckets (at cfg2: 19 k of 245 k).  live_scan_kernel prefix-sums them; the backward kernel is persistent
// (a fixed number of CTAs, warps stride over the live list), so every resident warp does useful work and no
// launch geometry depends on a device-side count.
__global__ void __launch_bounds__(1024)
live_scan_kernel(int T, const uint32_t* __restrict__ max_contrib, uint32_t* __restrict__ live_offsets, ImageHeader* hdr) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < T; base += 1024) {
        const int t = base + tid;
        const uint32_t v = t < T ? (max_contrib[t] + BUCKET - 1) / BUCKET : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_tot[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += n;
            }
            warp_tot[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t out = carry + warp_tot[warp] + incl;
        if (t < T) live_offsets[t] = out;
        __syncthreads();
        if (tid == 1023) carry = out;
        __syncthreads();
    }
    if (tid == 0) { hdr->num_live_buckets = carry; hdr->bwd_ticket = 0; }
}

constexpr int BWD_WARPS = 4;              // warps per CTA; every warp owns one (live bucket, tile part) at a time

// Backward of the blend, pixel-major inside a 32-splat bucket (replaces PerGaussianRenderCUDA, backward.cu:400-597).
//
// One warp per (LIVE bucket, tile part) in a persistent grid-stride loop; a part is RP consecutive row pairs of the 16x16
// tile (RP = 8: whole tile, RP = 4: half), and lane l owns the RP pixels (column l & 15, rows 2j + (l >> 4)) of its part.  Every pixel restarts from the bucket's checkpoint (T, C) and walks the bucket's splats in list order
// exactly like the forward, so T and the colour prefix need no hand-over between lanes and the eight pixels of a lane
// are eight independent dependency chains.  A (splat, row pair) is visited only if the splat's alpha >= 1/255 ellipse can
// reach those rows (the forward's conservative half-extent test => skipped pairs contribute exactly zero) and some pixel
// of the pair got this far.  The nine per-splat sums are accumulated over the lane's pixels in registers, combined across
// the warp by one reduce-scatter (8 values in 3 halving exchanges + 2 butterflies; the 9th by butterfly) and leave as
// 9 RED per splat and bucket -- the same global-atomic count as the reference's per-splat formulation, without its
// 287-step shuffle pipeline (4 shuffles + bookkeeping per pixel-splat pair).
template <int BWD_ROWPAIRS, int GRP, int MIN_CTAS>
__global__ void __launch_bounds__(BWD_WARPS * 32, MIN_CTAS)
render_backward_kernel(ViewParams vp, int T_tiles, ImageHeader* __restrict__ hdr, const uint2* __restrict__ ranges,
                       const uint32_t* __restrict__ point_list, const float4* __restrict__ rec,
                       const uint32_t* __restrict__ bucket_offsets, const uint32_t* __restrict__ live_offsets,
                       const float4* __restrict__ ckpt, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ pixel_colors, const float* __restrict__ dL_dpix,
                       float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity,
                       float* __restrict__ dL_dcolors) {
    __shared__ float4 s_sp[BWD_WARPS][BUCKET][3];            // the bucket's splat records, per warp
    __shared__ uint32_t s_gid[BWD_WARPS][BUCKET];
    __shared__ uint32_t s_rows[BWD_WARPS][BUCKET];           // row pairs each splat can reach (8-bit mask)

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_live = hdr->num_live_buckets;
    const size_t HW = (size_t)vp.W * vp.H;
    const float ddelx_dx = 0.5f * vp.W, ddely_dy = 0.5f * vp.H;
    float4 (*sp)[3] = s_sp[warp];

    // destination of this lane's reduce-scatter result: value slot v = lane >> 2 (held by lanes with (lane & 3) == 0)
    //   slot: 0 mean.x, 1 mean.y, 2 conic.x, 3 conic.y, 4 conic.w, 5..7 colour; lane 1 additionally carries opacity
    const int slot = lane >> 2;
    float* dst_base; int dst_stride, dst_off; float f_op, f_one;
    if (lane == 1) { dst_base = dL_dopacity; dst_stride = 1; dst_off = 0; f_op = 0.f; f_one = 1.f; }
    else if (slot < 2) { dst_base = dL_dmean2D; dst_stride = 3; dst_off = slot; f_op = slot == 0 ? -ddelx_dx : -ddely_dy; f_one = 0.f; }
    else if (slot < 5) { dst_base = dL_dconic; dst_stride = 4; dst_off = slot == 4 ? 3 : slot - 2; f_op = -0.5f; f_one = 0.f; }
    else { dst_base = dL_dcolors; dst_stride = 3; dst_off = slot - 5; f_op = 0.f; f_one = 1.f; }
    const bool i_write = (lane & 3) == 0 || lane == 1;

    constexpr int PARTS = (TILE / 2) / BWD_ROWPAIRS;
    // work items are handed out dynamically (one ticket per warp and item): buckets differ by more than 10x in cost
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(&hdr->bwd_ticket, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_live * PARTS) break;
        const uint32_t live = item / PARTS;
        const int part = (int)(item % PARTS);
        // tile of this live bucket: first t with live_offsets[t] > live -- 32-ary search, one probe per lane
        int lo = 0, cnt = T_tiles;
        while (cnt > 1) {
            const int step = (cnt + 31) >> 5;
            const int probe = lo + min((lane + 1) * step, cnt) - 1;              // last tile of this lane's sub-range
            const unsigned hit = __ballot_sync(0xffffffffu, live_offsets[probe] > live);
            const int first = __ffs(hit) - 1;                                      // hit != 0: live < live_offsets[T-1]
            lo += first * step;
            cnt = min(step, cnt - first * step);
        }
        const uint32_t tile = (uint32_t)lo;
        const int bucket_in_tile = (int)(live - (tile == 0 ? 0u : live_offsets[tile - 1]));
        const uint32_t bucket = (tile == 0 ? 0u : bucket_offsets[tile - 1]) + (uint32_t)bucket_in_tile;
        const uint2 range = ranges[tile];
        const int n_splats = (int)(range.y - range.x);
        const int bucket_start = bucket_in_tile * BUCKET;
        const int n_valid = min(BUCKET, n_splats - bucket_start);
        const int tile_x = tile % vp.grid_x, tile_y = tile / vp.grid_x;
        const int row0 = tile_y * TILE + 2 * BWD_ROWPAIRS * part;       // first tile row of this part
        const int qx = tile_x * TILE + (lane & (TILE - 1)), qy0 = row0 + (lane >> 4);
        const float pfx = (float)qx;

        // this lane's 8 pixels: contributors left inside this bucket (0 = never got here) and restart state
        int rel[BWD_ROWPAIRS];
        float Tr[BWD_ROWPAIRS], a0[BWD_ROWPAIRS], a1[BWD_ROWPAIRS], a2[BWD_ROWPAIRS];
        float g0[BWD_ROWPAIRS], g1[BWD_ROWPAIRS], g2[BWD_ROWPAIRS];
        int rel_max[BWD_ROWPAIRS];                            // warp-uniform: deepest pixel of each row pair
        int n_max = 0;
#pragma unroll
        for (int j = 0; j < BWD_ROWPAIRS; ++j) {
            const int qy = qy0 + 2 * j;
            int n = 0;
            size_t qi = 0;
            if (qx < vp.W && qy < vp.H) { qi = (size_t)qy * vp.W + qx; n = (int)n_contrib[qi]; }
            rel[j] = min(max(n - bucket_start, 0), BUCKET);
            Tr[j] = a0[j] = a1[j] = a2[j] = g0[j] = g1[j] = g2[j] = 0.f;
            if (rel[j] > 0) {
                const float4 k4 = ckpt[(size_t)bucket * TILE_PIX + (part * BWD_ROWPAIRS + j) * 32 + lane];
                Tr[j] = k4.x;
                a0[j] = k4.y - pixel_colors[qi]; a1[j] = k4.z - pixel_colors[HW + qi]; a2[j] = k4.w - pixel_colors[2 * HW + qi];
                g0[j] = dL_dpix[qi]; g1[j] = dL_dpix[HW + qi]; g2[j] = dL_dpix[2 * HW + qi];
            }
            rel_max[j] = (int)__reduce_max_sync(0xffffffffu, (unsigned)rel[j]);
            n_max = max(n_max, rel_max[j]);
        }

        // stage the bucket's splats (lane l <-> splat l) and the row pairs each one can reach
        {
            uint32_t gid = 0;
            float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
            uint32_t rows = 0;
            if (lane < n_valid && lane < n_max) {
                gid = point_list[range.x + bucket_start + lane];
                r0 = rec[3 * (size_t)gid + 0]; r1 = rec[3 * (size_t)gid + 1]; r2 = rec[3 * (size_t)gid + 2];
#pragma unroll
                for (int j = 0; j < BWD_ROWPAIRS; ++j) {
                    const float row_mid = (float)(row0 + 2 * j) + 0.5f;
                    if (lane < rel_max[j] && !(fabsf(row_mid - r0.y) > r2.w + 0.5f)) rows |= 1u << j;
                }
            }
            __syncwarp();                                    // previous bucket's readers are done
            sp[lane][0] = r0; sp[lane][1] = r1; sp[lane][2] = r2;
            s_gid[warp][lane] = gid;
            s_rows[warp][lane] = rows;
            __syncwarp();
        }

        uint32_t todo = __ballot_sync(0xffffffffu, s_rows[warp][lane] != 0);
        while (todo) {
            const int k = __ffs(todo) - 1;
            todo &= todo - 1;
            const float4 a = sp[k][0], b = sp[k][1];
            const float c2 = sp[k][2].x;
            const uint32_t rows = s_rows[warp][k];
            const float dx = fsub(a.x, pfx);
            const float cxdx = fmul(dx, a.z);                // shared by the lane's pixels: same column
            float v_mx = 0.f, v_my = 0.f, v_cx = 0.f, v_cy = 0.f, v_cw = 0.f, v_c0 = 0.f, v_c1 = 0.f, v_c2 = 0.f, v_o = 0.f;
            // Row pairs are taken in groups of GRP: one warp-uniform branch per group, and inside a group the pixels are
            // evaluated WITHOUT branches (a pixel that does not take the splat gets alpha = G = 0, which makes every term
            // below an exact zero and leaves T untouched), so the GRP dependency chains interleave in the issue stream.
#pragma unroll
            for (int jg = 0; jg < BWD_ROWPAIRS; jg += GRP) {
                if (!((rows >> jg) & ((1u << GRP) - 1u))) continue;              // warp-uniform
#pragma unroll
                for (int u = 0; u < GRP; ++u) {
                    const int j = jg + u;
                    const float dy = fsub(a.y, (float)(qy0 + 2 * j));
                    // splat_power(dx, dy, cx, cy, cz) with the dx*cx product hoisted (same operation order)
                    const float power = fsub(fmul(ffma(dx, cxdx, fmul(dy, fmul(dy, b.x))), -0.5f), fmul(dy, fmul(dx, a.w)));
                    const float Gx = expf(fminf(power, 0.0f));
                    const float ax = fminf(0.99f, fmul(b.y, Gx));
                    const bool take = (((rows >> j) & 1u) != 0) & (k < rel[j]) & (power <= 0.0f) & (ax >= (1.0f / 255.0f));
                    const float G = take ? Gx : 0.0f;
                    const float alpha = take ? ax : 0.0f;
                    const float one_m = fsub(1.0f, alpha);
                    const float T = Tr[j];
                    const float dchannel_dcolor = alpha * T;
                    float alpha_inverse;                                  // 1/(1-alpha), 1-alpha in [0.01, 1]: MUFU.RCP
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(alpha_inverse) : "f"(one_m));
                    float dL_dalpha;
                    a0[j] += dchannel_dcolor * b.z; v_c0 += dchannel_dcolor * g0[j]; dL_dalpha = ((b.z * T) + alpha_inverse * a0[j]) * g0[j];
                    a1[j] += dchannel_dcolor * b.w; v_c1 += dchannel_dcolor * g1[j]; dL_dalpha += ((b.w * T) + alpha_inverse * a1[j]) * g1[j];
                    a2[j] += dchannel_dcolor * c2; v_c2 += dchannel_dcolor * g2[j]; dL_dalpha += ((c2 * T) + alpha_inverse * a2[j]) * g2[j];
                    Tr[j] = fmul(T, one_m);
                    // constant factors (opacity, 0.5*W, 0.5*H, -0.5) are applied once per splat below
                    const float gdl = G * dL_dalpha;                      // = dL_dG / opacity
                    const float gdx = gdl * dx, gdy = gdl * dy;
                    v_mx += gdx * a.z + gdy * a.w;                        // -> * (-opacity * 0.5 * W)
                    v_my += gdy * b.x + gdx * a.w;                        // -> * (-opacity * 0.5 * H)
                    v_cx += gdx * dx;                                     // -> * (-0.5 * opacity)
                    v_cy += gdx * dy;
                    v_cw += gdy * dy;
                    v_o += gdl;
                }
            }
            // v_o != 0 iff some pixel of this lane contributed?  gdl can be exactly 0 (zero image gradient): then every
            // sum of the lane is 0 as well, so skipping is exact.
            if (!__any_sync(0xffffffffu, v_o != 0.f || v_c0 != 0.f || v_c1 != 0.f || v_c2 != 0.f)) continue;
            // reduce-scatter of 8 values over the 32 lanes: halve the value set at xor 16, 8, 4, then plain butterflies
            const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
            const float r0 = (h16 ? v_cw : v_mx) + __shfl_xor_sync(0xffffffffu, h16 ? v_mx : v_cw, 16);
            const float r1 = (h16 ? v_c0 : v_my) + __shfl_xor_sync(0xffffffffu, h16 ? v_my : v_c0, 16);
            const float r2 = (h16 ? v_c1 : v_cx) + __shfl_xor_sync(0xffffffffu, h16 ? v_cx : v_c1, 16);
            const float r3 = (h16 ? v_c2 : v_cy) + __shfl_xor_sync(0xffffffffu, h16 ? v_cy : v_c2, 16);
            const float q0 = (h8 ? r2 : r0) + __shfl_xor_sync(0xffffffffu, h8 ? r0 : r2, 8);
            const float q1 = (h8 ? r3 : r1) + __shfl_xor_sync(0xffffffffu, h8 ? r1 : r3, 8);
            float w0 = (h4 ? q1 : q0) + __shfl_xor_sync(0xffffffffu, h4 ? q0 : q1, 4);
            w0 += __shfl_xor_sync(0xffffffffu, w0, 2);
            w0 += __shfl_xor_sync(0xffffffffu, w0, 1);
            v_o += __shfl_xor_sync(0xffffffffu, v_o, 16);
            v_o += __shfl_xor_sync(0xffffffffu, v_o, 8);
            v_o += __shfl_xor_sync(0xffffffffu, v_o, 4);
            v_o += __shfl_xor_sync(0xffffffffu, v_o, 2);
            v_o += __shfl_xor_sync(0xffffffffu, v_o, 1);
            if (i_write) {
                const size_t gid = s_gid[warp][k];
                const float val = lane == 1 ? v_o : w0;
                atomicAdd(dst_base + dst_stride * gid + dst_off, val * (f_op * b.y + f_one));
            }
        }
    }
}

// ---- launchers ------------------------------------------------------------------------------
int launch_tile_ranges(int64_t R, const GeomHeader* ghdr, const uint32_t* keys_sorted, int T, ImageState img, bool buckets,
                       cudaStream_t s) {
    const unsigned int* r_dev = ghdr ? &ghdr->r_eff : nullptr;
    GLIC_CUDA_TRY(cudaMemsetAsync(img.ranges, 0, sizeof(uint2) * (size_t)T, s));
    if (R > 0) {
        tile_ranges_kernel<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(R, r_dev, keys_sorted, (uint32_t)T, img.ranges);
        GLIC_LAUNCH_CHECK();
    }
    bucket_scan_kernel<<<1, 1024, 0, s>>>(T, img.ranges, img.bucket_offsets, img.hdr, (long long)R, ghdr, buckets ? 1 : 0);
    GLIC_LAUNCH_CHECK();
    return GLIC_OK;
}

int launch_render_forward(const ViewParams& vp, bool no_color, const uint32_t* point_list, GeomState g, ImageState img,
                          SampleState smp, float* out_color, float* out_final_T, cudaStream_t s) {
    dim3 grid(vp.grid_x, vp.grid_y), block(TILE, TILE);
    render_forward_kernel<<<grid, block, 0, s>>>(vp, no_color, img.ranges, point_list, g.rec, img.bucket_offsets,
                                                 smp.bucket_to_tile, smp.ckpt,
                                                 img.n_contrib, img.max_contrib, img.pixel_colors, out_color,
                                                 out_final_T);
    GLIC_LAUNCH_CHECK();
    return GLIC_OK;
}

int launch_render_backward(int P, const ViewParams& vp, int64_t max_buckets, const uint32_t* point_list, GeomState g,
                           ImageState img, SampleState smp, const float* dL_dpix, float* dL_dmean2D, float* dL_dconic,
                           float* dL_dopacity, float* dL_dcolors, cudaStream_t s) {
    (void)P; (void)smp.bucket_to_tile;
    if (max_buckets <= 0) return GLIC_OK;
    const int T = vp.grid_x * vp.grid_y;
    live_scan_kernel<<<1, 1024, 0, s>>>(T, img.max_contrib, img.live_offsets, img.hdr);
    GLIC_LAUNCH_CHECK();
    // RP = row pairs per lane (8: one warp per bucket, 4: two warps share a bucket's tile); GRP = row pairs evaluated
    // branch-free together.  Tuning knobs for the profiles/ sweep; the default is the measured best.
    static const int rp = getenv("GLIC_BWD_RP") ? atoi(getenv("GLIC_BWD_RP")) : 8;
    static const int grp = getenv("GLIC_BWD_GRP") ? atoi(getenv("GLIC_BWD_GRP")) : 2;
    using Kern = void (*)(ViewParams, int, ImageHeader*, const uint2*, const uint32_t*, const float4*, const uint32_t*,
                          const uint32_t*, const float4*, const uint32_t*, const float*, const float*, float*, float*, float*, float*);
    static Kern kern = nullptr;
    static int blocks = 0;
    if (!kern) {
        kern = rp == 8 ? (grp >= 4 ? render_backward_kernel<8, 4, 3> : grp == 2 ? render_backward_kernel<8, 2, 4> : render_backward_kernel<8, 1, 4>)
                       : (grp >= 4 ? render_backward_kernel<4, 4, 5> : grp == 2 ? render_backward_kernel<4, 2, 6> : render_backward_kernel<4, 1, 6>);
        int dev = 0, sms = 148, per_sm = 4;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BWD_WARPS * 32, 0);
        blocks = sms * (per_sm > 0 ? per_sm : 1);
    }
    const int parts = rp == 8 ? 1 : 2;
    const int64_t need = (max_buckets * parts + BWD_WARPS - 1) / BWD_WARPS;
    kern<<<(unsigned)std::min<int64_t>(blocks, need), BWD_WARPS * 32, 0, s>>>(
        vp, T, img.hdr, img.ranges, point_list, g.rec, img.bucket_offsets, img.live_offsets, smp.ckpt, img.n_contrib,
        img.pixel_colors, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors);
    GLIC_LAUNCH_CHECK();
    return GLIC_OK;
}

}  // namespace glic
