// ssim.cu -- fused 11x11 separable-Gaussian SSIM map (+ partial-derivative maps), its backward,
// and the whole L1 + D-SSIM photometric loss of one mapping iteration.
//
// Replaces fusedssimCUDA / fusedssim_backwardCUDA (reference fused-ssim/ssim.cu:186-365) and, for
// glic_l1_ssim_loss, the chain l1_loss + fused_ssim + loss-combine + their autograd backward
// (gaussian.cpp:685-697, loss_utils.h:30-33,135-193).
//
// Design: one CTA (256 threads) = one 32x32 output tile of one channel.  Both 42x42 halo tiles are loaded
// once; the horizontal pass produces all five moment rows (x, y, xx, yy, xy) into shared memory in one
// register-tiled sweep (4 outputs per item) and the vertical pass finishes 4 rows per thread: 3 barriers
// per tile instead of the reference's 18 per channel, ~3x fewer shared-memory loads than one-output-per-
// thread, no scratch-buffer re-zeroing.  Zero ("same") padding as the reference.
// The convolutions are pure multiply-add streams, so moments travel in PAIRS through packed fp32x2 instructions
// (f32x2.cuh) in the FORWARD kernel: (x, y) and (xx, yy) share one FFMA2 per tap, xy keeps a scalar FFMA -- 3 issue slots per
// tap instead of 5 -- and the pairs sit interleaved in shared memory so one 64-bit load fetches both (-8 % measured).  Each
// half is the same fmaf(g, v, acc) chain as before: bit-identical results.  The backward (3 maps) measured SLOWER packed
// (+9..16 %: 64-bit shared-memory traffic costs more than the saved issue slots) and stays scalar.
#include "common.cuh"
#include "f32x2.cuh"
#include <algorithm>

namespace glic {

namespace {

constexpr int SB = 32;              // output tile edge
constexpr int HALO = 5;
constexpr int SH_ = SB + 2 * HALO;  // 42
constexpr int ST = 256;             // threads per CTA: 32 columns x 8 row groups, 4 outputs per thread
constexpr int RPT = 4;              // rows (vertical pass) / columns (horizontal pass) per thread

__device__ __constant__ float kG[11] = {0.001028380123898387f, 0.0075987582094967365f, 0.036000773310661316f,
                                        0.10936068743467331f,  0.21300552785396576f,   0.26601171493530273f,
                                        0.21300552785396576f,  0.10936068743467331f,   0.036000773310661316f,
                                        0.0075987582094967365f, 0.001028380123898387f};

__device__ __forceinline__ float pix_or_zero(const float* __restrict__ img, int y, int x, int H, int W) {
    return (x >= 0 && y >= 0 && x < W && y < H) ? img[(size_t)y * W + x] : 0.0f;
}

// Register-tiled separable 11-tap convolution.  Horizontal pass: one work item = 4 consecutive outputs
// of one halo row (14 loads per image instead of 44); vertical pass: one thread = 4 consecutive rows of
// one column (14 loads per moment instead of 44).  Row strides 43 / 33 keep both passes bank-conflict free.
// LOSS = true: additionally accumulates (1-lambda)*|a-b| - lambda*ssim, scaled by 1/N, into *loss.
template <bool LOSS>
__global__ void __launch_bounds__(ST)
ssim_forward_kernel(int H, int W, float C1, float C2, const float* __restrict__ img1, const float* __restrict__ img2,
                    float* __restrict__ ssim_map, float* __restrict__ dm_dmu1, float* __restrict__ dm_dsigma1_sq,
                    float* __restrict__ dm_dsigma12, float lambda_dssim, float inv_n, float* __restrict__ loss) {
    __shared__ float s1[SH_][SH_ + 1];
    __shared__ float s2[SH_][SH_ + 1];
    __shared__ float2 h01[SH_][SB + 1];          // horizontal pass of (x, y)
    __shared__ float2 h23[SH_][SB + 1];          // horizontal pass of (xx, yy)
    __shared__ float h4[SH_][SB + 1];            // horizontal pass of xy
    __shared__ float red[ST / 32];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const float* a = img1 + plane;
    const float* b = img2 + plane;
    const int x0 = blockIdx.x * SB, y0 = blockIdx.y * SB;

    for (int i = tid; i < SH_ * SH_; i += ST) {
        const int ly = i / SH_, lx = i % SH_;
        s1[ly][lx] = pix_or_zero(a, y0 + ly - HALO, x0 + lx - HALO, H, W);
        s2[ly][lx] = pix_or_zero(b, y0 + ly - HALO, x0 + lx - HALO, H, W);
    }
    __syncthreads();
    for (int i = tid; i < SH_ * (SB / RPT); i += ST) {
        const int ly = i % SH_, lx = (i / SH_) * RPT;     // consecutive lanes = consecutive rows: conflict-free for the 64-bit pair arrays
        f2 pq[RPT + 10], sq[RPT + 10];
        float xy[RPT + 10];
#pragma unroll
        for (int k = 0; k < RPT + 10; ++k) {
            const float pp = s1[ly][lx + k], qq = s2[ly][lx + k];
            pq[k] = f2_pack(pp, qq);
            sq[k] = f2_mul(pq[k], pq[k]);
            xy[k] = pp * qq;
        }
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            f2 m = f2_bcast(0.f), q = m;
            float q12 = 0.f;
#pragma unroll
            for (int k = 0; k < 11; ++k) {
                const float g = kG[k];
                m = f2_fma(pq[j + k], g, m);
                q = f2_fma(sq[j + k], g, q);
                q12 = fmaf(g, xy[j + k], q12);
            }
            float lo, hi;
            f2_unpack(m, lo, hi); h01[ly][lx + j] = make_float2(lo, hi);
            f2_unpack(q, lo, hi); h23[ly][lx + j] = make_float2(lo, hi);
            h4[ly][lx + j] = q12;
        }
    }
    __syncthreads();
    float v[5][RPT];
    {
        f2 c01[RPT + 10], c23[RPT + 10];
        float c4[RPT + 10];
#pragma unroll
        for (int k = 0; k < RPT + 10; ++k) {
            const float2 u = h01[ty * RPT + k][tx], w = h23[ty * RPT + k][tx];
            c01[k] = f2_pack(u.x, u.y); c23[k] = f2_pack(w.x, w.y);
            c4[k] = h4[ty * RPT + k][tx];
        }
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            f2 m = f2_bcast(0.f), q = m;
            float q12 = 0.f;
#pragma unroll
            for (int k = 0; k < 11; ++k) {
                const float g = kG[k];
                m = f2_fma(c01[j + k], g, m);
                q = f2_fma(c23[j + k], g, q);
                q12 = fmaf(g, c4[j + k], q12);
            }
            f2_unpack(m, v[0][j], v[1][j]);
            f2_unpack(q, v[2][j], v[3][j]);
            v[4][j] = q12;
        }
    }
    float part = 0.f;
    const int px = x0 + tx;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const float mu1 = v[0][j], mu2 = v[1][j];
        const float sigma1_sq = v[2][j] - mu1 * mu1, sigma2_sq = v[3][j] - mu2 * mu2, sigma12 = v[4][j] - mu1 * mu2;
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
        const float Cq = 2.0f * mu1_mu2 + C1, Dq = 2.0f * sigma12 + C2;
        const float A = mu1_sq + mu2_sq + C1, B = sigma1_sq + sigma2_sq + C2;
        const float m = (Cq * Dq) / (A * B);
        const int ry = ty * RPT + j, py = y0 + ry;
        if (px < W && py < H) {
            const size_t gi = plane + (size_t)py * W + px;
            if (ssim_map) ssim_map[gi] = m;
            if (dm_dmu1) {
                dm_dmu1[gi] = ((mu2 * 2.0f * Dq) / (A * B) - (mu2 * 2.0f * Cq) / (A * B) - (mu1 * 2.0f * Cq * Dq) / (A * A * B) +
                               (mu1 * 2.0f * Cq * Dq) / (A * B * B));
                dm_dsigma1_sq[gi] = ((-Cq * Dq) / (A * B * B));
                dm_dsigma12[gi] = ((2.0f * Cq) / (A * B));
            }
            if (LOSS) part += (1.0f - lambda_dssim) * fabsf(s1[ry + HALO][tx + HALO] - s2[ry + HALO][tx + HALO]) - lambda_dssim * m;
        }
    }
    if (LOSS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (tx == 0) red[ty] = part;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < ST / 32; ++w) t += red[w];
            if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) t = fmaf(t, inv_n, lambda_dssim);
            else t *= inv_n;
            atomicAdd(loss, t);
        }
    }
}

// dL/dimg1 = conv(dL*dm_dmu1) + 2*img1*conv(dL*dm_dsigma1_sq) + img2*conv(dL*dm_dsigma12)
// CONST_DL: dL_dmap is the constant `dl_const` (fused loss) and the L1 sign term is added.
template <bool CONST_DL>
__global__ void __launch_bounds__(ST)
ssim_backward_kernel(int H, int W, const float* __restrict__ img1, const float* __restrict__ img2,
                     const float* __restrict__ dL_dmap, const float* __restrict__ dm_dmu1,
                     const float* __restrict__ dm_dsigma1_sq, const float* __restrict__ dm_dsigma12,
                     float* __restrict__ dL_dimg1, float dl_const, float l1_scale) {
    __shared__ float s[3][SH_][SH_ + 1];
    __shared__ float h[3][SH_][SB + 1];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * SB, y0 = blockIdx.y * SB;
    for (int i = tid; i < SH_ * SH_; i += ST) {
        const int ly = i / SH_, lx = i % SH_;
        const int y = y0 + ly - HALO, x = x0 + lx - HALO;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
        if (x >= 0 && y >= 0 && x < W && y < H) {
            const size_t gi = plane + (size_t)y * W + x;
            const float dl = CONST_DL ? dl_const : dL_dmap[gi];
            v0 = dm_dmu1[gi] * dl; v1 = dm_dsigma1_sq[gi] * dl; v2 = dm_dsigma12[gi] * dl;
        }
        s[0][ly][lx] = v0; s[1][ly][lx] = v1; s[2][ly][lx] = v2;
    }
    __syncthreads();
    for (int i = tid; i < SH_ * (SB / RPT); i += ST) {
        const int ly = i / (SB / RPT), lx = (i % (SB / RPT)) * RPT;
#pragma unroll
        for (int qn = 0; qn < 3; ++qn) {
            float p[RPT + 10];
#pragma unroll
            for (int k = 0; k < RPT + 10; ++k) p[k] = s[qn][ly][lx + k];
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 11; ++k) acc = fmaf(kG[k], p[j + k], acc);
                h[qn][ly][lx + j] = acc;
            }
        }
    }
    __syncthreads();
    float c[3][RPT];
#pragma unroll
    for (int qn = 0; qn < 3; ++qn) {
        float col[RPT + 10];
#pragma unroll
        for (int k = 0; k < RPT + 10; ++k) col[k] = h[qn][ty * RPT + k][tx];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 11; ++k) acc = fmaf(kG[k], col[j + k], acc);
            c[qn][j] = acc;
        }
    }
    const int px = x0 + tx;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int py = y0 + ty * RPT + j;
        if (px < W && py < H) {
            const size_t gi = plane + (size_t)py * W + px;
            const float p1 = img1[gi], p2 = img2[gi];
            float out = c[0][j];
            out += p1 * 2.0f * c[1][j];
            out += p2 * c[2][j];
            if (CONST_DL) {
                const float d = p1 - p2;
                out += l1_scale * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
            }
            dL_dimg1[gi] = out;
        }
    }
}

}  // namespace

}  // namespace glic

using namespace glic;

extern "C" int glic_fused_ssim(int B, int CH, int H, int W, float C1, float C2, const float* img1, const float* img2,
                               float* ssim_map, float* dm_dmu1, float* dm_dsigma1_sq, float* dm_dsigma12, void* stream) {
    if (B < 0 || CH < 0 || H <= 0 || W <= 0 || !img1 || !img2 || !ssim_map) { set_error("fused_ssim: bad arguments"); return GLIC_ERR_INVALID_ARGUMENT; }
    if ((dm_dmu1 != nullptr) != (dm_dsigma1_sq != nullptr) || (dm_dmu1 != nullptr) != (dm_dsigma12 != nullptr)) {
        set_error("fused_ssim: partial maps must be all NULL or all non-NULL"); return GLIC_ERR_INVALID_ARGUMENT;
    }
    if (B * CH == 0) return GLIC_OK;
    dim3 grid((W + SB - 1) / SB, (H + SB - 1) / SB, B * CH), block(ST);
    ssim_forward_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(H, W, C1, C2, img1, img2, ssim_map, dm_dmu1,
                                                                          dm_dsigma1_sq, dm_dsigma12, 0.f, 0.f, nullptr);
    GLIC_LAUNCH_CHECK();
    return GLIC_OK;
}

extern "C" int glic_fused_ssim_backward(int B, int CH, int H, int W, float C1, float C2, const float* img1,
                                        const float* img2, const float* dL_dmap, const float* dm_dmu1,
                                        const float* dm_dsigma1_sq, const float* dm_dsigma12, float* dL_dimg1, void* stream) {
    (void)C1; (void)C2;
    if (H <= 0 || W <= 0 || !img1 || !img2 || !dL_dmap || !dm_dmu1 || !dm_dsigma1_sq || !dm_dsigma12 || !dL_dimg1) {
        set_error("fused_ssim_backward: bad arguments"); return GLIC_ERR_INVALID_ARGUMENT;
    }
    if (B * CH == 0) return GLIC_OK;
    dim3 grid((W + SB - 1) / SB, (H + SB - 1) / SB, B * CH), block(ST);
    ssim_backward_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(H, W, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq,
                                                                           dm_dsigma12, dL_dimg1, 0.f, 0.f);
    GLIC_LAUNCH_CHECK();
    return GLIC_OK;
}

extern "C" size_t glic_loss_scratch_bytes(int CH, int H, int W) { return sizeof(float) * 3 * (size_t)CH * H * W + 256; }

extern "C" int glic_l1_ssim_loss(int CH, int H, int W, float lambda_dssim, const float* img, const float* gt, float* loss_out,
                                 float* dL_dimg, void* scratch, size_t scratch_bytes, void* stream) {
    if (CH <= 0 || H <= 0 || W <= 0 || !img || !gt || !loss_out || !scratch) { set_error("l1_ssim_loss: bad arguments"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (scratch_bytes < glic_loss_scratch_bytes(CH, H, W)) { set_error("l1_ssim_loss: scratch too small"); return GLIC_ERR_WORKSPACE; }
    const size_t N = (size_t)CH * H * W;
    float* d1 = static_cast<float*>(scratch);
    float* d2 = d1 + N;
    float* d3 = d2 + N;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;     // loss_utils.h:130-131
    const float inv_n = (float)(1.0 / (double)N);
    cudaStream_t s = (cudaStream_t)stream;
    GLIC_CUDA_TRY(cudaMemsetAsync(loss_out, 0, sizeof(float), s));
    dim3 grid((W + SB - 1) / SB, (H + SB - 1) / SB, CH), block(ST);
    { StageTimer _t(GLIC_STAGE_LOSS_FWD, s);
    ssim_forward_kernel<true><<<grid, block, 0, s>>>(H, W, C1, C2, img, gt, nullptr, d1, d2, d3, lambda_dssim, inv_n, loss_out); }
    GLIC_LAUNCH_CHECK();
    if (dL_dimg) {
        StageTimer _t(GLIC_STAGE_LOSS_BWD, s);
        ssim_backward_kernel<true><<<grid, block, 0, s>>>(H, W, img, gt, nullptr, d1, d2, d3, dL_dimg, -lambda_dssim * inv_n,
                                                          (1.0f - lambda_dssim) * inv_n);
        GLIC_LAUNCH_CHECK();
    }
    return GLIC_OK;
}


// ---- evaluation metrics (evaluateVisualQuality, gaussian.cpp:756-760,795-799; loss_utils.h:35-39,84-127) --------------
namespace glic {
namespace {

// clamp both images to [0,1] (gaussian.cpp:756-757) into scratch copies and accumulate the squared error
__global__ void __launch_bounds__(256)
eval_clamp_mse_kernel(size_t N, const float* __restrict__ img, const float* __restrict__ gt, float* __restrict__ a,
                      float* __restrict__ b, double* __restrict__ acc2) {
    double se = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) {
        const float x = fminf(fmaxf(img[i], 0.f), 1.f), y = fminf(fmaxf(gt[i], 0.f), 1.f);
        a[i] = x; b[i] = y;
        const float d = x - y;
        se += (double)(d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    if ((threadIdx.x & 31) == 0 && se != 0.0) atomicAdd(&acc2[0], se);
}

__global__ void __launch_bounds__(256)
eval_sum_kernel(size_t N, const float* __restrict__ map, double* __restrict__ acc2) {
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) s += (double)map[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&acc2[1], s);
}

__global__ void eval_finish_kernel(size_t N, const double* __restrict__ acc2, float* __restrict__ out2) {
    const float mse = (float)(acc2[0] / (double)N);
    out2[0] = 10.0f * log10f(1.0f / mse);          // loss_utils.h:35-39
    out2[1] = (float)(acc2[1] / (double)N);        // ssim_map.mean(), loss_utils.h:107
}

}  // namespace
}  // namespace glic

extern "C" size_t glic_eval_scratch_bytes(int CH, int H, int W) { return sizeof(float) * 3 * (size_t)CH * H * W + 256; }

extern "C" int glic_eval_psnr_ssim(int CH, int H, int W, const float* img, const float* gt, float* out2, void* scratch,
                                   size_t scratch_bytes, void* stream) {
    if (CH <= 0 || H <= 0 || W <= 0 || !img || !gt || !out2 || !scratch) { set_error("eval_psnr_ssim: bad arguments"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (scratch_bytes < glic_eval_scratch_bytes(CH, H, W)) { set_error("eval_psnr_ssim: scratch too small"); return GLIC_ERR_WORKSPACE; }
    const size_t N = (size_t)CH * H * W;
    cudaStream_t s = (cudaStream_t)stream;
    double* acc2 = static_cast<double*>(scratch);                     // first 256 bytes: the two accumulators
    float* a = reinterpret_cast<float*>(static_cast<char*>(scratch) + 256);
    float* b = a + N;
    float* map = b + N;
    GLIC_CUDA_TRY(cudaMemsetAsync(acc2, 0, 2 * sizeof(double), s));
    const unsigned blocks = (unsigned)std::min<size_t>((N + 255) / 256, (size_t)148 * 8);
    eval_clamp_mse_kernel<<<blocks, 256, 0, s>>>(N, img, gt, a, b, acc2);
    GLIC_LAUNCH_CHECK();
    if (int e = glic_fused_ssim(1, CH, H, W, 0.01f * 0.01f, 0.03f * 0.03f, a, b, map, nullptr, nullptr, nullptr, stream)) return e;
    eval_sum_kernel<<<blocks, 256, 0, s>>>(N, map, acc2);
    GLIC_LAUNCH_CHECK();
    eval_finish_kernel<<<1, 1, 0, s>>>(N, acc2, out2);
    GLIC_LAUNCH_CHECK();
    return GLIC_OK;
}
