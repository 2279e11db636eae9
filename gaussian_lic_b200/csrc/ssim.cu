// ssim.cu -- fused 11x11 separable-Gaussian SSIM map (+ partial-derivative maps), its backward,
// and the whole L1 + D-SSIM photometric loss of one mapping iteration.
//
// Replaces fusedssimCUDA / fusedssim_backwardCUDA (reference fused-ssim/ssim.cu:186-365) and, for
// glic_l1_ssim_loss, the chain l1_loss + fused_ssim + loss-combine + their autograd backward
// (gaussian.cpp:685-697, loss_utils.h:30-33,135-193).
//
// Design: one CTA = one 32x32 output tile of one channel.  Both 42x42 halo tiles are loaded once;
// the horizontal pass produces all five moment rows (x, y, xx, yy, xy) into shared memory in one
// sweep and the vertical pass finishes them: 3 barriers per tile instead of the reference's 18 per
// channel, and no scratch-buffer re-zeroing.  Zero ("same") padding as the reference.
#include "common.cuh"

namespace glic {

namespace {

constexpr int SB = 32;            // output tile edge
constexpr int HALO = 5;
constexpr int SH_ = SB + 2 * HALO;  // 42

__device__ __constant__ float kG[11] = {0.001028380123898387f, 0.0075987582094967365f, 0.036000773310661316f,
                                        0.10936068743467331f,  0.21300552785396576f,   0.26601171493530273f,
                                        0.21300552785396576f,  0.10936068743467331f,   0.036000773310661316f,
                                        0.0075987582094967365f, 0.001028380123898387f};

__device__ __forceinline__ float pix_or_zero(const float* __restrict__ img, int y, int x, int H, int W) {
    return (x >= 0 && y >= 0 && x < W && y < H) ? img[(size_t)y * W + x] : 0.0f;
}

// LOSS = true: additionally accumulates (1-lambda)*|a-b| - lambda*ssim, scaled by 1/N, into *loss.
template <bool LOSS>
__global__ void __launch_bounds__(SB * SB)
ssim_forward_kernel(int H, int W, float C1, float C2, const float* __restrict__ img1, const float* __restrict__ img2,
                    float* __restrict__ ssim_map, float* __restrict__ dm_dmu1, float* __restrict__ dm_dsigma1_sq,
                    float* __restrict__ dm_dsigma12, float lambda_dssim, float inv_n, float* __restrict__ loss) {
    __shared__ float s1[SH_][SH_ + 1];
    __shared__ float s2[SH_][SH_ + 1];
    __shared__ float h[5][SH_][SB];
    __shared__ float red[SB];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * SB + tx;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const float* a = img1 + plane;
    const float* b = img2 + plane;
    const int x0 = blockIdx.x * SB, y0 = blockIdx.y * SB;

    for (int i = tid; i < SH_ * SH_; i += SB * SB) {
        const int ly = i / SH_, lx = i % SH_;
        s1[ly][lx] = pix_or_zero(a, y0 + ly - HALO, x0 + lx - HALO, H, W);
        s2[ly][lx] = pix_or_zero(b, y0 + ly - HALO, x0 + lx - HALO, H, W);
    }
    __syncthreads();
    for (int i = tid; i < SH_ * SB; i += SB * SB) {
        const int ly = i / SB, lx = i % SB;
        float m1 = 0.f, m2 = 0.f, q11 = 0.f, q22 = 0.f, q12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float p = s1[ly][lx + k], q = s2[ly][lx + k], g = kG[k];
            m1 = fmaf(g, p, m1); m2 = fmaf(g, q, m2);
            q11 = fmaf(g, p * p, q11); q22 = fmaf(g, q * q, q22); q12 = fmaf(g, p * q, q12);
        }
        h[0][ly][lx] = m1; h[1][ly][lx] = m2; h[2][ly][lx] = q11; h[3][ly][lx] = q22; h[4][ly][lx] = q12;
    }
    __syncthreads();
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
        const float g = kG[k];
        mu1 = fmaf(g, h[0][ty + k][tx], mu1); mu2 = fmaf(g, h[1][ty + k][tx], mu2);
        e11 = fmaf(g, h[2][ty + k][tx], e11); e22 = fmaf(g, h[3][ty + k][tx], e22);
        e12 = fmaf(g, h[4][ty + k][tx], e12);
    }
    const float sigma1_sq = e11 - mu1 * mu1, sigma2_sq = e22 - mu2 * mu2, sigma12 = e12 - mu1 * mu2;
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
    const float Cq = 2.0f * mu1_mu2 + C1, Dq = 2.0f * sigma12 + C2;
    const float A = mu1_sq + mu2_sq + C1, B = sigma1_sq + sigma2_sq + C2;
    const float m = (Cq * Dq) / (A * B);
    const int px = x0 + tx, py = y0 + ty;
    const bool inside = px < W && py < H;
    if (inside) {
        const size_t gi = plane + (size_t)py * W + px;
        if (ssim_map) ssim_map[gi] = m;
        if (dm_dmu1) {
            dm_dmu1[gi] = ((mu2 * 2.0f * Dq) / (A * B) - (mu2 * 2.0f * Cq) / (A * B) - (mu1 * 2.0f * Cq * Dq) / (A * A * B) +
                           (mu1 * 2.0f * Cq * Dq) / (A * B * B));
            dm_dsigma1_sq[gi] = ((-Cq * Dq) / (A * B * B));
            dm_dsigma12[gi] = ((2.0f * Cq) / (A * B));
        }
    }
    if (LOSS) {
        float part = 0.f;
        if (inside) part = (1.0f - lambda_dssim) * fabsf(s1[ty + HALO][tx + HALO] - s2[ty + HALO][tx + HALO]) - lambda_dssim * m;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (tx == 0) red[ty] = part;
        __syncthreads();
        if (ty == 0) {
            float v = red[tx];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (tx == 0) {
                if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) v = fmaf(v, inv_n, lambda_dssim);
                else v *= inv_n;
                atomicAdd(loss, v);
            }
        }
    }
}

// dL/dimg1 = conv(dL*dm_dmu1) + 2*img1*conv(dL*dm_dsigma1_sq) + img2*conv(dL*dm_dsigma12)
// CONST_DL: dL_dmap is the constant `dl_const` (fused loss) and the L1 sign term is added.
template <bool CONST_DL>
__global__ void __launch_bounds__(SB * SB)
ssim_backward_kernel(int H, int W, const float* __restrict__ img1, const float* __restrict__ img2,
                     const float* __restrict__ dL_dmap, const float* __restrict__ dm_dmu1,
                     const float* __restrict__ dm_dsigma1_sq, const float* __restrict__ dm_dsigma12,
                     float* __restrict__ dL_dimg1, float dl_const, float l1_scale) {
    __shared__ float s[3][SH_][SH_ + 1];
    __shared__ float h[3][SH_][SB];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * SB + tx;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * SB, y0 = blockIdx.y * SB;
    for (int i = tid; i < SH_ * SH_; i += SB * SB) {
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
    for (int i = tid; i < SH_ * SB; i += SB * SB) {
        const int ly = i / SB, lx = i % SB;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float g = kG[k];
            a0 = fmaf(g, s[0][ly][lx + k], a0); a1 = fmaf(g, s[1][ly][lx + k], a1); a2 = fmaf(g, s[2][ly][lx + k], a2);
        }
        h[0][ly][lx] = a0; h[1][ly][lx] = a1; h[2][ly][lx] = a2;
    }
    __syncthreads();
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
        const float g = kG[k];
        c0 = fmaf(g, h[0][ty + k][tx], c0); c1 = fmaf(g, h[1][ty + k][tx], c1); c2 = fmaf(g, h[2][ty + k][tx], c2);
    }
    const int px = x0 + tx, py = y0 + ty;
    if (px < W && py < H) {
        const size_t gi = plane + (size_t)py * W + px;
        const float p1 = img1[gi], p2 = img2[gi];
        float out = c0;
        out += p1 * 2.0f * c1;
        out += p2 * c2;
        if (CONST_DL) {
            const float d = p1 - p2;
            out += l1_scale * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        }
        dL_dimg1[gi] = out;
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
    dim3 grid((W + SB - 1) / SB, (H + SB - 1) / SB, B * CH), block(SB, SB);
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
    dim3 grid((W + SB - 1) / SB, (H + SB - 1) / SB, B * CH), block(SB, SB);
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
    dim3 grid((W + SB - 1) / SB, (H + SB - 1) / SB, CH), block(SB, SB);
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
