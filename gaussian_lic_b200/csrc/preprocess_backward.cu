// preprocess_backward.cu -- per-Gaussian backward: 2-D gradients -> dL/d{mean3D, cov3D, dc, sh,
// scale, rotation}.
//
// Replaces computeCov2DCUDA + preprocessCUDA(bwd) + computeColorFromSH(bwd) + computeCov3D(bwd)
// (reference backward.cu:27-377) in ONE kernel, and the ten torch::zeros of
// rasterize_points.cu:192-201: every output row is written here (exact zeros for culled
// Gaussians).  Forward intermediates (cov3D, T) are recomputed from scale/rotation instead of
// being stored (24 B/Gaussian less state, and the inputs are read anyway).
#include "geom_math.cuh"
#include "warp_rows.cuh"

namespace glic {

__device__ __constant__ float bSH_C1 = 0.4886025119029199f;
__device__ __constant__ float bSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                           -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float bSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                           0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                           -0.5900435899266435f};
constexpr float bSH_C0 = 0.28209479177387814f;

constexpr int PB_THREADS = 128;

__global__ void __launch_bounds__(PB_THREADS)
preprocess_backward_kernel(int P, int D, int M, const float* __restrict__ means, const float* __restrict__ scales,
                           float mod, const float4* __restrict__ rots, const float* __restrict__ sh, ViewParams vp,
                           const int* __restrict__ radii, const uint8_t* __restrict__ clamped, float lambda_erank,
                           const float* __restrict__ dL_dmean2D, const float* __restrict__ dL_dconic,
                           const float* __restrict__ dL_dcolors, float* __restrict__ dL_dmeans3D,
                           float* __restrict__ dL_dcov3D, float* __restrict__ dL_ddc, float* __restrict__ dL_dsh,
                           float* __restrict__ dL_dscales, float4* __restrict__ dL_drots) {
    __shared__ float s_view[16], s_proj[16], s_cam[3];
    __shared__ __align__(16) float s_sh[PB_THREADS / 32][32 * SH_ROW_MAX];    // SH rows in, dL/dSH rows out
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 16) s_view[tid] = vp.view[tid];
    else if (tid < 32) s_proj[tid - 16] = vp.proj[tid - 16];
    else if (tid < 35) s_cam[tid - 32] = vp.campos[tid - 32];
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + tid;
    const int K = 3 * M;
    const bool staged = K > 0 && K <= SH_ROW_MAX && sh != nullptr;
    const int wfirst = blockIdx.x * blockDim.x + warp * 32;
    const int wcnt = min(32, P - wfirst);
    if (wcnt <= 0) return;                                         // whole warp beyond P (warp-uniform)
    float* slab = s_sh[warp];
    if (staged && D > 0) warp_load_rows(sh, (size_t)wfirst, wcnt, K, slab, lane);
    // dL/dSH row of this thread: shared-memory slab (written back coalesced below) or global memory
    float* dsh = staged ? (slab + lane * K) : (dL_dsh + (size_t)(idx < P ? idx : 0) * K);
    const bool visible = idx < P && radii[idx] > 0;
    if (!visible) {
        if (idx < P) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { dL_dmeans3D[3 * idx + k] = 0.f; dL_ddc[3 * idx + k] = 0.f; dL_dscales[3 * idx + k] = 0.f; }
#pragma unroll
            for (int k = 0; k < 6; ++k) dL_dcov3D[6 * idx + k] = 0.f;
            dL_drots[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (idx < P || staged) for (int k = 0; k < K; ++k) dsh[k] = 0.f;
    } else {
    const float px = means[3 * idx], py = means[3 * idx + 1], pz = means[3 * idx + 2];
    const float s0 = scales[3 * idx], s1 = scales[3 * idx + 1], s2 = scales[3 * idx + 2];
    const float4 q = rots[idx];
    Cov3 c3;
    cov3d_from_scale_rot(s0, s1, s2, mod, q, c3);
    Cov2 c2;
    cov2d_project(px, py, pz, s_view, vp.focal_x, vp.focal_y, vp.limx_neg, vp.limx_pos, vp.limy_neg, vp.limy_pos, c3.c, c2);
    const float fx = vp.focal_x, fy = vp.focal_y;
    const float* v = s_view;

    // ---- EWA backward (backward.cu:138-255) -------------------------------------------------
    const float x_grad_mul = (c2.txtz < vp.limx_neg || c2.txtz > vp.limx_pos) ? 0.f : 1.f;
    const float y_grad_mul = (c2.tytz < vp.limy_neg || c2.tytz > vp.limy_pos) ? 0.f : 1.f;
    const float tz = c2.tz, tcx = c2.lx * tz, tcy = c2.ly * tz;
    const float a = c2.a, b = c2.b, c = c2.c;
    const float gxx = dL_dconic[4 * idx], gxy = dL_dconic[4 * idx + 1], gyy = dL_dconic[4 * idx + 3];
    const float denom = a * c - b * b;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
    const float T00 = c2.T00, T01 = c2.T01, T02 = c2.T02, T10 = c2.T10, T11 = c2.T11, T12 = c2.T12;
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (denom2inv != 0.f) {
        dL_da = denom2inv * (-c * c * gxx + 2 * b * c * gxy + (denom - a * c) * gyy);
        dL_dc = denom2inv * (-a * a * gyy + 2 * a * b * gxy + (denom - a * c) * gxx);
        dL_db = denom2inv * 2 * (b * c * gxx - (denom + 2 * b * b) * gxy + a * b * gyy);
        dcov[0] = (T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc);
        dcov[3] = (T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc);
        dcov[5] = (T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc);
        dcov[1] = 2 * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2 * T10 * T11 * dL_dc;
        dcov[2] = 2 * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2 * T10 * T12 * dL_dc;
        dcov[4] = 2 * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2 * T11 * T12 * dL_dc;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) dL_dcov3D[6 * idx + k] = dcov[k];
    const float* V = c3.c;
    // rows of T^T Vrk products: r0k = sum_m T0m * Vrk[k][m], r1k = sum_m T1m * Vrk[k][m]
    const float r00 = T00 * V[0] + T01 * V[1] + T02 * V[2], r10 = T10 * V[0] + T11 * V[1] + T12 * V[2];
    const float r01 = T00 * V[1] + T01 * V[3] + T02 * V[4], r11 = T10 * V[1] + T11 * V[3] + T12 * V[4];
    const float r02 = T00 * V[2] + T01 * V[4] + T02 * V[5], r12 = T10 * V[2] + T11 * V[4] + T12 * V[5];
    const float dT00 = 2 * r00 * dL_da + r10 * dL_db, dT01 = 2 * r01 * dL_da + r11 * dL_db, dT02 = 2 * r02 * dL_da + r12 * dL_db;
    const float dT10 = 2 * r10 * dL_dc + r00 * dL_db, dT11 = 2 * r11 * dL_dc + r01 * dL_db, dT12 = 2 * r12 * dL_dc + r02 * dL_db;
    const float dJ00 = v[0] * dT00 + v[4] * dT01 + v[8] * dT02;
    const float dJ02 = v[2] * dT00 + v[6] * dT01 + v[10] * dT02;
    const float dJ11 = v[1] * dT10 + v[5] * dT11 + v[9] * dT12;
    const float dJ12 = v[2] * dT10 + v[6] * dT11 + v[10] * dT12;
    const float itz = 1.f / tz, itz2 = itz * itz, itz3 = itz2 * itz;
    const float dtx = x_grad_mul * -fx * itz2 * dJ02;
    const float dty = y_grad_mul * -fy * itz2 * dJ12;
    const float dtz = -fx * itz2 * dJ00 - fy * itz2 * dJ11 + (2 * fx * tcx) * itz3 * dJ02 + (2 * fy * tcy) * itz3 * dJ12;
    float dmx = v[0] * dtx + v[1] * dty + v[2] * dtz;
    float dmy = v[4] * dtx + v[5] * dty + v[6] * dtz;
    float dmz = v[8] * dtx + v[9] * dty + v[10] * dtz;

    // ---- projection term (backward.cu:339-350) ------------------------------------------------
    {
        const float* pr = s_proj;
        const float hx = xform_row(pr, 0, px, py, pz), hy = xform_row(pr, 1, px, py, pz), hw = xform_row(pr, 3, px, py, pz);
        const float m_w = 1.0f / (hw + 0.0000001f);
        const float mul1 = hx * m_w * m_w, mul2 = hy * m_w * m_w;
        const float g2x = dL_dmean2D[3 * idx], g2y = dL_dmean2D[3 * idx + 1];
        dmx += (pr[0] * m_w - pr[3] * mul1) * g2x + (pr[1] * m_w - pr[3] * mul2) * g2y;
        dmy += (pr[4] * m_w - pr[7] * mul1) * g2x + (pr[5] * m_w - pr[7] * mul2) * g2y;
        dmz += (pr[8] * m_w - pr[11] * mul1) * g2x + (pr[9] * m_w - pr[11] * mul2) * g2y;
    }

    // ---- SH backward (backward.cu:27-136).  Like the reference (`if (shs)`, backward.cu:352) the whole colour
    // backward -- dL/ddc included -- is skipped when no SH-rest tensor is bound (M == 0). -------------------------
    if (sh == nullptr) {
        dL_ddc[3 * idx] = 0.f; dL_ddc[3 * idx + 1] = 0.f; dL_ddc[3 * idx + 2] = 0.f;
    } else {
        const float ox = px - s_cam[0], oy = py - s_cam[1], oz = pz - s_cam[2];
        const float sum2 = ox * ox + oy * oy + oz * oz;
        const float inv = 1.0f / sqrtf(sum2);
        const float x = ox * inv, y = oy * inv, z = oz * inv;
        const unsigned cb = clamped[idx];
        float dRGB[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) dRGB[ch] = (cb >> ch) & 1u ? 0.f : dL_dcolors[3 * idx + ch];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) dL_ddc[3 * idx + ch] = bSH_C0 * dRGB[ch];
        float ddx = 0.f, ddy = 0.f, ddz = 0.f;   // dL/d(dir)
        if (D > 0) {
            float sv[SH_ROW_MAX];                                  // input row -> registers (the slab row is reused for the output)
            if (staged) {
#pragma unroll
                for (int k = 0; k < SH_ROW_MAX; ++k) sv[k] = k < K ? slab[lane * K + k] : 0.f;
            } else {
#pragma unroll
                for (int k = 0; k < SH_ROW_MAX; ++k) sv[k] = k < K ? sh[(size_t)idx * K + k] : 0.f;
            }
            const float* s = sv;   // rows are lane-private: no cross-lane hazard between this read and the dsh writes
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            float bs[15];
            bs[0] = -bSH_C1 * y; bs[1] = bSH_C1 * z; bs[2] = -bSH_C1 * x;
            int nb = 3;
            if (D > 1) {
                bs[3] = bSH_C2[0] * xy; bs[4] = bSH_C2[1] * yz; bs[5] = bSH_C2[2] * (2.f * zz - xx - yy);
                bs[6] = bSH_C2[3] * xz; bs[7] = bSH_C2[4] * (xx - yy);
                nb = 8;
                if (D > 2) {
                    bs[8] = bSH_C3[0] * y * (3.f * xx - yy);
                    bs[9] = bSH_C3[1] * xy * z;
                    bs[10] = bSH_C3[2] * y * (4.f * zz - xx - yy);
                    bs[11] = bSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                    bs[12] = bSH_C3[4] * x * (4.f * zz - xx - yy);
                    bs[13] = bSH_C3[5] * z * (xx - yy);
                    bs[14] = bSH_C3[6] * x * (xx - 3.f * yy);
                    nb = 15;
                }
            }
#pragma unroll
            for (int k = 0; k < 15; ++k) {
                if (k < nb) {
                    dsh[3 * k + 0] = bs[k] * dRGB[0];
                    dsh[3 * k + 1] = bs[k] * dRGB[1];
                    dsh[3 * k + 2] = bs[k] * dRGB[2];
                }
            }
            for (int k = nb; k < M; ++k) { dsh[3 * k] = 0.f; dsh[3 * k + 1] = 0.f; dsh[3 * k + 2] = 0.f; }
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
#define SHc(k) s[(k) * 3 + ch]
                float dx_ = -bSH_C1 * SHc(2), dy_ = -bSH_C1 * SHc(0), dz_ = bSH_C1 * SHc(1);
                if (D > 1) {
                    dx_ += bSH_C2[0] * y * SHc(3) + bSH_C2[2] * 2.f * -x * SHc(5) + bSH_C2[3] * z * SHc(6) + bSH_C2[4] * 2.f * x * SHc(7);
                    dy_ += bSH_C2[0] * x * SHc(3) + bSH_C2[1] * z * SHc(4) + bSH_C2[2] * 2.f * -y * SHc(5) + bSH_C2[4] * 2.f * -y * SHc(7);
                    dz_ += bSH_C2[1] * y * SHc(4) + bSH_C2[2] * 2.f * 2.f * z * SHc(5) + bSH_C2[3] * x * SHc(6);
                    if (D > 2) {
                        dx_ += bSH_C3[0] * SHc(8) * 3.f * 2.f * xy + bSH_C3[1] * SHc(9) * yz + bSH_C3[2] * SHc(10) * -2.f * xy +
                               bSH_C3[3] * SHc(11) * -3.f * 2.f * xz + bSH_C3[4] * SHc(12) * (-3.f * xx + 4.f * zz - yy) +
                               bSH_C3[5] * SHc(13) * 2.f * xz + bSH_C3[6] * SHc(14) * 3.f * (xx - yy);
                        dy_ += bSH_C3[0] * SHc(8) * 3.f * (xx - yy) + bSH_C3[1] * SHc(9) * xz +
                               bSH_C3[2] * SHc(10) * (-3.f * yy + 4.f * zz - xx) + bSH_C3[3] * SHc(11) * -3.f * 2.f * yz +
                               bSH_C3[4] * SHc(12) * -2.f * xy + bSH_C3[5] * SHc(13) * -2.f * yz + bSH_C3[6] * SHc(14) * -3.f * 2.f * xy;
                        dz_ += bSH_C3[1] * SHc(9) * xy + bSH_C3[2] * SHc(10) * 4.f * 2.f * yz +
                               bSH_C3[3] * SHc(11) * 3.f * (2.f * zz - xx - yy) + bSH_C3[4] * SHc(12) * 4.f * 2.f * xz +
                               bSH_C3[5] * SHc(13) * (xx - yy);
                    }
                }
#undef SHc
                ddx += dx_ * dRGB[ch]; ddy += dy_ * dRGB[ch]; ddz += dz_ * dRGB[ch];
            }
        } else {
            for (int k = 0; k < 3 * M; ++k) dsh[k] = 0.f;
        }
        // d normalize(v)/dv applied to dL/d(dir)  (auxiliary.h:103-114)
        const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
        dmx += ((+sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * invsum32;
        dmy += (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * invsum32;
        dmz += (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * invsum32;
    }
    dL_dmeans3D[3 * idx] = dmx; dL_dmeans3D[3 * idx + 1] = dmy; dL_dmeans3D[3 * idx + 2] = dmz;

    // ---- cov3D backward (backward.cu:257-310) ------------------------------------------------------
    {
        const float sv[3] = {mod * s0, mod * s1, mod * s2};
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        const float dS[9] = {dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], 0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
                             0.5f * dcov[2], 0.5f * dcov[4], dcov[5]};
        float dMt[9];   // dMt[3j+k] = dL_dM[k][j] with dL_dM = 2*M*dL_dSigma (column-major sense)
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                dMt[3 * j + k] = 2.f * c3.M[j] * dS[3 * k] + 2.f * c3.M[3 + j] * dS[3 * k + 1] + 2.f * c3.M[6 + j] * dS[3 * k + 2];
        float ds[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) ds[j] = c3.R[j] * dMt[3 * j] + c3.R[3 + j] * dMt[3 * j + 1] + c3.R[6 + j] * dMt[3 * j + 2];
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) dMt[3 * j + k] *= sv[j];
#define DMT(a_, b_) dMt[3 * (a_) + (b_)]
        float4 dq;
        dq.x = 2 * z * (DMT(0, 1) - DMT(1, 0)) + 2 * y * (DMT(2, 0) - DMT(0, 2)) + 2 * x * (DMT(1, 2) - DMT(2, 1));
        dq.y = 2 * y * (DMT(0, 1) + DMT(1, 0)) + 2 * z * (DMT(2, 0) + DMT(0, 2)) + 2 * r * (DMT(1, 2) - DMT(2, 1)) - 4 * x * (DMT(2, 2) + DMT(1, 1));
        dq.z = 2 * x * (DMT(0, 1) + DMT(1, 0)) + 2 * r * (DMT(2, 0) - DMT(0, 2)) + 2 * z * (DMT(1, 2) + DMT(2, 1)) - 4 * y * (DMT(2, 2) + DMT(0, 0));
        dq.w = 2 * r * (DMT(0, 1) - DMT(1, 0)) + 2 * x * (DMT(2, 0) + DMT(0, 2)) + 2 * y * (DMT(1, 2) + DMT(2, 1)) - 4 * z * (DMT(1, 1) + DMT(0, 0));
#undef DMT
        if (lambda_erank > 0.f) {   // effective-rank regulariser, backward.cu:358-375 (off in every shipped config)
            const float s1s1 = s0 * s0, s2s2 = s1 * s1, s3s3 = s2 * s2, sum = s1s1 + s2s2 + s3s3;
            const float q1 = s0 / sum, q2 = s1 / sum, q3 = s2 / sum;
            const float erank = expf(-q1 * logf(q1) - q2 * logf(q2) - q3 * logf(q3));
            if (-log((double)erank - 1 + 1e-5) > 0) {
                const float f = (float)(erank / (erank - 1 + 1e-5));
                const float e1 = f * (-logf(q1) - 1), e2 = f * (-logf(q2) - 1), e3 = f * (-logf(q3) - 1);
                const float le = lambda_erank * 2.f / (sum * sum);
                ds[0] += le * s0 * (e1 * (s2s2 + s3s3) - e2 * s2s2 - e3 * s3s3);
                ds[1] += le * s1 * (-e1 * s1s1 + e2 * (s1s1 + s3s3) - e3 * s3s3);
                ds[2] += le * s2 * (-e1 * s1s1 - e2 * s2s2 + e3 * (s1s1 + s2s2));
            }
            ds[2] += 1.f;           // backward.cu:374 adds this unconditionally; kept for parity
        }
        dL_dscales[3 * idx] = ds[0]; dL_dscales[3 * idx + 1] = ds[1]; dL_dscales[3 * idx + 2] = ds[2];
        dL_drots[idx] = dq;
    }
    }   // visible
    if (staged) warp_store_rows(dL_dsh, (size_t)wfirst, wcnt, K, slab, lane);
}

int launch_preprocess_backward(int P, int D, int M, const float* means, const float* scales, float mod,
                               const float* rots, const float* sh, const ViewParams& vp, const int* radii, GeomState g,
                               float lambda_erank, const float* dL_dmean2D, const float* dL_dconic,
                               const float* dL_dcolors, float* dL_dmeans3D, float* dL_dcov3D, float* dL_ddc,
                               float* dL_dsh, float* dL_dscales, float* dL_drots, cudaStream_t s) {
    preprocess_backward_kernel<<<(P + PB_THREADS - 1) / PB_THREADS, PB_THREADS, 0, s>>>(
        P, D, M, means, scales, mod, reinterpret_cast<const float4*>(rots), sh, vp, radii, g.clamped, lambda_erank,
        dL_dmean2D, dL_dconic, dL_dcolors, dL_dmeans3D, dL_dcov3D, dL_ddc, dL_dsh, dL_dscales,
        reinterpret_cast<float4*>(dL_drots));
    GLIC_LAUNCH_CHECK();
    return GLIC_OK;
}

}  // namespace glic
