// model_step.cu -- the optimiser side of a mapping iteration on the packed model.
//
// The reference keeps six parameter tensors, derives the rasterizer inputs with three elementwise torch ops
// (GaussianModel::getOpacity / getScaling / getRotation, gaussian.cpp:147-175: sigmoid, exp, normalize), lets autograd
// run their three backward ops, clones every gradient and launches adamUpdateCUDA six times (optim_utils.h:102-137).
// Here the model lives in ONE planar buffer in the layout of the packed gradient buffer
//     rotation[4P] | xyz[3P] | log-scale[3P] | opacity logit[P] | dc[3P] | sh-rest[3MP]
// so that: the activations are one kernel (8 floats in, 8 out per Gaussian), their chain rule is applied IN PLACE on
// the gradient buffer the backward just wrote, and the visibility-masked Adam of all six groups is a single launch
// that consumes that buffer (and, multi-GPU, exactly the bytes the exchange step just reduced).
#include "common.cuh"
#include "adam_math.cuh"
#include <algorithm>

namespace glic {
namespace {

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// torch::sigmoid / torch::exp / torch::nn::functional::normalize (p = 2, eps = 1e-12) of the raw parameters
__global__ void __launch_bounds__(256)
activations_forward_kernel(int P, const float* __restrict__ opacity_logit, const float* __restrict__ log_scale,
                           const float4* __restrict__ rot_raw, float* __restrict__ opacity, float* __restrict__ scale,
                           float4* __restrict__ rot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    opacity[i] = 1.0f / (1.0f + expf(-opacity_logit[i]));
#pragma unroll
    for (int c = 0; c < 3; ++c) scale[3 * (size_t)i + c] = expf(log_scale[3 * (size_t)i + c]);
    const float4 q = rot_raw[i];
    const float n = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
    rot[i] = make_float4(q.x / n, q.y / n, q.z / n, q.w / n);
}

// dL/d(activated) -> dL/d(raw), in place: sigmoid' = s(1-s); exp' = exp; normalize: (g - q (q.g)) / |r| (|r| > eps)
__global__ void __launch_bounds__(256)
activations_backward_kernel(int P, const float* __restrict__ opacity, const float* __restrict__ scale,
                            const float4* __restrict__ rot_raw, float* __restrict__ dL_dopacity, float* __restrict__ dL_dscale,
                            float4* __restrict__ dL_drot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float s = opacity[i];
    dL_dopacity[i] = dL_dopacity[i] * (s * (1.0f - s));
#pragma unroll
    for (int c = 0; c < 3; ++c) dL_dscale[3 * (size_t)i + c] *= scale[3 * (size_t)i + c];
    const float4 r = rot_raw[i];
    const float4 g = dL_drot[i];
    const float nn = sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w);
    if (nn > 1e-12f) {
        const float inv = 1.0f / nn;
        const float4 q = make_float4(r.x * inv, r.y * inv, r.z * inv, r.w * inv);
        const float qg = q.x * g.x + q.y * g.y + q.z * g.z + q.w * g.w;
        dL_drot[i] = make_float4((g.x - q.x * qg) * inv, (g.y - q.y * qg) * inv, (g.z - q.z * qg) * inv, (g.w - q.w * qg) * inv);
    } else {
        const float inv = 1.0f / 1e-12f;                    // clamped branch of normalize: plain scaling
        dL_drot[i] = make_float4(g.x * inv, g.y * inv, g.z * inv, g.w * inv);
    }
}

struct PackedLayout {
    size_t end[6];       // exclusive end offset (in floats) of each group inside the planar buffer
    size_t begin[6];
    uint32_t k[6];       // floats per Gaussian
    float lr[6];
};

// same arithmetic, element by element, as adam_kernel (adamUpdateCUDA, adam.cu:9-38): only the launch count changes
__global__ void __launch_bounds__(256)
adam_packed_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ exp_avg,
                   float* __restrict__ exp_avg_sq, const uint8_t* __restrict__ visible, PackedLayout L, float b1, float b2,
                   float eps, size_t total) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (size_t)gridDim.x * blockDim.x) {
        int grp = 0;
#pragma unroll
        for (int q = 0; q < 5; ++q) grp += j >= L.end[q];
        const size_t gaussian = (j - L.begin[grp]) / L.k[grp];
        if (!visible[gaussian]) continue;
        float p = param[j], m = exp_avg[j], v = exp_avg_sq[j];
        adam_element(p, m, v, grad[j], L.lr[grp], b1, b2, eps);
        param[j] = p;
        exp_avg[j] = m;
        exp_avg_sq[j] = v;
    }
}

}  // namespace
}  // namespace glic

using namespace glic;

extern "C" {

size_t glic_packed_floats(uint32_t P, uint32_t M) { return (size_t)P * (4 + 3 + 3 + 1 + 3 + 3 * (size_t)M); }

int glic_packed_offsets(uint32_t P, uint32_t M, size_t* offsets6) {
    if (!offsets6) { set_error("packed_offsets: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    const uint32_t k[6] = {4, 3, 3, 1, 3, 3 * M};
    size_t off = 0;
    for (int q = 0; q < 6; ++q) { offsets6[q] = off; off += (size_t)P * k[q]; }
    return GLIC_OK;
}

int glic_activations_forward(int P, const float* opacity_logit, const float* log_scale, const float* rot_raw, float* opacity,
                             float* scale, float* rot, void* stream) {
    if (P < 0) { set_error("activations_forward: bad P"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (P == 0) return GLIC_OK;
    if (!opacity_logit || !log_scale || !rot_raw || !opacity || !scale || !rot) { set_error("activations_forward: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (!aligned16(rot_raw) || !aligned16(rot)) { set_error("activations_forward: rotations must be 16-byte aligned"); return GLIC_ERR_INVALID_ARGUMENT; }
    StageTimer _t(GLIC_STAGE_ADAM, (cudaStream_t)stream);
    activations_forward_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        P, opacity_logit, log_scale, reinterpret_cast<const float4*>(rot_raw), opacity, scale, reinterpret_cast<float4*>(rot));
    GLIC_LAUNCH_CHECK();
    return GLIC_OK;
}

int glic_activations_backward(int P, const float* opacity, const float* scale, const float* rot_raw, float* dL_dopacity,
                              float* dL_dscale, float* dL_drot, void* stream) {
    if (P < 0) { set_error("activations_backward: bad P"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (P == 0) return GLIC_OK;
    if (!opacity || !scale || !rot_raw || !dL_dopacity || !dL_dscale || !dL_drot) { set_error("activations_backward: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (!aligned16(rot_raw) || !aligned16(dL_drot)) { set_error("activations_backward: rotations must be 16-byte aligned"); return GLIC_ERR_INVALID_ARGUMENT; }
    StageTimer _t(GLIC_STAGE_ADAM, (cudaStream_t)stream);
    activations_backward_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        P, opacity, scale, reinterpret_cast<const float4*>(rot_raw), dL_dopacity, dL_dscale, reinterpret_cast<float4*>(dL_drot));
    GLIC_LAUNCH_CHECK();
    return GLIC_OK;
}

int glic_adam_update_packed(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const uint8_t* visible,
                            const float* lr6_host, float b1, float b2, float eps, uint32_t P, uint32_t M, void* stream) {
    if (P == 0) return GLIC_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq || !visible || !lr6_host) { set_error("adam_update_packed: null pointer"); return GLIC_ERR_INVALID_ARGUMENT; }
    PackedLayout L;
    const uint32_t k[6] = {4, 3, 3, 1, 3, 3 * M};
    size_t off = 0;
    for (int q = 0; q < 6; ++q) {
        L.begin[q] = off; L.k[q] = k[q] ? k[q] : 1; L.lr[q] = lr6_host[q];
        off += (size_t)P * k[q];
        L.end[q] = off;
    }
    const size_t total = off;
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)148 * 16);
    StageTimer _t(GLIC_STAGE_ADAM, (cudaStream_t)stream);
    adam_packed_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, visible, L, b1, b2, eps, total);
    GLIC_LAUNCH_CHECK();
    return GLIC_OK;
}

}  // extern "C"
