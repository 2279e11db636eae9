// sh_math.cuh -- real spherical-harmonics basis (degree <= 3), its gradient, and the rotation-matrix -> quaternion
// chain rule, shared by the per-Gaussian backward kernels and the SH-rebuilding Adam kernel.
//
// Same functions as the reference's computeColorFromSH backward (backward.cu:27-136) and computeCov3D backward
// (backward.cu:257-310) -- the maths is fixed by the parity contract -- but derived and factored here on their own:
//   colour_c(dir) = sum_k b_k(dir) * sh[k][c]   with b_k a polynomial in the unit direction (x, y, z)
//   dL/dsh[k][c]  = b_k * g_c                        (g = clamp-masked dL/dcolour)
//   dL/ddir       = sum_k grad b_k * w_k,   w_k = sum_c sh[k][c] * g_c
// i.e. ONE table of basis values and ONE table of basis gradients per direction, contracted with a per-coefficient
// scalar w_k; the reference instead expands the derivative channel by channel.  The two agree to fp32 rounding.
#pragma once
#include <cuda_runtime.h>

namespace glic {

constexpr float kSH0 = 0.28209479177387814f;
constexpr float kSH1 = 0.4886025119029199f;
constexpr float kSH2a = 1.0925484305920792f, kSH2b = 0.31539156525252005f, kSH2c = 0.5462742152960396f;
constexpr float kSH3a = 0.5900435899266435f, kSH3b = 2.890611442640554f, kSH3c = 0.4570457994644658f,
                kSH3d = 0.3731763325901154f, kSH3e = 1.445305721320277f;

// number of SH-rest basis functions of degree D (1..3): 3, 8, 15
__host__ __device__ __forceinline__ int sh_rest_count(int D) { return (D + 1) * (D + 1) - 1; }

// b[k], k = 0..sh_rest_count(D)-1: the degree >= 1 basis at unit direction (x, y, z).  The products are formed exactly as
// in the forward's colour evaluation (forward.cu:29-77), so dL/dsh = b_k * g reproduces the reference's bits.
template <int MAXK = 15>
__device__ __forceinline__ void sh_basis(int D, float x, float y, float z, float (&b)[MAXK]) {
    b[0] = -kSH1 * y; b[1] = kSH1 * z; b[2] = -kSH1 * x;
    if (D < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[3] = kSH2a * xy; b[4] = -kSH2a * yz; b[5] = kSH2b * (2.f * zz - xx - yy); b[6] = -kSH2a * xz; b[7] = kSH2c * (xx - yy);
    if (D < 3) return;
    b[8] = -kSH3a * y * (3.f * xx - yy);
    b[9] = kSH3b * xy * z;
    b[10] = -kSH3c * y * (4.f * zz - xx - yy);
    b[11] = kSH3d * z * (2.f * zz - 3.f * xx - 3.f * yy);
    b[12] = -kSH3c * x * (4.f * zz - xx - yy);
    b[13] = kSH3e * z * (xx - yy);
    b[14] = -kSH3a * x * (xx - 3.f * yy);
}

// dL/ddir = sum_k w[k] * grad b_k(x, y, z); grad b_k written out per degree from the polynomials above.
__device__ __forceinline__ void sh_dir_gradient(int D, float x, float y, float z, const float* w, float& gx, float& gy, float& gz) {
    // degree 1: b0 = -c1 y, b1 = c1 z, b2 = -c1 x
    gx = -kSH1 * w[2]; gy = -kSH1 * w[0]; gz = kSH1 * w[1];
    if (D < 2) return;
    // degree 2: a xy | -a yz | b (2zz - xx - yy) | -a xz | c (xx - yy)
    const float a3 = kSH2a * w[3], a4 = -kSH2a * w[4], a5 = kSH2b * w[5], a6 = -kSH2a * w[6], a7 = kSH2c * w[7];
    gx += a3 * y - 2.f * a5 * x + a6 * z + 2.f * a7 * x;
    gy += a3 * x + a4 * z - 2.f * a5 * y - 2.f * a7 * y;
    gz += a4 * y + 4.f * a5 * z + a6 * x;
    if (D < 3) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    const float c8 = -kSH3a * w[8], c9 = kSH3b * w[9], c10 = -kSH3c * w[10], c11 = kSH3d * w[11], c12 = -kSH3c * w[12],
                c13 = kSH3e * w[13], c14 = -kSH3a * w[14];
    // y(3xx - yy) | xyz | y(4zz - xx - yy) | z(2zz - 3xx - 3yy) | x(4zz - xx - yy) | z(xx - yy) | x(xx - 3yy)
    gx += c8 * 6.f * xy + c9 * yz - c10 * 2.f * xy - c11 * 6.f * xz + c12 * (4.f * zz - 3.f * xx - yy) + c13 * 2.f * xz + c14 * 3.f * (xx - yy);
    gy += c8 * 3.f * (xx - yy) + c9 * xz + c10 * (4.f * zz - xx - 3.f * yy) - c11 * 6.f * yz - c12 * 2.f * xy - c13 * 2.f * yz - c14 * 6.f * xy;
    gz += c9 * xy + c10 * 8.f * yz + c11 * 3.f * (2.f * zz - xx - yy) + c12 * 8.f * xz + c13 * (xx - yy);
}

// Chain rule through R(q), q = (r, x, y, z) unit quaternion:
//   R = [1-2(yy+zz)  2(xy-rz)    2(xz+ry) ]
//       [2(xy+rz)    1-2(xx+zz)  2(yz-rx) ]
//       [2(xz-ry)    2(yz+rx)    1-2(xx+yy)]
// G[3i+j] = dL/dR_ij  ->  dL/dq.  Antisymmetric parts of G drive r, symmetric parts the vector components.
__device__ __forceinline__ float4 quat_gradient(float4 q, const float* G) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    const float s01 = G[1] + G[3], s02 = G[2] + G[6], s12 = G[5] + G[7];       // symmetric sums
    const float a10 = G[3] - G[1], a02 = G[2] - G[6], a21 = G[7] - G[5];       // antisymmetric differences
    float4 d;
    d.x = 2.f * (z * a10 + y * a02 + x * a21);
    d.y = 2.f * (y * s01 + z * s02 + r * a21) - 4.f * x * (G[4] + G[8]);
    d.z = 2.f * (x * s01 + r * a02 + z * s12) - 4.f * y * (G[0] + G[8]);
    d.w = 2.f * (r * a10 + x * s02 + y * s12) - 4.f * z * (G[0] + G[4]);
    return d;
}

}  // namespace glic
