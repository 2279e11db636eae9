// geom_math.cuh -- per-Gaussian projective geometry with a FIXED fp32 operation order.
//
// The integer part of the parity contract (radii, tiles_touched, the sorted (tile|depth) list,
// tile ranges) hangs on the exact bits of depth, xy, conic and radius.  nvcc is free to contract
// a*b+c into fma differently from one build to the next, so every operation below is an explicit
// round-to-nearest intrinsic (__fmaf_rn/__fmul_rn/__fadd_rn/__fdiv_rn/__frcp_rn/__fsqrt_rn): the
// compiler may neither fuse nor reassociate them.  The order chosen is the one the reference
// build evaluates (DESIGN.md "contraction table"; maths: SURVEY.md App. A.2-A.3, reference
// forward.cu:79-149,232-319, forward.h:34-78, auxiliary.h:41-90).
#pragma once
#include "common.cuh"

namespace glic {

#define GLIC_DI __device__ __forceinline__

GLIC_DI float fmul(float a, float b) { return __fmul_rn(a, b); }
GLIC_DI float fadd(float a, float b) { return __fadd_rn(a, b); }
GLIC_DI float fsub(float a, float b) { return __fsub_rn(a, b); }
GLIC_DI float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// a0*b0 + a1*b1 + a2*b2 evaluated as fma(a2,b2, fma(a0,b0, a1*b1))
GLIC_DI float dot3c(float a0, float b0, float a1, float b1, float a2, float b2) {
    return ffma(a2, b2, ffma(a0, b0, fmul(a1, b1)));
}

// row `r` of a column-major 4x4 (m[4c+r]) applied to (x,y,z,1): m[12+r] + fma(z,m[8+r], fma(x,m[r], y*m[4+r]))
GLIC_DI float xform_row(const float* __restrict__ m, int r, float x, float y, float z) {
    return fadd(m[12 + r], ffma(z, m[8 + r], ffma(x, m[r], fmul(y, m[4 + r]))));
}

struct Cov3 {
    float c[6];      // Sigma: [00,01,02,11,12,22]
    float M[9];      // M[3j+i] = s_i * R[j][i]   (column-major glm sense)
    float R[9];      // R[3j+i]
};

// Sigma = (S R)^T (S R), quaternion (r,x,y,z) already normalised, scale already activated.
GLIC_DI void cov3d_from_scale_rot(float sx, float sy, float sz, float mod, float4 q, Cov3& o) {
    sx = fmul(mod, sx); sy = fmul(mod, sy); sz = fmul(mod, sz);
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    // products kept as separate roundings: yy, zz, xz, rx, rz; xy, ry, yz are fused into the sums
    // (this is what ptxas emits for the reference: SASS of preprocessCUDA)
    const float yy = fmul(y, y), zz = fmul(z, z);
    const float xz = fmul(x, z), rx = fmul(r, x), rz = fmul(r, z);
    float t;
    t = fadd(yy, zz);               o.R[0] = fsub(1.0f, fadd(t, t));
    t = ffma(x, y, -rz);            o.R[1] = fadd(t, t);
    t = ffma(r, y, xz);             o.R[2] = fadd(t, t);
    t = ffma(x, y, rz);             o.R[3] = fadd(t, t);
    t = ffma(x, x, zz);             o.R[4] = fsub(1.0f, fadd(t, t));
    t = ffma(y, z, -rx);            o.R[5] = fadd(t, t);
    t = ffma(-r, y, xz);            o.R[6] = fadd(t, t);
    t = ffma(y, z, rx);             o.R[7] = fadd(t, t);
    t = ffma(x, x, yy);             o.R[8] = fsub(1.0f, fadd(t, t));
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        o.M[3 * j + 0] = fmul(sx, o.R[3 * j + 0]);
        o.M[3 * j + 1] = fmul(sy, o.R[3 * j + 1]);
        o.M[3 * j + 2] = fmul(sz, o.R[3 * j + 2]);
    }
    const float* M = o.M;
    o.c[0] = dot3c(M[0], M[0], M[1], M[1], M[2], M[2]);
    o.c[1] = dot3c(M[3], M[0], M[4], M[1], M[5], M[2]);
    o.c[2] = dot3c(M[6], M[0], M[7], M[1], M[8], M[2]);
    o.c[3] = dot3c(M[3], M[3], M[4], M[4], M[5], M[5]);
    o.c[4] = dot3c(M[6], M[3], M[7], M[4], M[8], M[5]);
    o.c[5] = dot3c(M[6], M[6], M[7], M[7], M[8], M[8]);
}

struct Cov2 {
    float tz;                 // depth (p_view.z)
    float txtz, tytz;         // unclamped ratios
    float lx, ly;             // clamped ratios
    float T00, T01, T02, T10, T11, T12;
    float a, b, c;            // EWA covariance with the +0.3 dilation
};

// EWA projection of Sigma (forward.cu:79-118).  view = column-major Rt in shared/const memory.
GLIC_DI void cov2d_project(float px, float py, float pz, const float* __restrict__ v, float focal_x, float focal_y,
                           float limx_neg, float limx_pos, float limy_neg, float limy_pos, const float* c3, Cov2& o) {
    const float tx = xform_row(v, 0, px, py, pz);
    const float ty = xform_row(v, 1, px, py, pz);
    const float tz = xform_row(v, 2, px, py, pz);
    o.tz = tz;
    o.txtz = __fdiv_rn(tx, tz);
    o.tytz = __fdiv_rn(ty, tz);
    o.lx = fminf(limx_pos, fmaxf(limx_neg, o.txtz));
    o.ly = fminf(limy_pos, fmaxf(limy_neg, o.tytz));
    const float tz2 = fmul(tz, tz);
    const float J00 = __fdiv_rn(focal_x, tz);
    const float J02 = __fdiv_rn(fmul(focal_x, fmul(o.lx, -tz)), tz2);
    const float J11 = __fdiv_rn(focal_y, tz);
    const float J12 = __fdiv_rn(fmul(focal_y, fmul(o.ly, -tz)), tz2);
    o.T00 = ffma(v[2], J02, fmul(v[0], J00));
    o.T01 = ffma(v[6], J02, fmul(v[4], J00));
    o.T02 = ffma(v[10], J02, fmul(v[8], J00));
    o.T10 = ffma(v[2], J12, fmul(v[1], J11));
    o.T11 = ffma(v[6], J12, fmul(v[5], J11));
    o.T12 = ffma(v[10], J12, fmul(v[9], J11));
    const float A00 = dot3c(o.T00, c3[0], o.T01, c3[1], o.T02, c3[2]);
    const float A01 = dot3c(o.T10, c3[0], o.T11, c3[1], o.T12, c3[2]);
    const float A10 = dot3c(o.T00, c3[1], o.T01, c3[3], o.T02, c3[4]);
    const float A11 = dot3c(o.T10, c3[1], o.T11, c3[3], o.T12, c3[4]);
    const float A20 = dot3c(o.T00, c3[2], o.T01, c3[4], o.T02, c3[5]);
    const float A21 = dot3c(o.T10, c3[2], o.T11, c3[4], o.T12, c3[5]);
    o.a = fadd(dot3c(A00, o.T00, A10, o.T01, A20, o.T02), 0.3f);
    o.b = dot3c(A01, o.T00, A11, o.T01, A21, o.T02);
    o.c = fadd(dot3c(A01, o.T10, A11, o.T11, A21, o.T12), 0.3f);
}

// NDC -> pixel centre, evaluated in double like auxiliary.h:41-44.
GLIC_DI float ndc_to_pix(float v, int S) {
    return (float)(__fma_rn((double)v + 1.0, (double)S, -1.0) * 0.5);
}

struct TileRect { int x0, y0, x1, y1; };

// getRect (auxiliary.h:46-56): float arithmetic, truncation, clamp to the tile grid.
GLIC_DI TileRect tile_rect(float x, float y, int radius, int gx, int gy) {
    const float fr = (float)radius;
    TileRect r;
    r.x0 = min(gx, max(0, (int)fmul(fsub(x, fr), 0.0625f)));
    r.y0 = min(gy, max(0, (int)fmul(fsub(y, fr), 0.0625f)));
    r.x1 = min(gx, max(0, (int)fmul(fadd(fadd(fadd(x, fr), 16.0f), -1.0f), 0.0625f)));
    r.y1 = min(gy, max(0, (int)fmul(fadd(fadd(fadd(y, fr), 16.0f), -1.0f), 0.0625f)));
    return r;
}

// Exact tile test (StopThePop, forward.h:34-78): largest exponent "power" the Gaussian reaches
// inside the pixel-centre rectangle of tile (tx,ty); the tile is kept iff result <= log(o*255).
GLIC_DI float tile_max_power(float cox, float coy, float coz, float mx, float my, int tx, int ty) {
    const float tminx = (float)(tx * TILE), tminy = (float)(ty * TILE);
    const float tmaxx = (float)((tx * TILE) | (TILE - 1)), tmaxy = (float)((ty * TILE) | (TILE - 1));
    const float x_min_diff = fsub(tminx, mx);
    const float x_left = x_min_diff > 0.0f ? 1.0f : 0.0f;
    const float not_in_x = fadd(x_left, mx > tmaxx ? 1.0f : 0.0f);
    const float y_min_diff = fsub(tminy, my);
    const float y_above = y_min_diff > 0.0f ? 1.0f : 0.0f;
    const float not_in_y = fadd(y_above, my > tmaxy ? 1.0f : 0.0f);
    if (!(fadd(not_in_y, not_in_x) > 0.0f)) return 0.0f;
    const float sx = fsub(tmaxx, tminx), sy = fsub(tmaxy, tminy);
    const float px = ffma(tminx, x_left, fmul(fsub(1.0f, x_left), tmaxx));
    const float py = ffma(tminy, y_above, fmul(fsub(1.0f, y_above), tmaxy));
    const float dx = copysignf(sx, x_min_diff);
    const float dy = copysignf(sy, y_min_diff);
    const float diffx = fsub(mx, px), diffy = fsub(my, py);
    const float rcpx = __frcp_rn(fmul(cox, fmul(sx, sx)));
    const float rcpy = __frcp_rn(fmul(coz, fmul(sy, sy)));
    const float ux = fmul(ffma(fmul(coy, dx), diffy, fmul(fmul(cox, dx), diffx)), rcpx);
    const float tx_ = fmul(not_in_y, __saturatef(ux));
    const float uy = fmul(ffma(fmul(coz, dy), diffy, fmul(fmul(coy, dy), diffx)), rcpy);
    const float ty_ = fmul(not_in_x, __saturatef(uy));
    const float qx = ffma(dx, tx_, px), qy = ffma(dy, ty_, py);
    const float ex = fsub(mx, qx), ey = fsub(my, qy);
    const float h = fmul(ffma(ex, fmul(cox, ex), fmul(ey, fmul(coz, ey))), 0.5f);
    return ffma(fmul(coy, ex), ey, h);
}

// Blend exponent of one (pixel, splat) pair (forward.cu:430): d = xy - pix.
GLIC_DI float splat_power(float dx, float dy, float cx, float cy, float cz) {
    return fsub(fmul(ffma(dx, fmul(dx, cx), fmul(dy, fmul(dy, cz))), -0.5f), fmul(dy, fmul(dx, cy)));
}

}  // namespace glic
