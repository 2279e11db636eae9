// Minimal stand-in for the parts of GLM (OpenGL Mathematics) that the reference
// rasterizer uses.  GLM is an UN-VENDORED third-party dependency of the reference
// (#include <glm/glm.hpp> at src/rasterizer/cuda_rasterizer/forward.h:26,
// backward.h:26, adam.h:8, rasterizer_impl.cu:30) and is not installed in this
// image.  This header exists ONLY so that the reference's own .cu files can be
// compiled unmodified into oracle/_ref/ (test/bench infrastructure).  It is not
// part of the product and nothing under gaussian_lic_b200/ includes it.
//
// Semantics follow GLM's published conventions: column-major matrices,
// M[c][r] indexing, mat3(9 scalars) fills column by column, mat3(s) = s*I,
// mat*mat accumulates left to right (a[0]*b[j].x + a[1]*b[j].y + a[2]*b[j].z),
// dot(a,b) = a.x*b.x + a.y*b.y + a.z*b.z.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define GLMS_FN __host__ __device__ inline
#else
#define GLMS_FN inline
#endif

namespace glm {

struct vec2 {
    float x, y;
    vec2() = default;
    template <class A, class B> GLMS_FN vec2(A a, B b) : x(float(a)), y(float(b)) {}
};

struct vec3 {
    float x, y, z;
    vec3() = default;
    template <class A, class B, class C> GLMS_FN vec3(A a, B b, C c) : x(float(a)), y(float(b)), z(float(c)) {}
    GLMS_FN float& operator[](int i) { return (&x)[i]; }
    GLMS_FN const float& operator[](int i) const { return (&x)[i]; }
    GLMS_FN vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    GLMS_FN vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
    GLMS_FN vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4 {
    float x, y, z, w;
    vec4() = default;
    template <class A, class B, class C, class D> GLMS_FN vec4(A a, B b, C c, D d) : x(float(a)), y(float(b)), z(float(c)), w(float(d)) {}
};

GLMS_FN vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLMS_FN vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLMS_FN vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLMS_FN vec3 operator*(float s, const vec3& v) { return vec3(s * v.x, s * v.y, s * v.z); }
GLMS_FN vec3 operator*(const vec3& v, float s) { return vec3(v.x * s, v.y * s, v.z * s); }
GLMS_FN vec3 operator/(const vec3& v, float s) { return vec3(v.x / s, v.y / s, v.z / s); }

GLMS_FN float dot(const vec3& a, const vec3& b) {
    vec3 t = a * b;
    return t.x + t.y + t.z;
}
GLMS_FN float length(const vec3& v) { return sqrtf(dot(v, v)); }
GLMS_FN vec3 max(const vec3& v, float s) { return vec3(fmaxf(v.x, s), fmaxf(v.y, s), fmaxf(v.z, s)); }

struct mat3 {
    vec3 c[3];
    mat3() = default;
    GLMS_FN explicit mat3(float s) {
        c[0] = vec3(s, 0.f, 0.f);
        c[1] = vec3(0.f, s, 0.f);
        c[2] = vec3(0.f, 0.f, s);
    }
    template <class A0, class A1, class A2, class B0, class B1, class B2, class C0, class C1, class C2>
    GLMS_FN mat3(A0 a0, A1 a1, A2 a2, B0 b0, B1 b1, B2 b2, C0 c0, C1 c1, C2 c2) {
        c[0] = vec3(a0, a1, a2);
        c[1] = vec3(b0, b1, b2);
        c[2] = vec3(c0, c1, c2);
    }
    GLMS_FN vec3& operator[](int i) { return c[i]; }
    GLMS_FN const vec3& operator[](int i) const { return c[i]; }
};

GLMS_FN mat3 operator*(const mat3& a, const mat3& b) {
    mat3 r;
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i)
            r[j][i] = a[0][i] * b[j][0] + a[1][i] * b[j][1] + a[2][i] * b[j][2];
    return r;
}
GLMS_FN mat3 operator*(float s, const mat3& m) {
    mat3 r;
    r[0] = m[0] * s;
    r[1] = m[1] * s;
    r[2] = m[2] * s;
    return r;
}
GLMS_FN mat3 operator*(const mat3& m, float s) { return s * m; }
GLMS_FN mat3 transpose(const mat3& m) {
    return mat3(m[0][0], m[1][0], m[2][0],
                m[0][1], m[1][1], m[2][1],
                m[0][2], m[1][2], m[2][2]);
}

}  // namespace glm
