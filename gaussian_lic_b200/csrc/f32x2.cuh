// f32x2.cuh -- sm_100 packed fp32 arithmetic (PTX fma/mul/add .f32x2 -> SASS FFMA2 / FMUL2 / FADD2).
//
// One instruction carries two independent IEEE round-to-nearest fp32 operations (no flush-to-zero), so each half is
// bit-identical to the scalar __fmaf_rn / __fmul_rn / __fadd_rn it replaces; operands may be a register pair or a
// broadcast scalar (ptxas folds the {s, s} pack into the instruction's .F32 operand form).  Measured on B200
// (profiles/r02_f32x2_microbench.txt): the FMA pipe retires the same 128 lanes/clk/SM either way -- what packing buys is
// ISSUE slots (one slot per two FMAs), which is exactly the limiter of the render and SSIM kernels (ncu: 75-93 %
// issue-active).  The kernels pack the two pixels a lane owns.
#pragma once
#include <cuda_runtime.h>

namespace glic {

struct f2 { unsigned long long v; };

__device__ __forceinline__ f2 f2_pack(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f2 f2_bcast(float s) { return f2_pack(s, s); }
__device__ __forceinline__ void f2_unpack(f2 a, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
}
__device__ __forceinline__ float f2_lo(f2 a) { float l, h; f2_unpack(a, l, h); return l; }
__device__ __forceinline__ float f2_hi(f2 a) { float l, h; f2_unpack(a, l, h); return h; }

__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
// a - b as fma(b, -1, a): one rounding of the exact difference, i.e. __fsub_rn(a, b) per half
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b) { return f2_fma(b, f2_bcast(-1.0f), a); }
__device__ __forceinline__ f2 f2_mul(f2 a, float s) { return f2_mul(a, f2_bcast(s)); }
__device__ __forceinline__ f2 f2_fma(f2 a, float s, f2 c) { return f2_fma(a, f2_bcast(s), c); }
__device__ __forceinline__ f2 f2_fma(float s, float t, f2 c) { return f2_fma(f2_bcast(s), f2_bcast(t), c); }

}  // namespace glic
