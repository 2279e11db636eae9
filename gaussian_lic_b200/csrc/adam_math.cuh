// adam_math.cuh -- ONE definition of the visibility-masked Adam element update (adamUpdateCUDA, reference
// rasterizer/cuda_rasterizer/adam.cu:27-36: no bias correction, no weight decay), shared by the per-tensor kernel
// (adam_knn.cu), the packed-model kernel (model_step.cu) and the fused exchange + optimiser kernel (p2p.cu), so that the
// three are bit-identical by construction.  The operation order is pinned with round-to-nearest intrinsics (the compiler
// may neither fuse nor reassociate them): it is the order nvcc 12.9 picks for the reference's expression on sm_100a,
//     m' = fma(1-b1, g, b1*m)      v' = fma(g, (1-b2)*g, b2*v)      p' = p + (-lr*m') / (sqrt(v') + eps)
// with IEEE division and square root (the reference builds without fast-math).
#pragma once
#include <cuda_runtime.h>

namespace glic {

__device__ __forceinline__ void adam_element(float& p, float& m, float& v, float g, float lr, float b1, float b2, float eps) {
    const float mm = __fmaf_rn(__fsub_rn(1.0f, b1), g, __fmul_rn(b1, m));
    const float vv = __fmaf_rn(g, __fmul_rn(__fsub_rn(1.0f, b2), g), __fmul_rn(b2, v));
    p = __fadd_rn(p, __fdiv_rn(__fmul_rn(-lr, mm), __fadd_rn(__fsqrt_rn(vv), eps)));
    m = mm;
    v = vv;
}

}  // namespace glic
