// microbench_f32x2.cu -- does sm_100a's packed fp32 (fma.rn.f32x2 / mul.f32x2 / add.f32x2, SASS FFMA2/FMUL2/FADD2)
// buy issue slots, FMA-pipe throughput, or both?  Decides how the render kernels are packed (profiles/r02_f32x2.txt).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gaussian_lic_b200/_build/mb_f32x2 scripts/microbench_f32x2.cu
// Prints, per variant and warps/SM: FMA lanes retired per clock per SM (128 = the plain-FFMA pipe limit).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;      // independent dependency chains per thread

__global__ void k_scalar(float* out, float a, float b) {
    float x[2 * CHAINS];
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 2 * CHAINS; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_packed(float* out, float a, float b) {
    unsigned long long x[CHAINS], aa, bb;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        float lo = threadIdx.x * 1e-3f + i, hi = lo + 0.5f;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(lo), "f"(hi));
    }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { float lo, hi; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed FMAs interleaved with ALU-pipe work (the render loop's mix): does packing free issue slots for the other pipe?
__global__ void k_mix_scalar(float* out, float a, float b) {
    float x[2 * CHAINS]; unsigned y[CHAINS];
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; ++i) x[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) y[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[2 * i]) : "f"(a), "f"(b));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[2 * i + 1]) : "f"(a), "f"(b));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(it), "r"(i));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(i), "r"(it));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += (float)y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mix_packed(float* out, float a, float b) {
    unsigned long long x[CHAINS], aa, bb; unsigned y[CHAINS];
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        float lo = threadIdx.x * 1e-3f + i, hi = lo + 0.5f;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(lo), "f"(hi));
        y[i] = threadIdx.x + i;
    }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(it), "r"(i));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(i), "r"(it));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { float lo, hi; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi + (float)y[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
void run(const char* name, K kern, int fma_lanes_per_thread_iter, int other_per_thread_iter) {
    int sms = 148, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 1024 * 2);
    for (int warps : {4, 8, 16, 32}) {
        const int threads = warps * 32 > 1024 ? 1024 : warps * 32;
        const int blocks = sms * (warps * 32 / threads);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        kern<<<blocks, threads>>>(out, 1.0001f, 1e-7f);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int r = 0; r < 5; ++r) kern<<<blocks, threads>>>(out, 1.0001f, 1e-7f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= 5;
        const double clocks = ms * 1e-3 * khz * 1e3;                 // at the nominal max clock
        const double fma = (double)warps * 32 * ITERS * fma_lanes_per_thread_iter / clocks;   // per SM
        const double oth = (double)warps * 32 * ITERS * other_per_thread_iter / clocks;
        printf("%-12s warps/SM %2d: %.3f ms  FMA lanes/clk/SM %.1f  other lanes/clk/SM %.1f\n", name, warps, ms, fma, oth);
    }
    cudaFree(out);
}

int main() {
    run("scalar", k_scalar, 2 * CHAINS, 0);
    run("packed", k_packed, 2 * CHAINS, 0);
    run("mix_scalar", k_mix_scalar, 2 * CHAINS, 2 * CHAINS);
    run("mix_packed", k_mix_packed, 2 * CHAINS, 2 * CHAINS);
    return 0;
}
