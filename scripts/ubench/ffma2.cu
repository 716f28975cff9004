// Micro-benchmark: issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100a) - decides whether the single-sweep
// TV kernel should use packed fp32 arithmetic.  Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2 scripts/ubench/ffma2.cu && /tmp/ffma2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
        }
    } else {
        unsigned long long p[8], pa = pk(a, a), pb = pk(b, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = pk(acc[2 * i], acc[2 * i + 1]);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) upk(p[i], acc[2 * i], acc[2 * i + 1]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 0.999f, 0.001f); else k<1><<<148 * 8, 256>>>(d, iters, 0.999f, 0.001f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double fmas = 148.0 * 8 * 256 * 16.0 * iters;   // scalar fma operations
            printf("%s rep %d: %.3f ms, %.2f Tfma/s, %.1f fma/clk/SM @1.965GHz\n", mode ? "FFMA2" : "FFMA ", rep, ms, fmas / ms / 1e9, fmas / (ms * 1e-3) / 148 / 1.965e9);
        }
    }
    return 0;
}
