// Micro-benchmark: scalar FFMA vs packed FFMA2 issue throughput on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench_ffma2.cu && /tmp/ubench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    float x[8]; u64 y[8];
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; y[i] = pk(x[i], x[i] + 1); }
    const u64 a2 = pk(a, a), b2 = pk(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) x[i] = fma1(x[i], a, b);
            else y[i] = fma2(y[i], a2, b2);
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += (MODE == 0) ? x[i] : __uint_as_float((unsigned)y[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 4, 512>>>(d, iters, 1.0001f, 0.5f); else k<1><<<148 * 4, 512>>>(d, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double inst = 148.0 * 4 * 512 / 32 * iters * 8;  // warp instructions
        printf("%s: %.3f ms, %.1f G warp-inst/s, %.2f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", ms, inst / ms / 1e6,
               inst * 32 * (mode ? 4 : 2) / ms / 1e9);
    }
    return 0;
}
