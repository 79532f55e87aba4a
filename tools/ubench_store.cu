// Micro-benchmark: L1 data-pipe cost of global stores by width / cache operator on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/ubench_store tools/ubench_store.cu
// Every block rewrites its own L2-resident 64 KB window, so the rate is set by the SM's store path, not HBM.
// Run plain for rates, and under `ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_lg.sum,...` for wavefronts.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float v) {
    float* base = out + (size_t)blockIdx.x * 16384;  // 64 KB per block
    for (int it = 0; it < iters; ++it) {
        const float x = v + it;
        if (MODE == 0 || MODE == 1) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {  // 16 x 1 KB rows per iteration
                float* p = base + ((it & 3) * 4096) + r * 256 + threadIdx.x;
                if (MODE == 0) asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(x) : "memory");
                else asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(x) : "memory");
            }
        } else if (MODE == 2 || MODE == 3) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float* p = base + ((it & 3) * 4096) + r * 1024 + threadIdx.x * 4;
                if (MODE == 2) asm volatile("st.global.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(x) : "memory");
                else asm volatile("st.global.cs.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(x) : "memory");
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                float* p = base + ((it & 3) * 4096) + r * 2048 + threadIdx.x * 8;
                asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "f"(x) : "memory");
            }
        }
    }
}
int main() {
    const int blocks = 148 * 8;
    float* d; cudaMalloc(&d, (size_t)blocks * 65536);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    const char* names[5] = {"st.f32      ", "st.cs.f32   ", "st.v4.f32   ", "st.cs.v4.f32", "st.v8.f32   "};
    for (int mode = 0; mode < 5; ++mode) for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        switch (mode) {
            case 0: k<0><<<blocks, 256>>>(d, iters, 1.f); break;
            case 1: k<1><<<blocks, 256>>>(d, iters, 1.f); break;
            case 2: k<2><<<blocks, 256>>>(d, iters, 1.f); break;
            case 3: k<3><<<blocks, 256>>>(d, iters, 1.f); break;
            default: k<4><<<blocks, 256>>>(d, iters, 1.f); break;
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)blocks * iters * 16384;
        if (rep) printf("%s: %.3f ms, %.2f TB/s, %.1f B/clk/SM @1.965GHz (err %d)\n", names[mode], ms, bytes / ms / 1e9,
                        bytes / ms / 1e3 / 148 / 1965e3, (int)cudaGetLastError());
    }
    return 0;
}
