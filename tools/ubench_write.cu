// Micro-benchmark: what a WRITE-ONLY stream to HBM reaches on the B200 (the ceiling of K1, which writes
// 604 MB and reads 15 MB per 256^3 launch), next to a copy (the read+write figure MEASURED_PEAKS.json quotes).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/ubench_write tools/ubench_write.cu
// Buffers are 1 GiB (8x the L2), each kernel runs 5 times, the best time is printed.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>  // 0 st.f32, 1 st.cs.f32, 2 st.v4, 3 st.cs.v4, 4 planes like K1 (9 x 128 B rows per warp)
__global__ void __launch_bounds__(256) fill(float* out, size_t n, float v) {
    const size_t tid = (size_t)blockIdx.x * 256 + threadIdx.x, nth = (size_t)gridDim.x * 256;
    if (MODE == 0 || MODE == 1) {
        for (size_t i = tid; i < n; i += nth) {
            if (MODE == 0) out[i] = v; else __stcs(out + i, v);
        }
    } else if (MODE == 2 || MODE == 3) {
        float4* o = (float4*)out;
        for (size_t i = tid; i < n / 4; i += nth) {
            if (MODE == 2) o[i] = make_float4(v, v, v, v); else __stcs(o + i, make_float4(v, v, v, v));
        }
    } else {
        const size_t plane = n / 9;  // 9 channel planes, a thread writes the same offset of each
        for (size_t i = tid; i < plane; i += nth)
#pragma unroll
            for (int k = 0; k < 9; ++k) __stcs(out + k * plane + i, v);
    }
}
__global__ void __launch_bounds__(256) copy4(const float4* __restrict__ in, float4* __restrict__ out, size_t n4) {
    const size_t tid = (size_t)blockIdx.x * 256 + threadIdx.x, nth = (size_t)gridDim.x * 256;
    for (size_t i = tid; i < n4; i += nth) __stcs(out + i, __ldcs(in + i));
}
__global__ void __launch_bounds__(256) read4(const float4* __restrict__ in, float* sink, size_t n4) {
    const size_t tid = (size_t)blockIdx.x * 256 + threadIdx.x, nth = (size_t)gridDim.x * 256;
    float acc = 0.f;
    for (size_t i = tid; i < n4; i += nth) { const float4 v = __ldcs(in + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 12345.678f) *sink = acc;
}
// non-persistent variant: one thread = one float4, grid covers the buffer
__global__ void __launch_bounds__(256) fill_flat(float4* out, size_t n4, float v) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n4) __stcs(out + i, make_float4(v, v, v, v));
}

int main() {
    const size_t n = (size_t)1 << 28;  // floats = 1 GiB
    float *a, *b;
    cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4);
    cudaMemset(a, 0, n * 4); cudaMemset(b, 0, n * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](const char* name, double bytes, auto launch) {
        float best = 1e9f;
        for (int r = 0; r < 6; ++r) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r && ms < best) best = ms;
        }
        printf("%-34s %8.3f ms  %7.1f GB/s  (err %d)\n", name, best, bytes / best / 1e6, (int)cudaGetLastError());
    };
    const double B = (double)n * 4;
    for (int bps : {4, 8, 16}) {
        const int g = 148 * bps;
        char nm[64];
        snprintf(nm, 64, "write st.f32 grid=148x%d", bps);      time(nm, B, [&] { fill<0><<<g, 256>>>(a, n, 1.f); });
        snprintf(nm, 64, "write st.cs.f32 grid=148x%d", bps);   time(nm, B, [&] { fill<1><<<g, 256>>>(a, n, 1.f); });
        snprintf(nm, 64, "write st.v4 grid=148x%d", bps);       time(nm, B, [&] { fill<2><<<g, 256>>>(a, n, 1.f); });
        snprintf(nm, 64, "write st.cs.v4 grid=148x%d", bps);    time(nm, B, [&] { fill<3><<<g, 256>>>(a, n, 1.f); });
        snprintf(nm, 64, "write 9 planes st.cs grid=148x%d", bps); time(nm, B / 9 * 9, [&] { fill<4><<<g, 256>>>(a, n / 9 * 9, 1.f); });
        snprintf(nm, 64, "copy ld.cs->st.cs v4 grid=148x%d", bps); time(nm, 2 * B, [&] { copy4<<<g, 256>>>((float4*)a, (float4*)b, n / 4); });
        snprintf(nm, 64, "read ld.cs v4 grid=148x%d", bps);     time(nm, B, [&] { read4<<<g, 256>>>((float4*)a, b, n / 4); });
    }
    time("write st.cs.v4 flat grid", B, [&] { fill_flat<<<(unsigned)(n / 4 / 256), 256>>>((float4*)a, n / 4, 1.f); });
    time("cudaMemsetAsync", B, [&] { cudaMemsetAsync(a, 0, n * 4); });
    time("cudaMemcpyAsync D2D (r+w bytes)", 2 * B, [&] { cudaMemcpyAsync(b, a, n * 4, cudaMemcpyDeviceToDevice); });
    // 604 MB like one K1 launch, after an L2 flush by the copy above
    time("write st.cs.v4 604 MB flat", 604e6, [&] { fill_flat<<<(unsigned)(151000000 / 4 / 256), 256>>>((float4*)a, 151000000 / 4, 1.f); });
    return 0;
}
