// Micro-benchmark: write bandwidth of K1's OUTPUT PATTERN -- P channel planes, `stride` floats apart, filled
// concurrently -- as a function of how many contiguous bytes each block writes to one plane before moving on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/ubench_planes tools/ubench_planes.cu
// The 256^3 launch of K1 writes 9 planes of 64 MiB; a warp store instruction is 128 contiguous bytes of ONE plane.
#include <cstdio>
#include <cuda_runtime.h>

// Each block walks chunks of CH floats (grid-stride over chunks); inside a chunk it writes plane after plane,
// every thread VEC floats per store.  CH = 32*VEC*8 ... : contiguous bytes per plane per block visit = 4*CH.
template <int VEC>
__global__ void __launch_bounds__(256) planes_chunked(float* out, size_t plane_elems, size_t stride, int P, int CH, float v) {
    const size_t n_chunks = plane_elems / CH;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        for (int k = 0; k < P; ++k) {
            float* base = out + k * stride + c * CH;
            for (int i = threadIdx.x * VEC; i < CH; i += 256 * VEC) {
                if (VEC == 1) __stcs(base + i, v);
                else __stcs((float4*)(base + i), make_float4(v, v, v, v));
            }
        }
    }
}
// K1-like: a thread owns element i of every plane and writes the P values back to back (scalar stores).
__global__ void __launch_bounds__(256) planes_interleaved(float* out, size_t plane_elems, size_t stride, int P, float v) {
    const size_t tid = (size_t)blockIdx.x * 256 + threadIdx.x, nth = (size_t)gridDim.x * 256;
    for (size_t i = tid; i < plane_elems; i += nth)
        for (int k = 0; k < P; ++k) __stcs(out + k * stride + i, v);
}
// K1-like geometry: block = 8 rows x 64 z tile of a 256x256 plane (x = blockIdx.z), thread writes z and z+32.
__global__ void __launch_bounds__(256) planes_k1_tiles(float* out, int D, size_t stride, int P, float v) {
    const int c0 = blockIdx.x * 64 + (threadIdx.x & 31), b = blockIdx.y * 8 + (threadIdx.x >> 5);
    const size_t row = ((size_t)blockIdx.z * D + b) * D + c0;
    for (int j = 0; j < 2; ++j)
        for (int k = 0; k < P; ++k) __stcs(out + k * stride + row + 32 * j, v);
}

int main() {
    const size_t plane = (size_t)1 << 24;  // 64 MiB planes like D = 256
    float* a;
    cudaMalloc(&a, (plane + 4096) * 9 * 4 + (1 << 20));
    cudaMemset(a, 0, (plane + 4096) * 9 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](const char* name, double bytes, auto launch) {
        float best = 1e9f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r && ms < best) best = ms;
        }
        printf("%-58s %8.3f ms  %7.1f GB/s  (err %d)\n", name, best, bytes / best / 1e6, (int)cudaGetLastError());
    };
    char nm[96];
    const int g = 148 * 8;
    for (int P : {1, 2, 3, 5, 9}) {
        snprintf(nm, 96, "interleaved scalar, P=%d, stride 2^24", P);
        time(nm, (double)P * plane * 4, [&] { planes_interleaved<<<g, 256>>>(a, plane, plane, P, 1.f); });
    }
    for (size_t pad : {(size_t)0, (size_t)32, (size_t)1024, (size_t)4096}) {
        snprintf(nm, 96, "interleaved scalar, P=9, stride 2^24+%zu", pad);
        time(nm, 9.0 * plane * 4, [&] { planes_interleaved<<<g, 256>>>(a, plane, plane + pad, 9, 1.f); });
    }
    for (int CH : {256, 512, 2048, 8192, 32768, 131072}) {
        snprintf(nm, 96, "chunked scalar st, P=9, %d B contiguous per plane visit", CH * 4);
        time(nm, 9.0 * plane * 4, [&] { planes_chunked<1><<<g, 256>>>(a, plane, plane, 9, CH, 1.f); });
    }
    for (int CH : {1024, 2048, 8192, 32768, 131072}) {
        snprintf(nm, 96, "chunked v4 st, P=9, %d B contiguous per plane visit", CH * 4);
        time(nm, 9.0 * plane * 4, [&] { planes_chunked<4><<<g, 256>>>(a, plane, plane, 9, CH, 1.f); });
    }
    for (int gg : {148 * 4, 148 * 16, 148 * 32}) {
        snprintf(nm, 96, "chunked v4 st, P=9, 8192 B, grid %d", gg);
        time(nm, 9.0 * plane * 4, [&] { planes_chunked<4><<<gg, 256>>>(a, plane, plane, 9, 2048, 1.f); });
    }
    time("K1 tiles (8 rows x 64 z), P=9", 9.0 * plane * 4, [&] { planes_k1_tiles<<<dim3(4, 32, 256), 256>>>(a, 256, plane, 9, 1.f); });
    time("K1 tiles (8 rows x 64 z), P=1", 1.0 * plane * 4, [&] { planes_k1_tiles<<<dim3(4, 32, 256), 256>>>(a, 256, plane, 1, 1.f); });
    return 0;
}
