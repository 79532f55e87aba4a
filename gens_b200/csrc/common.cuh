// Shared helpers for the gens_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gens_b200.h"

#define GENS_MAX_VIEWS 16

#define GENS_CHECK_ARG(cond) \
    do {                     \
        if (!(cond)) return GENS_E_BADARG; \
    } while (0)

static inline int gens_launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

// Row of a 4x4 matrix times (x0,x1,x2,x3), k-ascending fma chain.  This is the order the
// reference's matmul uses for K=4 (MKL on CPU, verified bit-exact; see oracle/gens_oracle.c)
// and it is spelled with _rn intrinsics so that -fmad cannot re-associate it.
__device__ __forceinline__ float row_dot4(const float* a, float x0, float x1, float x2, float x3) {
    float t = __fmul_rn(a[0], x0);
    t = __fmaf_rn(a[1], x1, t);
    t = __fmaf_rn(a[2], x2, t);
    t = __fmaf_rn(a[3], x3, t);
    return t;
}

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

static inline __host__ __device__ int ceil_div_i(long long a, long long b) { return (int)((a + b - 1) / b); }
