// Device self-tests of the exact-division shortcuts K1 relies on (called from tests/ only).
#include "common.cuh"
#include "f32x2.cuh"

namespace {

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// s / n for every fp32 bit pattern s and n = 1..max_n, Markstein form vs div.rn.f32.
__global__ void selftest_div_count_kernel(int max_n, unsigned long long* mismatches) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ULL << 32); i += stride) {
        const float s = __uint_as_float((unsigned)i);
        if (!(fabsf(s) <= 3.0e38f) || (s != 0.f && fabsf(s) < 1e-30f)) continue;  // NaN/Inf/near-denormal: not produced
        for (int n = 1; n <= max_n; ++n) {
            const float fn = (float)n, r = __frcp_rn(fn);
            const f32x2 q0 = mul2(bc(s), bc(r));
            const f32x2 rem = fma2(bc(-fn), q0, bc(s));
            const float q = lo(fma2(rem, bc(r), q0));
            bad += !(q == __fdiv_rn(s, fn));  // value compare: the only bit difference allowed is the sign of zero
        }
    }
    atomicAdd(mismatches, bad);
}

// a / b via the shared-reciprocal fast path vs div.rn.f32, b in [1e-8, 2^100), a random finite.
__global__ void selftest_div_pair_kernel(unsigned long long n_cases, unsigned seed, unsigned long long* mismatches) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_cases; i += stride) {
        const unsigned h0 = hash32((unsigned)i ^ seed), h1 = hash32(h0 + 0x9e3779b9U), h2 = hash32(h1 ^ (unsigned)(i >> 32));
        // denominator: exponent in [-27, 99], any mantissa, positive, >= 1e-8
        float b = __uint_as_float(((100u + (h0 % 127u)) << 23) | (h1 & 0x7fffffu));
        b = fmaxf(b, 1e-8f);
        // numerators: any finite value (sign, exponent 1..254)
        const float a0 = __uint_as_float((h1 & 0x80000000u) | ((1u + (h2 % 254u)) << 23) | (h2 >> 9));
        const float a1 = __uint_as_float((h2 & 0x80000000u) | ((1u + (h0 % 254u)) << 23) | (h0 >> 9));
        const Recip2 rc = recip2(pk(b, b));
        const f32x2 q = div2(pk(a0, a1), rc);
        const float e0 = __fdiv_rn(a0, b), e1 = __fdiv_rn(a1, b);
        // contract (f32x2.cuh): exact whenever the true quotient is a normal number >= 2^-76 in magnitude
        if (fabsf(e0) >= 1.4e-23f && fabsf(e0) <= 3.0e38f) bad += __float_as_uint(lo(q)) != __float_as_uint(e0);
        if (fabsf(e1) >= 1.4e-23f && fabsf(e1) <= 3.0e38f) bad += __float_as_uint(hi(q)) != __float_as_uint(e1);
    }
    atomicAdd(mismatches, bad);
}

}  // namespace

extern "C" int gens_selftest_division(int max_n, unsigned long long n_pair_cases, unsigned long long* d_mismatches2,
                                      void* stream) {
    GENS_CHECK_ARG(d_mismatches2 && max_n > 0 && max_n <= 64);
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(d_mismatches2, 0, 2 * sizeof(unsigned long long), st);
    selftest_div_count_kernel<<<148 * 8, 256, 0, st>>>(max_n, d_mismatches2);
    selftest_div_pair_kernel<<<148 * 8, 256, 0, st>>>(n_pair_cases, 0x1234567u, d_mismatches2 + 1);
    return gens_launch_status();
}
