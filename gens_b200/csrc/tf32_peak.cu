// Measurement probe: dense TF32 throughput of the 5th-generation tensor cores with resident operands.
//
// MEASURED_PEAKS.json holds a bf16 cuBLAS figure only; the SDF MLP kernels (sdf_mlp_tc.cu) issue
// tcgen05.mma.kind::tf32, so their tensor-pipe fraction needs a TF32 denominator measured on the same part.
// One CTA per SM; a single elected thread issues `iters` 128x256x8 kind::tf32 MMAs from two shared-memory
// operand tiles into two alternating TMEM accumulators (all 512 columns), with nothing else in flight: no loads,
// no epilogue.  2 * 128 * 256 * 8 flop per instruction; bench.py divides by the CUDA-event time.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t kmajor_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3ffff) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__global__ void __launch_bounds__(128, 1) tf32_peak_kernel(int iters, double* __restrict__ out) {
    // A: 128 rows x 8 tf32 (two 16-byte K chunks), B: 256 rows x 8 tf32; core matrices of 8 rows x 16 bytes
    __shared__ __align__(128) float s_a[128 * 8];
    __shared__ __align__(128) float s_b[256 * 8];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_tmem;
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) s_a[i] = 1.0f + 0.001f * (float)(i % 97);
    for (int i = threadIdx.x; i < 256 * 8; i += blockDim.x) s_b[i] = 0.5f - 0.002f * (float)(i % 89);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t bar = s_u32(&s_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&s_tmem)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&s_tmem);

    if (threadIdx.x == 0) {
        const uint64_t a_desc = kmajor_desc(s_u32(s_a), 16 * 128, 128);
        const uint64_t b_desc = kmajor_desc(s_u32(s_b), 32 * 128, 128);
        // D fp32, A/B tf32, both K-major, N = 256, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem + ((i & 1) ? 256u : 0u);
            const uint32_t acc = i >= 2 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        uint32_t done;
        do {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(0u)
                : "memory");
        } while (!done);
        // keep the accumulators observable: one element of each into the output
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) {
        uint32_t v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)__uint_as_float(v);
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)gridDim.x;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

}  // namespace

// out[0] = number of CTAs launched (one per SM), out[1] = one accumulator element (keeps the MMAs alive).
extern "C" int gens_tf32_mma_peak(int iters, double* out, void* stream) {
    GENS_CHECK_ARG(out && iters > 0);
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return gens_launch_status();
    tf32_peak_kernel<<<sms, 128, 0, (cudaStream_t)stream>>>(iters, out);
    return gens_launch_status();
}
