// K4: the SDF MLP on tcgen05 tensor cores -- persistent kernels, error-compensated 3xTF32.
//
// Three kernels replace, for inference (torch.no_grad), the per-layer cuBLAS SGEMM + glue-kernel chains of
// gens_b200/sdf_analytic.py whose (n,128) activations travel through HBM between every two launches
// (reference models/modules/sdf_network.py:98-153: forward, .sdf and .gradient):
//
//   sdf_mlp_fwd_kernel<false>   value pass: encodings -> SDF, 128 points per tile
//   sdf_mlp_fwd_kernel<true>    value + directional-derivative pass (rows 2i / 2i+1 of a tile = primal /
//                               tangent of point i, 64 points per tile); also stores sp'(a) and sp''(a) da
//   sdf_mlp_rev_kernel          reverse sweep over the same 64-point tiles: cotangents of the encodings
//                               (and their tangents) for grad sdf and the second-order "smooth" term
//
// Common design:
//   * the running activations (A operand of the next GEMM) live in TENSOR MEMORY (lane = row, column =
//     channel), split into a TF32-exact high part and the fp32 remainder;
//   * the encodings every forward layer re-reads (position encoding P, 27 ch; volume-feature encoding F,
//     100 ch) sit in shared memory in the canonical K-major no-swizzle UMMA layout ([K/4][128 rows][16 B]),
//     also split hi/lo;
//   * weights stream through a ring of 16 KB shared-memory slots with cp.async.bulk (one k-step = N x 16
//     weights, hi and lo, pre-packed on the host in exactly the slot image, gens_b200/mlp_tc.py);
//   * per 8 channels three tcgen05.mma.kind::tf32 accumulate  Ah.Bh + Al.Bh + Ah.Bl  into fp32 accumulators
//     in TMEM: what is dropped (Al.Bl, the truncation of the low parts) is below 2^-21 relative, i.e. the
//     result matches an fp32 SGEMM to ~1e-5 after seven layers -- plain TF32 (2^-11) would not survive
//     softplus(beta = 100) at the stated 1e-4 tolerance;
//   * 16 epilogue warps pull the accumulators with tcgen05.ld, apply bias / softplus (and its derivatives),
//     split and write the next A operand back with tcgen05.st; in the forward kernels the accumulator is
//     double-buffered so the F/P k-steps of layer l+1 (independent of h_l) overlap layer l's epilogue.
//
// TMEM columns: [0,128) A hi | [128,256) A lo | [256,384) accumulator 0 | [384,512) accumulator 1.
// Warp roles: 0-15 epilogue / input staging (thread t <-> row t & 127, column quarter t >> 7),
//             16 weight producer (+ TMEM allocation), 17 MMA issuer (whole warp walks, one elected lane issues).
#include "common.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kEpiThreads = 512;
constexpr int kProducerWarp = 16;
constexpr int kMmaWarp = 17;
constexpr int kThreads = 576;
constexpr int kWSlotBytes = 16384;        // one k-step = 16 input channels: N x 16 weights, hi and lo
constexpr int kWStages = 4;
constexpr int kFChunks = 28;              // feature encoding: K = 112 (100 + 12 zero columns)
constexpr int kPChunks = 8;               // position encoding: K = 32 (27 + 5 zero columns)
constexpr int kChunkBytes = kTileM * 16;  // one 16-byte K chunk of all 128 rows
constexpr int kMaxLayers = 8;
constexpr int kMaxKSteps = 128;
constexpr int kNF = 100, kNP = 27;        // real widths of the encodings

// shared-memory map (bytes)
constexpr int kOffFhi = 0;
constexpr int kOffFlo = kOffFhi + kFChunks * kChunkBytes;
constexpr int kOffPhi = kOffFlo + kFChunks * kChunkBytes;
constexpr int kOffPlo = kOffPhi + kPChunks * kChunkBytes;
constexpr int kOffW = kOffPlo + kPChunks * kChunkBytes;
constexpr int kOffBias = kOffW + kWStages * kWSlotBytes;
constexpr int kOffSteps = kOffBias + kMaxLayers * 128 * 4;  // per k-step issue records (uint4)
constexpr int kOffBar = kOffSteps + kMaxKSteps * 16;
constexpr int kNumBars = 2 * kWStages + 6;  // full[], empty[], acc[2], a_ready, in_ready, tape[2]
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
// the reverse kernel keeps no encodings in shared memory: the ring starts at 0
constexpr int kRevOffW = 0;
// ... and stages the s1 / t2 tapes there instead: two buffers x two tapes x 32 KB (64 points x 128 channels)
constexpr int kRevOffStage = kWStages * kWSlotBytes;
// ... and, behind them, the skip layer's position cotangents of the tile in flight ([128 rows][27] floats)
constexpr int kRevOffSkip = kRevOffStage + 4 * 32 * 64 * 16;
static_assert(kRevOffSkip + kTileM * kNP * 4 <= kOffBias, "reverse-kernel staging overlaps the bias block");
static_assert(kNF % 4 == 0, "feature-encoding rows are read / written with 16-byte accesses");

constexpr uint32_t kColAhi = 0, kColAlo = 128, kColAcc0 = 256, kColAcc1 = 384;

// The tape between the JVP forward and the reverse kernel: per 64-point tile and hidden layer one contiguous 64 KB
// block [which: sp'(a) | sp''(a) da][chunk of 4 channels: 32][slot: 64] float4, point pt of the tile in slot
// (pt + chunk) & 63.  That is exactly the shared-memory image the reverse kernel's epilogue reads without bank conflicts
// (16 consecutive points of one chunk are 256 contiguous bytes), so the reverse kernel fetches a layer's tapes with ONE
// cp.async.bulk; the forward's stores are coalesced the same way.
constexpr uint32_t kTapeBytes = 32u * 64u * 16u;  // one tape of one layer of one tile
__device__ __forceinline__ uint32_t tape_slot(int pt, int chunk) { return (uint32_t)chunk * 1024u + (uint32_t)((pt + chunk) & 63) * 16u; }

struct KStep {
    uint32_t w_off;    // byte offset of this k-step's [hi | lo] weight block in the stream (blocks are contiguous)
    uint32_t w_bytes;  // N * 128
    uint32_t a;        // bits 0-7: A source (0 = F smem, 1 = P smem, 2 = TMEM); bits 8-15: k-step inside it (K / 16);
                       // bits 16-27: accumulator column in TMEM
    uint32_t flags;    // bit 0 overwrite the accumulator, bit 1 commit after this k-step (bit 4: to barrier acc1),
                       // bit 2 wait for the epilogue's A operand first; bits 16-24 N
};

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred;
}

// D[tmem] (+)= A[smem desc] . B[smem desc]^T, M = 128, kind::tf32
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// same with A from tensor memory
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}

// K-major, no swizzle: element (row r, 16-byte chunk c) at start + (r % 8) * 16 + (r / 8) * SBO + c * LBO
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3ffff) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base offset 0, layout type 0 (no swizzle)
}
// instruction descriptor: D fp32, A/B tf32, both K-major, M = 128
__device__ __forceinline__ uint32_t instr_desc(uint32_t n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return v;
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// TF32-exact high part (round to nearest on the 13 dropped bits) and the exact fp32 remainder
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    const uint32_t h = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    hi = h;
    lo = __float_as_uint(x - __uint_as_float(h));
}


#define TC_LIST16(v) v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// log1p(x) on [0,1] through the special-function unit: lg2.approx has an absolute error of ~2^-22 on [1,2], i.e.
// 1.7e-7 after the ln 2 factor -- the same as the degree-8 polynomial it replaces, at one MUFU + two FP32
// instructions instead of a chain of nine dependent FMAs (the epilogue warps are issue/latency bound).
__device__ __forceinline__ float log1p_unit(float x) {
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + x));
    return l * 0.6931471805599453f;
}
__device__ __forceinline__ float exp_neg_abs(float t) {  // exp(-|t|)
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-fabsf(t) * 1.4426950408889634f));
    return e;
}
// torch.nn.Softplus(beta = 100, threshold = 20) (reference sdf_network.py:95): a if 100 a > 20 else
// log1p(exp(100 a)) / 100, in the overflow-free form max(t,0) + log1p(exp(-|t|)).  Absolute error ~2e-9.
__device__ __forceinline__ float softplus100(float a) {
    const float t = a * 100.0f;
    const float sp = (fmaxf(t, 0.0f) + log1p_unit(exp_neg_abs(t))) * 0.01f;
    return t > 20.0f ? a : sp;
}
// sp'(a) = sigmoid(100 a) and sp''(a) = 100 sp' (1 - sp'), torch's linear region (100 a > 20) has (1, 0)
__device__ __forceinline__ void softplus100_d12(float a, float& d1, float& d2) {
    const float t = a * 100.0f;
    const float e = exp_neg_abs(t);
    const float r = __frcp_rn(1.0f + e);
    const float er = e * r;
    d1 = t > 0.0f ? r : er;
    d2 = 100.0f * er * r;
    if (t > 20.0f) {
        d1 = 1.0f;
        d2 = 0.0f;
    }
}

// softplus and both derivatives from one exponential: exactly softplus100() and softplus100_d12()
__device__ __forceinline__ void softplus100_all(float a, float& sp, float& d1, float& d2) {
    const float t = a * 100.0f;
    const float e = exp_neg_abs(t);
    const float r = __frcp_rn(1.0f + e);
    const float er = e * r;
    sp = (fmaxf(t, 0.0f) + log1p_unit(e)) * 0.01f;
    d1 = t > 0.0f ? r : er;
    d2 = 100.0f * er * r;
    if (t > 20.0f) {
        sp = a;
        d1 = 1.0f;
        d2 = 0.0f;
    }
}

struct Ctx {
    uint8_t* smem;
    uint32_t s_base, bar0, tmem;
    __device__ __forceinline__ uint32_t bar_full(int s) const { return bar0 + 8u * s; }
    __device__ __forceinline__ uint32_t bar_empty(int s) const { return bar0 + 8u * (kWStages + s); }
    __device__ __forceinline__ uint32_t bar_acc(int i) const { return bar0 + 8u * (2 * kWStages + i); }
    __device__ __forceinline__ uint32_t bar_a() const { return bar0 + 8u * (2 * kWStages + 2); }
    __device__ __forceinline__ uint32_t bar_in() const { return bar0 + 8u * (2 * kWStages + 3); }
    __device__ __forceinline__ uint32_t bar_tape(int i) const { return bar0 + 8u * (2 * kWStages + 4 + i); }
};

// One-time setup shared by the kernels: bias / constant rows and the k-step issue records to shared memory,
// barriers, TMEM allocation.  Issue record: x = A hi (descriptor low word, or TMEM column), y = A lo,
// z = flags (bit 3 added: A in TMEM) | N << 16, w = bytes | accumulator column << 16.
__device__ __forceinline__ Ctx setup(uint8_t* smem, const KStep* __restrict__ ksteps, int n_ksteps,
                                     const float* __restrict__ bias, int n_bias_rows) {
    Ctx c;
    c.smem = smem;
    c.s_base = smem_u32(smem);
    c.bar0 = c.s_base + kOffBar;
    float* s_bias = reinterpret_cast<float*>(smem + kOffBias);
    for (int i = threadIdx.x; i < n_bias_rows * 128; i += kThreads) s_bias[i] = bias[i];
    uint4* s_steps = reinterpret_cast<uint4*>(smem + kOffSteps);
    for (int i = threadIdx.x; i < n_ksteps; i += kThreads) {
        const KStep st = ksteps[i];
        const uint32_t kind = st.a & 0xff, kidx = (st.a >> 8) & 0xff, acc_col = (st.a >> 16) & 0xfff;
        uint4 r;
        if (kind == 2) {
            r.x = kColAhi + kidx * 16;
            r.y = kColAlo + kidx * 16;
        } else {
            const uint32_t off_hi = kind == 0 ? kOffFhi : kOffPhi, off_lo = kind == 0 ? kOffFlo : kOffPlo;
            r.x = (uint32_t)smem_desc(c.s_base + off_hi + kidx * 4 * kChunkBytes, kChunkBytes, 128);
            r.y = (uint32_t)smem_desc(c.s_base + off_lo + kidx * 4 * kChunkBytes, kChunkBytes, 128);
        }
        r.z = st.flags | (kind == 2 ? 8u : 0u);
        r.w = st.w_bytes | (acc_col << 16);
        s_steps[i] = r;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kWStages; ++s) {
            mbar_init(c.bar_full(s), 1);
            mbar_init(c.bar_empty(s), 1);
        }
        mbar_init(c.bar_acc(0), 1);
        mbar_init(c.bar_acc(1), 1);
        mbar_init(c.bar_a(), kEpiThreads);
        mbar_init(c.bar_in(), kEpiThreads);
        mbar_init(c.bar_tape(0), 1);
        mbar_init(c.bar_tape(1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if ((threadIdx.x >> 5) == kProducerWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(c.s_base + kOffTmemPtr),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    c.tmem = *reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr);
    return c;
}

__device__ __forceinline__ void teardown(const Ctx& c) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == kProducerWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "r"(512u) : "memory");
    }
}

// ===== weight producer (one lane): the same k-step stream for every tile ===================================
__device__ __forceinline__ void produce_weights(const Ctx& c, const float* __restrict__ wstream, int n_ksteps,
                                                long long n_tiles, uint32_t w_ring) {
    const uint4* s_steps = reinterpret_cast<const uint4*>(c.smem + kOffSteps);
    uint32_t slot = 0, phase = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint32_t w_off = 0;
        for (int ks = 0; ks < n_ksteps; ++ks) {
            const uint32_t w_bytes = s_steps[ks].w & 0xffffu;
            mbar_wait(c.bar_empty(slot), phase ^ 1);
            mbar_arrive_expect_tx(c.bar_full(slot), w_bytes);
            bulk_g2s(c.s_base + w_ring + slot * kWSlotBytes, reinterpret_cast<const uint8_t*>(wstream) + w_off, w_bytes,
                     c.bar_full(slot));
            w_off += w_bytes;
            if (++slot == kWStages) {
                slot = 0;
                phase ^= 1;
            }
        }
    }
}

// ===== MMA issuer: the whole warp walks the k-steps (uniform control flow and addresses), one elected lane
// issues.  Per k-step: 2 x (Ah.Bh + Al.Bh + Ah.Bl) over 8 channels each.  WAIT_IN: a tile starts when the
// epilogue warps have staged its encodings (forward kernels).
// FOUR: also accumulate Al.Bl (the value kernel: its SDF decides where the hierarchical sampling puts the samples, and
// the inverse CDF at inv_s up to 512 amplifies an SDF error of 1e-5 into sample depths moved by 1e-4; the fourth term
// brings the products to fp32 accuracy for +1/3 of that kernel's MMA work).
template <bool WAIT_IN, bool FOUR = false>
__device__ __forceinline__ void issue_mmas(const Ctx& c, int n_ksteps, long long n_tiles, uint32_t w_ring,
                                           long long* prof = nullptr) {
    long long t_wa = 0, t_wf = 0, t_is = 0, n_ks = 0;
    const uint4* s_steps = reinterpret_cast<const uint4*>(c.smem + kOffSteps);
    const uint32_t elected = elect_one();
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);  // SBO = 128 B, descriptor version 1 (bits 32-47)
    uint32_t slot = 0, phase = 0, in_phase = 0, a_phase = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (WAIT_IN) {
            mbar_wait(c.bar_in(), in_phase);
            in_phase ^= 1;
            tc_fence_after();
        }
        uint32_t acc_on = 0;
        uint4 rec = s_steps[0];
        for (int ks = 0; ks < n_ksteps; ++ks) {
            const uint4 cur = rec;
            if (ks + 1 < n_ksteps) rec = s_steps[ks + 1];
            const uint32_t flags = cur.z, nn = (flags >> 16) & 0x1ff;
            const uint32_t idesc = instr_desc(nn);
            const uint32_t d_tmem = c.tmem + (cur.w >> 16);
            if (flags & 1u) acc_on = 0;
            const long long q0 = prof ? clock64() : 0;
            if (flags & 4u) {  // the epilogue's A operand must be in tensor memory
                mbar_wait(c.bar_a(), a_phase);
                a_phase ^= 1;
            }
            const long long q1 = prof ? clock64() : 0;
            mbar_wait(c.bar_full(slot), phase);
            tc_fence_after();
            const long long q2 = prof ? clock64() : 0;
            // B: [hi: 4 chunks][lo: 4 chunks], chunk = N rows x 16 B
            const uint32_t w_addr = c.s_base + w_ring + slot * kWSlotBytes;
            const uint32_t b_lo32 = ((w_addr & 0x3ffff) >> 4) | (nn << 16);  // LBO = 16 N bytes
            const uint32_t b_step = 2 * nn, b_lopart = 4 * nn;               // in 16-byte units
            if (elected) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint64_t b_hi = ((uint64_t)desc_hi << 32) | (b_lo32 + j * b_step);
                    const uint64_t b_lo = ((uint64_t)desc_hi << 32) | (b_lo32 + b_lopart + j * b_step);
                    if (flags & 8u) {
                        const uint32_t a_hi = c.tmem + cur.x + j * 8, a_lo = c.tmem + cur.y + j * 8;
                        // smallest terms first: the accumulator adds with truncation
                        if (FOUR) {
                            mma_ts(d_tmem, a_lo, b_lo, idesc, acc_on);
                            mma_ts(d_tmem, a_lo, b_hi, idesc, 1);
                            mma_ts(d_tmem, a_hi, b_lo, idesc, 1);
                            mma_ts(d_tmem, a_hi, b_hi, idesc, 1);
                        } else {
                            mma_ts(d_tmem, a_hi, b_hi, idesc, acc_on);
                            mma_ts(d_tmem, a_lo, b_hi, idesc, 1);
                            mma_ts(d_tmem, a_hi, b_lo, idesc, 1);
                        }
                    } else {
                        const uint64_t a_hi = ((uint64_t)desc_hi << 32) | (cur.x + j * (2 * kChunkBytes >> 4));
                        const uint64_t a_lo = ((uint64_t)desc_hi << 32) | (cur.y + j * (2 * kChunkBytes >> 4));
                        if (FOUR) {
                            mma_ss(d_tmem, a_lo, b_lo, idesc, acc_on);
                            mma_ss(d_tmem, a_lo, b_hi, idesc, 1);
                            mma_ss(d_tmem, a_hi, b_lo, idesc, 1);
                            mma_ss(d_tmem, a_hi, b_hi, idesc, 1);
                        } else {
                            mma_ss(d_tmem, a_hi, b_hi, idesc, acc_on);
                            mma_ss(d_tmem, a_lo, b_hi, idesc, 1);
                            mma_ss(d_tmem, a_hi, b_lo, idesc, 1);
                        }
                    }
                    acc_on = 1;
                }
                tc_commit(c.bar_empty(slot));  // slot free once these MMAs have read it
                if (flags & 2u) tc_commit(c.bar_acc((flags >> 4) & 1));
            }
            acc_on = 1;
            __syncwarp();
            if (prof) {
                t_wa += q1 - q0;
                t_wf += q2 - q1;
                t_is += clock64() - q2;
                ++n_ks;
            }
            if (++slot == kWStages) {
                slot = 0;
                phase ^= 1;
            }
        }
    }
    if (prof && blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
        prof[8] = t_wa; prof[9] = t_wf; prof[10] = t_is; prof[11] = n_ks;
    }
}

// ------------------------------------------------------------------------------------------------------
// Forward kernels.  JVP = false: rows are points, pos (n,27), fe (n,100) -> sdf (n).
// JVP = true: row 2i / 2i+1 = primal / tangent of point i; pos (2n,27) and fe (2n,100) hold the primal rows
// [0,n) and the tangent rows [n,2n) (gens_sdf_encode); additionally stores, per hidden layer l and point p,
// sp'(a) and sp''(a) da of every channel into the tape (layout at tape_slot) for the reverse sweep.
template <bool JVP, bool FOUR = false>
__global__ void __launch_bounds__(kThreads, 1)
sdf_mlp_fwd_kernel(const float* __restrict__ pos, const float* __restrict__ fe, long long n,
                   const float* __restrict__ wstream, const KStep* __restrict__ ksteps, int n_ksteps,
                   const float* __restrict__ bias, int n_layers, float scale, float* __restrict__ sdf_out,
                   float* __restrict__ tape_out, long long* __restrict__ prof) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const long long k_start = prof ? clock64() : 0;
    const Ctx c = setup(smem, ksteps, n_ksteps, bias, n_layers);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kPts = JVP ? kTileM / 2 : kTileM;  // points per tile
    const long long n_tiles = (n + kPts - 1) / kPts;

    if (warp == kProducerWarp) {
        if (lane == 0) produce_weights(c, wstream, n_ksteps, n_tiles, kOffW);
    } else if (warp == kMmaWarp) {
        issue_mmas<true, FOUR>(c, n_ksteps, n_tiles, kOffW, prof);
    } else {
        // ===== input staging + epilogue (threads 0..511) ===================================================
        const float* s_bias = reinterpret_cast<const float*>(smem + kOffBias);
        const int row = threadIdx.x & 127, cq = threadIdx.x >> 7;  // column quarter: 32 columns
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const bool tangent = JVP && (row & 1);
        uint32_t acc_phase[2] = {0, 0};
        // measurement knob (gens_debug_tc_profile): cycles of block 0 / thread 0 per phase --
        // [0] staging the encodings, [1] waiting for a layer's MMAs, [2] epilogue arithmetic + tape stores + A stores,
        // [3] tiles, [7] layer epilogues
        const bool timing = prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
        long long pt[4] = {0, 0, 0, 0}, n_epi = 0, pt_fence = 0;
        // columns 27..31 of the position encoding are K padding: zero once, the staging below never touches them
        for (int i2 = threadIdx.x; i2 < kTileM * (kPChunks * 4 - kNP); i2 += kEpiThreads) {
            const int srow = i2 / (kPChunks * 4 - kNP), col = kNP + i2 % (kPChunks * 4 - kNP);
            const int at = (col >> 2) * kChunkBytes + srow * 16 + (col & 3) * 4;
            *reinterpret_cast<uint32_t*>(smem + kOffPhi + at) = 0u;
            *reinterpret_cast<uint32_t*>(smem + kOffPlo + at) = 0u;
        }
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long f0 = timing ? clock64() : 0;
            // -- encodings of this tile -> shared memory, split hi/lo (the previous tile's MMAs are done: this
            //    thread has already waited for its last accumulator)
            const long long p = JVP ? tile * kPts + (row >> 1) : tile * kPts + row;
            const bool live = p < n;
            const long long in_row = tangent ? n + p : p;
            // Cooperative, sector-exact staging.  The tile's encodings are contiguous row-major blocks in global memory
            // (JVP: a primal and a tangent block); who loads a value is independent of who owns the row later.  Measured
            // (gens_debug_tc_profile): with every thread loading its own row's chunks, a warp request touched 32 rows
            // 400 bytes apart -- half a sector used per 16-byte load, an eighth per scalar -- 7.7 k sector requests and
            // 9-14 k cycles per tile with the tensor core idle.  Now a warp request covers 8 rows x 64 contiguous bytes
            // of the feature encoding (16 full sectors; the 16-byte stores of one 8-row group are one conflict-free
            // wavefront per chunk), and the position encoding is read as one contiguous float stream.
            {
                const int r8 = lane & 7, c4 = lane >> 3;
                float4 fv[7];
#pragma unroll
                for (int it = 0; it < 7; ++it) {
                    const int u = warp + 16 * it, rg = u / 7, cg = u - 7 * rg;
                    const int srow = rg * 8 + r8, chunk = cg * 4 + c4;
                    const long long pt_ = JVP ? tile * kPts + (srow >> 1) : tile * kPts + srow;
                    const long long grow = (JVP && (srow & 1)) ? n + pt_ : pt_;
                    fv[it] = (pt_ < n && 4 * chunk + 3 < kNF) ? __ldg(reinterpret_cast<const float4*>(fe + grow * kNF) + chunk)
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                constexpr int kPosFloats = kTileM * kNP;  // 3456 per tile
                float pvv[7];
#pragma unroll
                for (int it = 0; it < 7; ++it) {
                    const int idx = (int)threadIdx.x + 512 * it;
                    float v = 0.0f;
                    if (idx < kPosFloats) {
                        if (JVP) {
                            const int b = idx / (kPosFloats / 2), jj = idx - b * (kPosFloats / 2);
                            const long long pt_ = tile * kPts + jj / kNP;
                            if (pt_ < n) v = __ldg(pos + ((b ? n : 0) + tile * kPts) * kNP + jj);
                        } else {
                            if (tile * kPts + idx / kNP < n) v = __ldg(pos + tile * kPts * kNP + idx);
                        }
                    }
                    pvv[it] = v;
                }
#pragma unroll
                for (int it = 0; it < 7; ++it) {
                    const int u = warp + 16 * it, rg = u / 7, cg = u - 7 * rg;
                    const int srow = rg * 8 + r8, chunk = cg * 4 + c4;
                    uint4 hi, lo;
                    split_tf32(fv[it].x, hi.x, lo.x);
                    split_tf32(fv[it].y, hi.y, lo.y);
                    split_tf32(fv[it].z, hi.z, lo.z);
                    split_tf32(fv[it].w, hi.w, lo.w);
                    *reinterpret_cast<uint4*>(smem + kOffFhi + chunk * kChunkBytes + srow * 16) = hi;
                    *reinterpret_cast<uint4*>(smem + kOffFlo + chunk * kChunkBytes + srow * 16) = lo;
                }
#pragma unroll
                for (int it = 0; it < 7; ++it) {
                    const int idx = (int)threadIdx.x + 512 * it;
                    if (idx < kPosFloats) {
                        int srow, col;
                        if (JVP) {
                            const int b = idx / (kPosFloats / 2), jj = idx - b * (kPosFloats / 2);
                            srow = 2 * (jj / kNP) + b;
                            col = jj % kNP;
                        } else {
                            srow = idx / kNP;
                            col = idx % kNP;
                        }
                        uint32_t hi, lo;
                        split_tf32(pvv[it], hi, lo);
                        const int at = (col >> 2) * kChunkBytes + srow * 16 + (col & 3) * 4;
                        *reinterpret_cast<uint32_t*>(smem + kOffPhi + at) = hi;
                        *reinterpret_cast<uint32_t*>(smem + kOffPlo + at) = lo;
                    }
                }
            }
            const long long f0b = timing ? clock64() : 0;
            fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async proxy
            tc_fence_before();    // orders this thread's earlier tcgen05.ld of the accumulators
            mbar_arrive(c.bar_in());
            if (timing) {
                pt[0] += clock64() - f0;
                pt_fence += clock64() - f0b;
                pt[3] += 1;
            }

            for (int layer = 0; layer < n_layers; ++layer) {
                const int st = layer & 1;
                if (layer == 1 && tile + gridDim.x < n_tiles) {
                    // pull the NEXT tile's encodings towards L2 now: its staging prologue is a burst of loads from all
                    // SMs at once during which the tensor core idles (11 k cycles per tile from HBM, measured)
                    const long long np = JVP ? (tile + gridDim.x) * kPts + (row >> 1) : (tile + gridDim.x) * kPts + row;
                    if (np < n) {
                        const long long nr = tangent ? n + np : np;
                        const float* f = fe + nr * kNF + 4 * cq * (kFChunks / 4);
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(f));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(f + 27));
                        if (cq == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(pos + nr * kNP));
                    }
                }
                const long long f1 = timing ? clock64() : 0;
                mbar_wait(c.bar_acc(st), acc_phase[st]);
                acc_phase[st] ^= 1;
                tc_fence_after();
                const long long f2 = timing ? clock64() : 0;
                const uint32_t acc = c.tmem + lane_base + (st ? kColAcc1 : kColAcc0);
                if (layer + 1 < n_layers) {
                    const float* b = s_bias + layer * 128;
#pragma unroll 1
                    for (int blk = 0; blk < 2; ++blk) {
                        const int col0 = cq * 32 + blk * 16;
                        uint32_t v[16], hi[16], lo[16];
                        tmem_ld16(acc + col0, v);
                        tmem_wait_ld();
                        if (!JVP) {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                split_tf32(softplus100(__uint_as_float(v[j]) + b[col0 + j]), hi[j], lo[j]);
                        } else {
                            // Lanes (2i, 2i+1) hold the primal / tangent row of one point.  Instead of every lane
                            // walking all 16 channels down its own branch (softplus on the primal lane, the two
                            // derivatives on the tangent lane: both branches serialised in every warp), the pair
                            // SPLITS the channels: the primal lane takes channels [0,8), the tangent lane [8,16),
                            // each evaluates softplus, sp' and sp'' (one shared exponential) for its eight
                            // (point, channel) entries, and one shuffle per entry hands the results back.
                            const int cb = col0 + (tangent ? 8 : 0);
                            float a8[8], da8[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const float v_lo = __uint_as_float(v[k]), v_hi = __uint_as_float(v[8 + k]);
                                // primal lane sends a[8+k] and receives da[k]; tangent lane sends da[k], receives a[8+k]
                                const float got = __shfl_xor_sync(0xffffffffu, tangent ? v_lo : v_hi, 1);
                                a8[k] = (tangent ? got : v_lo) + b[cb + k];
                                da8[k] = tangent ? v_hi : got;
                            }
                            float o_p[8], o_t[8], s1v[8], t2v[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                float d2;
                                softplus100_all(a8[k], o_p[k], s1v[k], d2);
                                o_t[k] = s1v[k] * da8[k];   // dh = sp'(a) da
                                t2v[k] = d2 * da8[k];       // sp''(a) da
                            }
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                // primal lane needs h of channels [8,16) (computed by the tangent lane), the tangent
                                // lane needs dh of channels [0,8) (computed by the primal lane)
                                const float got = __shfl_xor_sync(0xffffffffu, tangent ? o_p[k] : o_t[k], 1);
                                split_tf32(tangent ? got : o_p[k], hi[k], lo[k]);
                                split_tf32(tangent ? o_t[k] : got, hi[8 + k], lo[8 + k]);
                            }
                            if (live) {
                                uint8_t* blk_base = reinterpret_cast<uint8_t*>(tape_out) +
                                                    ((size_t)tile * (n_layers - 1) + layer) * (2u * kTapeBytes);
#pragma unroll
                                for (int q = 0; q < 2; ++q) {
                                    const uint32_t at = tape_slot(row >> 1, cb / 4 + q);
                                    __stcs(reinterpret_cast<float4*>(blk_base + at),
                                           make_float4(s1v[4 * q], s1v[4 * q + 1], s1v[4 * q + 2], s1v[4 * q + 3]));
                                    __stcs(reinterpret_cast<float4*>(blk_base + kTapeBytes + at),
                                           make_float4(t2v[4 * q], t2v[4 * q + 1], t2v[4 * q + 2], t2v[4 * q + 3]));
                                }
                            }
                        }
                        tmem_st16(c.tmem + lane_base + kColAhi + col0, hi);
                        tmem_st16(c.tmem + lane_base + kColAlo + col0, lo);
                    }
                    tmem_wait_st();
                    tc_fence_before();
                    mbar_arrive(c.bar_a());
                } else {
                    if (cq == 0) {
                        const uint32_t v = tmem_ld1(acc);
                        tmem_wait_ld();
                        if (live && !tangent) sdf_out[p] = __fdiv_rn(__uint_as_float(v) + s_bias[layer * 128], scale);
                    }
                }
                if (timing) {
                    pt[1] += f2 - f1;
                    pt[2] += clock64() - f2;
                    ++n_epi;
                }
            }
        }
        if (timing) {
            prof[0] = pt[0]; prof[1] = pt[1]; prof[2] = pt[2]; prof[3] = pt[3]; prof[4] = pt_fence; prof[7] = n_epi;
            prof[12] = clock64() - k_start;
        }
    }
    teardown(c);
}

// ------------------------------------------------------------------------------------------------------
// Reverse sweep (forward-over-reverse along u): for hidden layers l = L-1 .. 0
//     ga = sp' g_h ,  dga = sp''(a) da g_h + sp' dg_h            (epilogue, rows 2i / 2i+1 of a 64-point tile)
//     [g_h ; dg_h]_{l-1} = [ga ; dga] Wx_l        -> accumulator 0 (overwritten per layer)
//     [g_fe ; dg_fe]    += [ga ; dga] Wf_l        -> accumulator 1 (running sum over layers, l >= 1)
// Per layer the x-part MMAs are issued and committed first (barrier 0), the feature-part MMAs second (barrier 1): the
// epilogue of the next layer reads accumulator 0 and does its arithmetic while the feature part still runs, and only
// waits for barrier 1 before it overwrites the A operand.
// with g_h of the last hidden layer = w_out / scale (consts row 0) and the output layer's own feature part
// (consts row 1) added to the primal rows of g_fe at the end.  The skip layer returns its position part in
// columns [skip_col, skip_col + 27) of accumulator 0; together with layer 0's result it forms g_pos.
// Outputs (sdf_analytic's layout, primal rows [0,n), tangent rows [n,2n)): g_pos (2n,27), g_fe (2n,100).
__global__ void __launch_bounds__(kThreads, 1)
sdf_mlp_rev_kernel(const float* __restrict__ tape, long long n,
                   const float* __restrict__ wstream, const KStep* __restrict__ ksteps, int n_ksteps,
                   const float* __restrict__ consts, int n_hidden, int skip_layer, int skip_col,
                   float* __restrict__ g_pos, float* __restrict__ g_fe, long long* __restrict__ prof) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const long long k_start = prof ? clock64() : 0;
    const Ctx c = setup(smem, ksteps, n_ksteps, consts, 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kPts = kTileM / 2;
    const long long n_tiles = (n + kPts - 1) / kPts;

    if (warp == kProducerWarp) {
        if (lane == 0) produce_weights(c, wstream, n_ksteps, n_tiles, kRevOffW);
    } else if (warp == kMmaWarp) {
        issue_mmas<false>(c, n_ksteps, n_tiles, kRevOffW, prof);
    } else {
        const float* s_const = reinterpret_cast<const float*>(smem + kOffBias);
        const int row = threadIdx.x & 127, cq = threadIdx.x >> 7;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const bool tangent = row & 1;
        uint32_t acc_phase = 0, afree_phase = 0;
        float* s_skip = reinterpret_cast<float*>(smem + kRevOffSkip);
        const bool timing = prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;  // measurement knob, see below
        long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        // The tapes (sp', sp'' da) of the NEXT layer are fetched while the tensor core works on the current one: ONE
        // cp.async.bulk of the layer's 64 KB block, issued by epilogue thread 0, completion on an mbarrier.  (Measured
        // with gens_debug_tc_profile: per-thread cp.async of the own 16-byte pieces took 39 % of a layer's epilogue
        // time just to issue, cooperative coalesced cp.async still 29 %.)  Two staging buffers alternate.  A buffer is
        // overwritten two layers after it was read: by then every epilogue thread has arrived on bar_a of the layer in
        // between, which thread 0 observes through that layer's MMA commit before it gets here.
        const uint32_t stage0 = c.s_base + kRevOffStage;
        uint32_t pf_buf = 0, use_buf = 0, tape_phase[2] = {0, 0};
        auto prefetch = [&](long long tile_, int layer) {
            if (threadIdx.x == 0) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(tape) +
                                     ((size_t)tile_ * n_hidden + layer) * (2u * kTapeBytes);
                mbar_arrive_expect_tx(c.bar_tape(pf_buf), 2u * kTapeBytes);
                bulk_g2s(stage0 + pf_buf * (2u * kTapeBytes), src, 2u * kTapeBytes, c.bar_tape(pf_buf));
            }
            pf_buf ^= 1u;
        };
        if ((long long)blockIdx.x < n_tiles) prefetch((long long)blockIdx.x, n_hidden - 1);
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long p = tile * kPts + (row >> 1);
            const bool live = p < n;
            const long long out_row = tangent ? n + p : p;
            float pos_part[16];  // the skip layer's position cotangents (threads owning those columns)
#pragma unroll
            for (int j = 0; j < 16; ++j) pos_part[j] = 0.0f;
            for (int layer = n_hidden - 1; layer >= 0; --layer) {
                const bool top = layer == n_hidden - 1;
                const long long e0 = timing ? clock64() : 0;
                if (!top) {
                    mbar_wait(c.bar_acc(0), acc_phase);
                    acc_phase ^= 1;
                    tc_fence_after();
                }
                long long e1 = timing ? clock64() : 0, e2 = e1, e3 = e1, e4 = e1;
                // Both 16-column blocks are loaded and differentiated BEFORE the wait for the feature-part MMAs of the
                // layer above (which still read the A operand), so that only the TF32 split and the tcgen05.st of the
                // new A operand remain behind that wait (before: the second block's load and arithmetic as well).
                float outv[2][16];
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
                    const int col0 = cq * 32 + blk * 16;
                    uint32_t v[16];
                    if (top) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(tangent ? 0.0f : s_const[col0 + j]);
                    } else {
                        tmem_ld16(c.tmem + lane_base + kColAcc0 + col0, v);
                    }
                    float d1[16], d2[16];
                    {
                        if (blk == 0) {
                            mbar_wait(c.bar_tape(use_buf), tape_phase[use_buf]);  // this layer's tapes (issued one layer ago)
                            tape_phase[use_buf] ^= 1u;
                            if (timing) e2 = clock64();
                        }
                        const uint32_t src = stage0 + use_buf * (2u * kTapeBytes);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t at = tape_slot(row >> 1, cq * 8 + blk * 4 + q);
                            const float4 a = live ? lds128(src + at) : make_float4(0.f, 0.f, 0.f, 0.f);
                            d1[4 * q] = a.x; d1[4 * q + 1] = a.y; d1[4 * q + 2] = a.z; d1[4 * q + 3] = a.w;
                            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (live && tangent) b = lds128(src + kTapeBytes + at);
                            d2[4 * q] = b.x; d2[4 * q + 1] = b.y; d2[4 * q + 2] = b.z; d2[4 * q + 3] = b.w;
                        }
                    }
                    if (!top) tmem_wait_ld();
                    // the skip layer's x-part covers [h | pos]: keep the pos columns, they are not activations.  They
                    // are owned by other threads than the ones that finish g_pos at the end of the tile, so they are
                    // handed over through shared memory ([row][27] floats; a round trip through g_pos in global memory
                    // put ~16 dependent loads per thread into the tile's tail: a third of the tile time).  Ordering:
                    // this thread's bar_a arrival below, the MMA commits in between, the readers' bar_acc wait.
                    if (layer == skip_layer - 1 && col0 + 16 > skip_col) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int cc = col0 + j - skip_col;
                            if (cc >= 0 && cc < kNP) s_skip[row * kNP + cc] = __uint_as_float(v[j]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float mine = __uint_as_float(v[j]);                    // g_h (primal) / dg_h (tangent)
                        const float other = __shfl_xor_sync(0xffffffffu, mine, 1);  // the pair's other row
                        outv[blk][j] = tangent ? fmaf(d2[j], other, d1[j] * mine) : d1[j] * mine;
                    }
                }
                if (timing) e3 = clock64();
                if (!top) {
                    // The layer above committed its x-part (barrier 0: accumulator 0 readable) BEFORE its
                    // feature-part MMAs, which still read the A operand this thread is about to overwrite and ran
                    // while the values above were loaded and computed; barrier 1 says they are done.
                    mbar_wait(c.bar_acc(1), afree_phase);
                    afree_phase ^= 1;
                    tc_fence_after();
                }
                if (timing) e4 = clock64();
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
                    const int col0 = cq * 32 + blk * 16;
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) split_tf32(outv[blk][j], hi[j], lo[j]);
                    tmem_st16(c.tmem + lane_base + kColAhi + col0, hi);
                    tmem_st16(c.tmem + lane_base + kColAlo + col0, lo);
                }
                (void)pos_part;
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(c.bar_a());
                const long long e5 = timing ? clock64() : 0;
                use_buf ^= 1u;
                // fetch the next layer's tapes (or the next tile's top layer) into the other staging buffer
                if (layer > 0) prefetch(tile, layer - 1);
                else if (tile + gridDim.x < n_tiles) prefetch(tile + gridDim.x, n_hidden - 1);
                if (timing) {
                    pt[0] += e1 - e0;            // wait for the x-part MMAs of the layer above
                    pt[1] += e2 - e1;            // wait for this layer's s1 / t2 (cp.async)
                    pt[2] += e3 - e2;            // accumulator load + arithmetic of both 16-column blocks
                    pt[3] += e4 - e3;            // wait for the feature-part MMAs (A operand free)
                    pt[4] += e5 - e4;            // TF32 split, stores, store fence, arrive
                    pt[5] += clock64() - e5;     // prefetch issue
                    pt[7] += 1;
                }
            }
            // -- results: accumulator 0 = layer 0's position cotangents, accumulator 1 = feature cotangents
            const long long tail0 = timing ? clock64() : 0;
            mbar_wait(c.bar_acc(0), acc_phase);
            acc_phase ^= 1;
            tc_fence_after();
            // The results leave through shared memory: rows of g_fe / g_pos are 400 / 108 bytes apart, so stores from the
            // row-owning threads touch one sector per lane (6.6 k sector requests per tile, most of the 11 k-cycle tail
            // the phase timers showed), while a tile's primal rows -- and its tangent rows -- are ONE contiguous block
            // each in global memory.  The tape buffer layer 0 has just finished with is free (every thread arrived on
            // bar_a before the MMAs awaited above could start): [half][point][100] floats, then [half][point][27].
            const uint32_t obuf = stage0 + (use_buf ^ 1u) * (2u * kTapeBytes);
            constexpr uint32_t kOutPosOff = 2u * 64u * kNF * 4u;
            const uint32_t orow = (uint32_t)((row & 1) * 64 + (row >> 1));
#pragma unroll 1
            for (int blk = 0; blk < 2; ++blk) {
                const int col0 = cq * 32 + blk * 16;
                uint32_t v[16];
                if (col0 < kNP) {  // warp-uniform
                    tmem_ld16(c.tmem + lane_base + kColAcc0 + col0, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (col0 + j < kNP) {
                            const float o = __uint_as_float(v[j]) + (skip_layer > 0 ? s_skip[row * kNP + col0 + j] : 0.0f);
                            asm volatile("st.shared.f32 [%0], %1;" ::"r"(obuf + kOutPosOff + (orow * kNP + col0 + j) * 4u), "f"(o));
                        }
                }
                if (col0 < kNF) {
                    tmem_ld16(c.tmem + lane_base + kColAcc1 + col0, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        if (col0 + j < kNF) {
                            float4 o;
                            o.x = __uint_as_float(v[j]) + (tangent ? 0.0f : s_const[128 + col0 + j]);
                            o.y = __uint_as_float(v[j + 1]) + (tangent ? 0.0f : s_const[128 + col0 + j + 1]);
                            o.z = __uint_as_float(v[j + 2]) + (tangent ? 0.0f : s_const[128 + col0 + j + 2]);
                            o.w = __uint_as_float(v[j + 3]) + (tangent ? 0.0f : s_const[128 + col0 + j + 3]);
                            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(obuf + (orow * kNF + col0 + j) * 4u),
                                         "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w));
                        }
                }
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            {
                const long long first = tile * kPts;
                const long long left = n - first;
                const int pts_here = left < kPts ? (int)left : kPts;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const long long grow = (h ? n : 0) + first;
                    float4* dfe = reinterpret_cast<float4*>(g_fe + grow * kNF);
                    for (int idx = threadIdx.x; idx < pts_here * (kNF / 4); idx += kEpiThreads)
                        dfe[idx] = lds128(obuf + (uint32_t)(h * 64 * kNF * 4) + (uint32_t)idx * 16u);
                    float* dpo = g_pos + grow * kNP;
                    for (int idx = threadIdx.x; idx < pts_here * kNP; idx += kEpiThreads) {
                        float o;
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o) : "r"(obuf + kOutPosOff + (uint32_t)(h * 64 * kNP + idx) * 4u));
                        dpo[idx] = o;
                    }
                }
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");  // the buffer is a prefetch target again from here on
            tc_fence_before();  // accumulator reads done before the next tile's MMAs may overwrite them
            if (timing) pt[6] += clock64() - tail0;
        }
        if (timing) {
            for (int i = 0; i < 8; ++i) prof[i] = pt[i];
            prof[12] = clock64() - k_start;
        }
    }
    teardown(c);
}

}  // namespace

namespace {
template <typename K>
int set_smem(K kernel) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    return e == cudaSuccess ? 0 : (int)e;
}
}  // namespace

namespace {
int g_tc_value_terms = 3;
long long* g_tc_prof = nullptr;
int g_tc_prof_target = 2;  // 0 value kernel, 1 JVP forward, 2 reverse
}
// Measurement knob: a device buffer of 16 int64 that the next launches of one kernel (target 0 value, 1 JVP forward,
// 2 reverse; the forward kernels' layout is documented at their timers) fill with cycle counts of block 0 -- reverse:
// (epilogue thread 0: [0] wait x-part MMAs, [1] wait s1/t2, [2] load + arithmetic, [3] wait feature-part MMAs,
// [4] stores + second half + arrive, [5] prefetch issue, [6] tile tails (last MMA wait + result stores), [7] layers; MMA issuer: [8] wait A operand, [9] wait weights,
// [10] issue, [11] k-steps; [12] kernel cycles).  nullptr switches it off.
extern "C" int gens_debug_tc_profile(long long* buf, int target) {
    g_tc_prof = buf;
    g_tc_prof_target = target;
    return 0;
}
extern "C" int gens_debug_set_tc_terms(int terms) {
    if (terms != 3 && terms != 4) return GENS_E_BADARG;
    g_tc_value_terms = terms;
    return 0;
}

// Value pass of the SDF MLP on the tensor cores.  pos (n,27) / fe (n,100): the encodings produced by
// gens_sdf_encode; wstream / ksteps / bias: the packed network (gens_b200/mlp_tc.py documents the format);
// sdf_out (n).  n_sm = number of CTAs to launch (<= SM count; one persistent CTA per SM).
extern "C" int gens_sdf_mlp_value_tc(const float* pos, const float* fe, long long n, const float* wstream,
                                     const void* ksteps, int n_ksteps, const float* bias, int n_layers, float scale,
                                     int n_sm, float* sdf_out, void* stream) {
    GENS_CHECK_ARG(pos && fe && wstream && ksteps && bias && sdf_out && n >= 0 && n_sm > 0);
    if (n_ksteps <= 0 || n_ksteps > kMaxKSteps || n_layers <= 0 || n_layers > kMaxLayers || scale == 0.f)
        return GENS_E_UNSUPPORTED;
    if (n == 0) return 0;
    const long long tiles = (n + kTileM - 1) / kTileM;
    const int grid = (int)(tiles < n_sm ? tiles : n_sm);
    if (g_tc_value_terms == 4) {  // measurement knob (gens_debug_set_tc_terms): fourth compensation term Al.Bl
        if (int rc = set_smem(sdf_mlp_fwd_kernel<false, true>)) return rc;
        sdf_mlp_fwd_kernel<false, true><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(
            pos, fe, n, wstream, reinterpret_cast<const KStep*>(ksteps), n_ksteps, bias, n_layers, scale, sdf_out,
            nullptr, nullptr);
        return gens_launch_status();
    }
    if (int rc = set_smem(sdf_mlp_fwd_kernel<false>)) return rc;
    sdf_mlp_fwd_kernel<false><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(
        pos, fe, n, wstream, reinterpret_cast<const KStep*>(ksteps), n_ksteps, bias, n_layers, scale, sdf_out, nullptr,
        g_tc_prof_target == 0 ? g_tc_prof : nullptr);
    return gens_launch_status();
}

// Value + tangent pass: pos (2n,27) / fe (2n,100) with the tangent rows (directional derivative of the
// encodings along u) in [n,2n); sdf_out (n); tape_out: ceil(n/64) x (n_layers-1) blocks of 64 KB holding sp'(a) and
// sp''(a) da of every hidden channel in the layout the reverse kernel stages verbatim (tape_slot above).
extern "C" int gens_sdf_mlp_jvp_tc(const float* pos, const float* fe, long long n, const float* wstream,
                                   const void* ksteps, int n_ksteps, const float* bias, int n_layers, float scale,
                                   int n_sm, float* sdf_out, float* tape_out, void* stream) {
    GENS_CHECK_ARG(pos && fe && wstream && ksteps && bias && sdf_out && tape_out && n >= 0 && n_sm > 0);
    if (n_ksteps <= 0 || n_ksteps > kMaxKSteps || n_layers <= 0 || n_layers > kMaxLayers || scale == 0.f)
        return GENS_E_UNSUPPORTED;
    if (n == 0) return 0;
    if (int rc = set_smem(sdf_mlp_fwd_kernel<true>)) return rc;
    const long long tiles = (n + kTileM / 2 - 1) / (kTileM / 2);
    const int grid = (int)(tiles < n_sm ? tiles : n_sm);
    sdf_mlp_fwd_kernel<true><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(
        pos, fe, n, wstream, reinterpret_cast<const KStep*>(ksteps), n_ksteps, bias, n_layers, scale, sdf_out, tape_out,
        g_tc_prof_target == 1 ? g_tc_prof : nullptr);
    return gens_launch_status();
}

// Reverse sweep: tape from gens_sdf_mlp_jvp_tc; wstream / ksteps = the transposed network
// (mlp_tc.PackedSDFReverse); consts (2,128): row 0 = output-layer weights of the last hidden activations /
// scale, row 1 = output-layer weights of the feature encoding / scale.  skip_layer / skip_col: the layer whose
// input concatenates the position encoding and the column where it starts.  g_pos (2n,27), g_fe (2n,100).
extern "C" int gens_sdf_mlp_rev_tc(const float* tape, long long n, const float* wstream,
                                   const void* ksteps, int n_ksteps, const float* consts, int n_hidden, int skip_layer,
                                   int skip_col, int n_sm, float* g_pos, float* g_fe, void* stream) {
    GENS_CHECK_ARG(tape && wstream && ksteps && consts && g_pos && g_fe && n >= 0 && n_sm > 0);
    if (n_ksteps <= 0 || n_ksteps > kMaxKSteps || n_hidden <= 0 || n_hidden >= kMaxLayers || skip_col < 0 ||
        skip_col + kNP > 128)
        return GENS_E_UNSUPPORTED;
    if (n == 0) return 0;
    if (int rc = set_smem(sdf_mlp_rev_kernel)) return rc;
    const long long tiles = (n + kTileM / 2 - 1) / (kTileM / 2);
    const int grid = (int)(tiles < n_sm ? tiles : n_sm);
    sdf_mlp_rev_kernel<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(
        tape, n, wstream, reinterpret_cast<const KStep*>(ksteps), n_ksteps, consts, n_hidden, skip_layer, skip_col,
        g_pos, g_fe, g_tc_prof_target == 2 ? g_tc_prof : nullptr);
    return gens_launch_status();
}
