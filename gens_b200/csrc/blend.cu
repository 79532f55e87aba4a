// K10: the colour-blending network over the source views as ONE kernel (sm_100a).
//
// Replaces, for inference, BlendingNetwork.forward (reference models/modules/blending_network.py:69-117): eleven
// small Linear layers on (n * n_src) rows plus ~40 element-wise / concat / reduction launches, all of whose
// activations travel through HBM (27 % of a 16 k-ray render chunk, profiles/r01_launches_bench.txt).  Here one
// thread owns one sample point, walks its source views, and keeps every activation in registers; the 11 k
// weights sit in the constant bank (or shared memory) in the order the loops read them:
//   * "A" layers (inputs in registers) produce their outputs four at a time in a rolled loop,
//   * the following "B" layer accumulates those four activations into its statically indexed outputs,
// which keeps the code a few thousand instructions instead of the 19 k a full unroll would need (unrolling the block
// loops by 2 / 4 is slower: 5.28 / 7.79 ms against 5.10 ms).
// Outputs are held as PAIRS (f32x2): one packed FFMA2 takes a pair of weights straight from a 64-bit uniform
// register (LDCU.64 -> UR.F32x2) and the input activation as a scalar broadcast, so every weight fetch feeds one
// instruction that does two FMAs; each half rounds exactly like the scalar fmaf, in the same order as before.
// The view-independent half of base_fc's first layer ([mean, var] -> 64) is evaluated once per point and
// parked in a per-thread shared-memory column.
// Bound: instruction issue (18.6 kMAC per point at n_src = 2); HBM traffic is 232 B read + 12 B written per point.
#include "common.cuh"
#include "f32x2.cuh"

namespace {

constexpr int kC = 23;           // 3 rgb + 20 feature channels per view
constexpr int kThreads = 192;
constexpr int kMaxSrc = 8;

// packed-weight offsets, in floats (gens_b200/networks.py:pack_blending builds exactly this image).
// A layer: [out/4][in] float4 (four outputs of one input).  B layer: [in/4][ceil(out/2)] x 8 floats
// (x_j, x_j+1, y_j, y_j+1, z_j, z_j+1, w_j, w_j+1: the pair of outputs (j, j+1) for the four inputs x..w of the block;
// an odd output count is padded with a zero-weight output).
constexpr int oW1 = 0;                    // ray_dir_fc.0   A  [4 blocks][4 in] float4
constexpr int oB1 = oW1 + 4 * 4 * 4;      //                   [16]
constexpr int oW2 = oB1 + 16;             // ray_dir_fc.2   B  [4 blocks][12 pairs] 8 floats
constexpr int oB2 = oW2 + 4 * 12 * 8;     //                   [23] (+1 pad)
constexpr int oW3s = oB2 + 24;            // base_fc.0 (mean,var part)  A [16][46] float4
constexpr int oB3 = oW3s + 16 * 46 * 4;   //                   [64]
constexpr int oW3f = oB3 + 64;            // base_fc.0 (per-view part)  A [16][23] float4
constexpr int oW4 = oW3f + 16 * kC * 4;   // base_fc.2      B  [16][16 pairs] 8 floats
constexpr int oB4 = oW4 + 16 * 16 * 8;    //                   [32]
constexpr int oW5 = oB4 + 32;             // vis_fc.0       A  [8][32] float4
constexpr int oB5 = oW5 + 8 * 32 * 4;     //                   [32]
constexpr int oW6 = oB5 + 32;             // vis_fc.2       B  [8][17 pairs] 8 floats
constexpr int oB6 = oW6 + 8 * 17 * 8;     //                   [33] (+3 pad)
constexpr int oW7 = oB6 + 36;             // vis_fc2.0      A  [8][32] float4
constexpr int oB7 = oW7 + 8 * 32 * 4;     //                   [32]
constexpr int oW8 = oB7 + 32;             // vis_fc2.2      B  [8][1 pair] 8 floats
constexpr int oB8 = oW8 + 8 * 8;          //                   [1] (+3 pad)
constexpr int oW9 = oB8 + 4;              // rgb_fc.0       A  [4][37] float4
constexpr int oB9 = oW9 + 4 * 37 * 4;     //                   [16]
constexpr int oW10 = oB9 + 16;            // rgb_fc.2       B  [4][4 pairs] 8 floats
constexpr int oB10 = oW10 + 4 * 4 * 8;    //                   [8]
constexpr int oW11 = oB10 + 8;            // rgb_fc.4          [8]
constexpr int oB11 = oW11 + 8;            //                   [1] ; then |s| of the anti-alias pooling, 2 pad
constexpr int oS = oB11 + 1;
constexpr int kWeightFloats = oS + 3;
static_assert(kWeightFloats % 4 == 0, "weight image must be float4 granular");

// The weight image in the constant bank (shipped; gens_debug_blend_const(0) selects the shared-memory variant):
// broadcast weights arrive as uniform-register operands instead of LDS.128 through the SM's L1 data pipe (69 % busy in
// the shared-memory variant under ncu), the kernel needs 128 registers instead of 168 with spills, and two blocks
// of 256 threads fit an SM.  Published per call by one stream-ordered device-to-device copy: launches that use
// DIFFERENT weights must not overlap on different streams of one device.
__constant__ __align__(16) float c_blend[kWeightFloats];

typedef f32x2 V;  // two OUTPUTS of a layer

__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// ELU without the branch: elu(x) = max(x, 0) + (2^(min(x, 0) log2 e) - 1).  ex2.approx is accurate to 2^-22 relative,
// i.e. |error| < 3e-7 absolute here (the branchy expf version cost ~13 issue slots per activation, this one 4.5).
__device__ __forceinline__ float elu(float x) {
    return fmaxf(x, 0.f) + (ex2_approx(fminf(x, 0.f) * 1.4426950408889634f) - 1.0f);
}
__device__ __forceinline__ V elu2(V x) {
    const float a = lo(x), b = hi(x);
    const V t = mul2(pk(fminf(a, 0.f), fminf(b, 0.f)), bc(1.4426950408889634f));
    const V e = add2(pk(ex2_approx(lo(t)), ex2_approx(hi(t))), bc(-1.0f));
    return add2(pk(fmaxf(a, 0.f), fmaxf(b, 0.f)), e);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

struct Quad {  // the four outputs of one A block
    V xy, zw;
};
__device__ __forceinline__ Quad elu4(Quad a) { return Quad{elu2(a.xy), elu2(a.zw)}; }
__device__ __forceinline__ float4 ld4(const float* p, int i) { return reinterpret_cast<const float4*>(p)[i]; }
__device__ __forceinline__ Quad ldq(const float* p, int i) {
    const float4 b = ld4(p, i);
    return Quad{pk(b.x, b.y), pk(b.z, b.w)};
}
__device__ __forceinline__ V ld2(const float* p, int i) { return pk(p[2 * i], p[2 * i + 1]); }

// four outputs (block ob) of an A layer whose IN inputs are in registers; w = [blocks][IN] float4
template <int IN>
__device__ __forceinline__ Quad a_block(const float4* __restrict__ w, int ob, const float (&x)[IN], Quad acc) {
    const float4* p = w + ob * IN;
#pragma unroll
    for (int i = 0; i < IN; ++i) {
        const float4 ww = p[i];
        acc.xy = fma2(pk(ww.x, ww.y), bc(x[i]), acc.xy);
        acc.zw = fma2(pk(ww.z, ww.w), bc(x[i]), acc.zw);
    }
    return acc;
}
// accumulate the four activations of block ob into the 2 * PAIRS outputs of a B layer; w = [blocks][PAIRS] 8 floats.
// Per output: acc += x.wx, then + y.wy ... innermost first, the order the scalar kernel of round 1 used.
template <int PAIRS>
__device__ __forceinline__ void b_accum(const float4* __restrict__ w, int ob, Quad a, V (&acc)[PAIRS]) {
    const float4* p = w + ob * (2 * PAIRS);
    const float ax = lo(a.xy), ay = hi(a.xy), az = lo(a.zw), aw = hi(a.zw);
#pragma unroll
    for (int j = 0; j < PAIRS; ++j) {
        const float4 wa = p[2 * j], wb = p[2 * j + 1];
        acc[j] = fma2(pk(wa.x, wa.y), bc(ax),
                      fma2(pk(wa.z, wa.w), bc(ay), fma2(pk(wb.x, wb.y), bc(az), fma2(pk(wb.z, wb.w), bc(aw), acc[j]))));
    }
}

// feat = rgb_feat + ray_dir_fc(ray_diff)   (blending_network.py:77-78)
__device__ __forceinline__ void view_features(const float* __restrict__ W, const float* __restrict__ rf,
                                              const float (&rd)[4], float (&feat)[kC]) {
    V acc[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = ld2(W + oB2, j);
#pragma unroll 1
    for (int ob = 0; ob < 4; ++ob) {
        const Quad a = elu4(a_block<4>(reinterpret_cast<const float4*>(W + oW1), ob, rd, ldq(W + oB1, ob)));
        b_accum<12>(reinterpret_cast<const float4*>(W + oW2), ob, a, acc);
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        const V e = elu2(acc[j]);
        feat[2 * j] = rf[2 * j] + lo(e);
        if (2 * j + 1 < kC) feat[2 * j + 1] = rf[2 * j + 1] + hi(e);
    }
}

template <bool CONSTW, int THREADS>
__global__ void __launch_bounds__(THREADS, CONSTW ? 512 / THREADS : 2)
blend_kernel(const float* __restrict__ rgb_feat, const float* __restrict__ ray_diff, const uint8_t* __restrict__ mask,
             long long n, int ns, const float* __restrict__ weights, float* __restrict__ rgb_out) {
    extern __shared__ __align__(16) float smem[];
    const float* W = CONSTW ? c_blend : smem;                 // kWeightFloats
    V* pre3 = reinterpret_cast<V*>(smem + (CONSTW ? 0 : kWeightFloats));  // [32 pairs][THREADS]: view-independent half of base_fc.0
    if (!CONSTW) {
        for (int i = threadIdx.x; i < kWeightFloats / 4; i += THREADS)
            reinterpret_cast<float4*>(smem)[i] = __ldg(reinterpret_cast<const float4*>(weights) + i);
        __syncthreads();
    }
    const long long pt = (long long)blockIdx.x * THREADS + threadIdx.x;
    if (pt >= n) return;
    const float* rf0 = rgb_feat + pt * ns * kC;
    const float* rd0 = ray_diff + pt * ns * 4;
    const uint8_t* m0 = mask + pt * ns;

    // anti-alias pooling weights over the views (:79-83)
    const float s_abs = W[oS];
    float wv[kMaxSrc];
    {
        float emin = 3.4e38f;
        for (int v = 0; v < ns; ++v) {
            wv[v] = expf(s_abs * (rd0[4 * v + 3] - 1.0f));
            emin = fminf(emin, wv[v]);
        }
        float sum = 0.f;
        for (int v = 0; v < ns; ++v) {
            wv[v] = (wv[v] - emin) * (m0[v] ? 1.0f : 0.0f);
            sum += wv[v];
        }
        for (int v = 0; v < ns; ++v) wv[v] = wv[v] / (sum + 1e-8f);
    }

    // weighted mean and variance of the per-view features (:88-89), two passes as the reference
    float mv[2 * kC];
#pragma unroll
    for (int j = 0; j < 2 * kC; ++j) mv[j] = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll 1
        for (int v = 0; v < ns; ++v) {
            const float4 r4 = ld4(rd0, v);
            const float rd[4] = {r4.x, r4.y, r4.z, r4.w};
            float feat[kC];
            view_features(W, rf0 + v * kC, rd, feat);
            const float w = wv[v];
            if (pass == 0) {
#pragma unroll
                for (int j = 0; j < kC; ++j) mv[j] = fmaf(feat[j], w, mv[j]);
            } else {
#pragma unroll
                for (int j = 0; j < kC; ++j) {
                    const float d = feat[j] - mv[j];
                    mv[kC + j] = fmaf(w, d * d, mv[kC + j]);
                }
            }
        }
    }
    // view-independent half of base_fc.0: b3 + W3[:, :46] . [mean, var]  -> shared-memory column of this thread
#pragma unroll 1
    for (int ob = 0; ob < 16; ++ob) {
        const Quad a = a_block<2 * kC>(reinterpret_cast<const float4*>(W + oW3s), ob, mv, ldq(W + oB3, ob));
        pre3[(2 * ob + 0) * THREADS + threadIdx.x] = a.xy;
        pre3[(2 * ob + 1) * THREADS + threadIdx.x] = a.zw;
    }

    float logit[kMaxSrc];
#pragma unroll 1
    for (int v = 0; v < ns; ++v) {
        const float4 r4 = ld4(rd0, v);
        const float rd[4] = {r4.x, r4.y, r4.z, r4.w};
        const float m = m0[v] ? 1.0f : 0.0f;
        V x[16];
        {   // base_fc: 69 -> 64 -> 32 (:91)
            float feat[kC];
            view_features(W, rf0 + v * kC, rd, feat);
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = ld2(W + oB4, j);
#pragma unroll 1
            for (int ob = 0; ob < 16; ++ob) {
                const Quad init = Quad{pre3[(2 * ob + 0) * THREADS + threadIdx.x], pre3[(2 * ob + 1) * THREADS + threadIdx.x]};
                const Quad a = elu4(a_block<kC>(reinterpret_cast<const float4*>(W + oW3f), ob, feat, init));
                b_accum<16>(reinterpret_cast<const float4*>(W + oW4), ob, a, x);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = elu2(x[j]);
        }
        float vis;
        {   // vis_fc on x * weight: 32 -> 32 -> 33; residual + visibility (:92-95)
            float xin[32];
            V y[17];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const V t = mul2(x[j], bc(wv[v]));
                xin[2 * j] = lo(t);
                xin[2 * j + 1] = hi(t);
            }
#pragma unroll
            for (int j = 0; j < 17; ++j) y[j] = ld2(W + oB6, j);
#pragma unroll 1
            for (int ob = 0; ob < 8; ++ob) {
                const Quad a = elu4(a_block<32>(reinterpret_cast<const float4*>(W + oW5), ob, xin, ldq(W + oB5, ob)));
                b_accum<17>(reinterpret_cast<const float4*>(W + oW6), ob, a, y);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = add2(x[j], elu2(y[j]));
            vis = sigmoidf_(elu(lo(y[16]))) * m;
        }
        {   // vis_fc2 on x * vis: 32 -> 32 -> 1, sigmoid (:96)
            float xin[32];
            V y[1] = {pk(W[oB8], 0.f)};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const V t = mul2(x[j], bc(vis));
                xin[2 * j] = lo(t);
                xin[2 * j + 1] = hi(t);
            }
#pragma unroll 1
            for (int ob = 0; ob < 8; ++ob) {
                const Quad a = elu4(a_block<32>(reinterpret_cast<const float4*>(W + oW7), ob, xin, ldq(W + oB7, ob)));
                b_accum<1>(reinterpret_cast<const float4*>(W + oW8), ob, a, y);
            }
            vis = sigmoidf_(lo(y[0])) * m;
        }
        {   // rgb_fc on [x, vis, ray_diff]: 37 -> 16 -> 8 -> 1 (:99-100)
            float xin[37];
            V y[4];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                xin[2 * j] = lo(x[j]);
                xin[2 * j + 1] = hi(x[j]);
            }
            xin[32] = vis;
#pragma unroll
            for (int j = 0; j < 4; ++j) xin[33 + j] = rd[j];
#pragma unroll
            for (int j = 0; j < 4; ++j) y[j] = ld2(W + oB10, j);
#pragma unroll 1
            for (int ob = 0; ob < 4; ++ob) {
                const Quad a = elu4(a_block<37>(reinterpret_cast<const float4*>(W + oW9), ob, xin, ldq(W + oB9, ob)));
                b_accum<4>(reinterpret_cast<const float4*>(W + oW10), ob, a, y);
            }
            float l = W[oB11];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const V e = elu2(y[j]);
                l = fmaf(W[oW11 + 2 * j], lo(e), l);
                l = fmaf(W[oW11 + 2 * j + 1], hi(e), l);
            }
            logit[v] = m != 0.f ? l : -1e9f;
        }
    }
    // softmax over the views, blend the sampled colours (:101-103)
    float lmax = -3.4e38f;
    for (int v = 0; v < ns; ++v) lmax = fmaxf(lmax, logit[v]);
    float den = 0.f, r = 0.f, g = 0.f, b = 0.f;
    for (int v = 0; v < ns; ++v) {
        const float e = expf(logit[v] - lmax);
        den += e;
        r = fmaf(rf0[v * kC], e, r);
        g = fmaf(rf0[v * kC + 1], e, g);
        b = fmaf(rf0[v * kC + 2], e, b);
    }
    rgb_out[3 * pt] = r / den;
    rgb_out[3 * pt + 1] = g / den;
    rgb_out[3 * pt + 2] = b / den;
}

}  // namespace

namespace {
// measured per 4.2 M points, n_src = 2: scalar FFMA kernel 8.45 ms (shared memory, 168 registers + spills, 12 warps/SM) /
// 6.65 ms (constant bank); outputs as FFMA2 pairs + branch-free ELU 6.28 / 5.10 ms (profiles/r02_k10_variants.txt)
int g_blend_const = 1;
}
// Tuning knob: 1 = weights from the constant bank (shipped), 0 = from shared memory.
extern "C" int gens_debug_blend_const(int on) {
    g_blend_const = on ? 1 : 0;
    return 0;
}

extern "C" int gens_blend_weight_floats(void) { return kWeightFloats; }

extern "C" int gens_blend_colour(const float* rgb_feat, const float* ray_diff, const uint8_t* mask, long long n,
                                 int n_src, const float* weights, float* rgb_out, void* stream) {
    if (n == 0) return 0;
    GENS_CHECK_ARG(rgb_feat && ray_diff && mask && weights && rgb_out && n > 0 && n_src > 0);
    if (n_src > kMaxSrc) return GENS_E_UNSUPPORTED;
    if (g_blend_const) {
        constexpr int kT = 256;  // 128 registers without the shared-memory weight loads: two blocks of 256
        const int smem = 64 * kT * (int)sizeof(float);
        auto kern = blend_kernel<true, kT>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        e = cudaMemcpyToSymbolAsync(c_blend, weights, sizeof(float) * kWeightFloats, 0, cudaMemcpyDeviceToDevice,
                                    (cudaStream_t)stream);
        if (e != cudaSuccess) return (int)e;
        kern<<<ceil_div_i(n, kT), kT, smem, (cudaStream_t)stream>>>(rgb_feat, ray_diff, mask, n, n_src, weights, rgb_out);
        return gens_launch_status();
    }
    const int smem = (kWeightFloats + 64 * kThreads) * (int)sizeof(float);
    const cudaError_t e = cudaFuncSetAttribute(blend_kernel<false, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    blend_kernel<false, kThreads><<<ceil_div_i(n, kThreads), kThreads, smem, (cudaStream_t)stream>>>(rgb_feat, ray_diff, mask, n,
                                                                                          n_src, weights, rgb_out);
    return gens_launch_status();
}
