// K11: patch normalised-cross-correlation score of the feature-metric consistency term (sm_100a).
//
// Replaces compute_LNCC (reference models/losses/ncc.py:7-50), the consumer of K8's patches
// (loss.py:36-38).  The reference permutes both patch tensors, builds three more of the same size
// (products and squares) and runs FIVE grouped patch x patch convolutions over zero-padded patches only
// to read the centre pixel -- i.e. it computes plain sums over the P = patch^2 samples -- followed by
// ~30 element-wise ops.  Here one block owns one ray: thread (s, c) accumulates the five sums of source
// view s / channel c in registers straight from K8's output layout (n,P,C) / (S,n,P,C) (consecutive
// threads read consecutive channels: coalesced), then the block reduces over channels, picks the two
// best source views and writes one score.  The backward kernel re-reads the patches once and writes both
// patch gradients.
//
// Arithmetic follows ncc.py:34-49 term by term (same association, fp32):
//   u = sum / P;  cross = rs - u_s r - u_r s + u_r u_s P;  var = sq - 2 u sum + u u P
//   cc = cross^2 / (var_r var_s + 1e-5);  ncc_c = clamp(1 - cc, 0, 2);  ncc_s = mean_c ncc_c
//   score = mean of the two smallest ncc_s                (torch.topk(k=2, largest=False))
// Sums over the P samples are sequential per (s, c) (the convolution's accumulation order is cuDNN's and
// not reproducible anyway: compared at 1e-5 in the tests).
#include "common.cuh"

namespace {

constexpr int kMaxSrc = 15;   // source views
constexpr int kMaxC = 16;     // channels of the fused feature image (12 in GenS)

struct Sums {
    float r, s, rr, ss, rs;
};

__device__ __forceinline__ void ncc_terms(const Sums& a, float P, float& cross, float& rvar, float& svar) {
    const float ur = a.r / P, us = a.s / P;
    cross = a.rs - us * a.r - ur * a.s + ur * us * P;
    rvar = a.rr - 2.0f * ur * a.r + ur * ur * P;
    svar = a.ss - 2.0f * us * a.s + us * us * P;
}

// ref (n,P,C), src (S,n,P,C) -> score (n); optional per-(ray, view) ncc (n,S) and the two selected views (n,2)
__global__ void __launch_bounds__(256)
lncc_fwd_kernel(const float* __restrict__ ref, const float* __restrict__ src, int n, int S, int P, int C,
                float* __restrict__ score, float* __restrict__ ncc_view, int* __restrict__ picked) {
    __shared__ float s_ncc[kMaxSrc * kMaxC];
    __shared__ float s_view[kMaxSrc];
    const int b = blockIdx.x, t = threadIdx.x;
    const int s = t / C, c = t % C;
    if (t < S * C) {
        const float* r = ref + (long long)b * P * C + c;
        const float* q = src + ((long long)s * n + b) * P * C + c;
        Sums a = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int p = 0; p < P; ++p) {
            const float x = __ldg(r + p * C), y = __ldg(q + p * C);
            a.r += x;
            a.s += y;
            a.rr += x * x;
            a.ss += y * y;
            a.rs += x * y;
        }
        float cross, rvar, svar;
        ncc_terms(a, (float)P, cross, rvar, svar);
        const float cc = cross * cross / (rvar * svar + 1e-5f);
        s_ncc[t] = fminf(fmaxf(1.0f - cc, 0.0f), 2.0f);
    }
    __syncthreads();
    if (t < S) {
        float m = 0.f;
        for (int k = 0; k < C; ++k) m += s_ncc[t * C + k];
        m /= (float)C;
        s_view[t] = m;
        if (ncc_view) ncc_view[(long long)b * S + t] = m;
    }
    __syncthreads();
    if (t == 0) {
        // two smallest, ties resolved towards the lower index like a stable selection
        int i0 = 0;
        for (int k = 1; k < S; ++k)
            if (s_view[k] < s_view[i0]) i0 = k;
        int i1 = -1;
        for (int k = 0; k < S; ++k)
            if (k != i0 && (i1 < 0 || s_view[k] < s_view[i1])) i1 = k;
        const float v = i1 >= 0 ? (s_view[i0] + s_view[i1]) / 2.0f : s_view[i0];
        score[b] = v;
        if (picked) {
            picked[2 * b] = i0;
            picked[2 * b + 1] = i1;
        }
    }
}

// d score / d patches.  g_score (n); picked (n,2) from the forward.  g_ref (n,P,C) and g_src (S,n,P,C) are fully
// written (zeros for the views that were not selected).
__global__ void __launch_bounds__(256)
lncc_bwd_kernel(const float* __restrict__ ref, const float* __restrict__ src, const float* __restrict__ g_score,
                const int* __restrict__ picked, int n, int S, int P, int C, float* __restrict__ g_ref,
                float* __restrict__ g_src) {
    // per (s, c): coefficients of  d ncc_c / d r_p = A (s_p - u_s) + B (r_p - u_r),  d / d s_p = A (r_p - u_r) + D (s_p - u_s)
    __shared__ float s_A[kMaxSrc * kMaxC], s_B[kMaxSrc * kMaxC], s_D[kMaxSrc * kMaxC], s_ur[kMaxSrc * kMaxC],
        s_us[kMaxSrc * kMaxC];
    const int b = blockIdx.x, t = threadIdx.x;
    const int s = t / C, c = t % C;
    const int i0 = picked[2 * b], i1 = picked[2 * b + 1];
    const float g = g_score[b];
    if (t < S * C) {
        const bool sel = s == i0 || s == i1;
        float A = 0.f, B = 0.f, D = 0.f, ur = 0.f, us = 0.f;
        if (sel) {
            const float* r = ref + (long long)b * P * C + c;
            const float* q = src + ((long long)s * n + b) * P * C + c;
            Sums a = {0.f, 0.f, 0.f, 0.f, 0.f};
            for (int p = 0; p < P; ++p) {
                const float x = __ldg(r + p * C), y = __ldg(q + p * C);
                a.r += x; a.s += y; a.rr += x * x; a.ss += y * y; a.rs += x * y;
            }
            float cross, rvar, svar;
            ncc_terms(a, (float)P, cross, rvar, svar);
            ur = a.r / (float)P;
            us = a.s / (float)P;
            const float den = rvar * svar + 1e-5f;
            const float cc = cross * cross / den;
            const float one_m = 1.0f - cc;
            if (one_m > 0.0f && one_m < 2.0f) {  // inside the clamp
                // weight of this (view, channel) in the score: 1/C (channel mean) x 1/2 (two views; 1 if S == 1)
                const float w = -g / (float)C * (i1 >= 0 ? 0.5f : 1.0f);  // d score / d cc
                const float dcross = 2.0f * cross / den, dvar = -cross * cross / (den * den);
                A = w * dcross;
                B = w * dvar * svar * 2.0f;  // d cc / d rvar * d rvar / d r_p = dvar * svar * 2 (r_p - u_r)
                D = w * dvar * rvar * 2.0f;
            }
        }
        s_A[t] = A; s_B[t] = B; s_D[t] = D; s_ur[t] = ur; s_us[t] = us;
    }
    __syncthreads();
    // gradient w.r.t. the source patches: one element per thread-iteration, (s, p, c) with c fastest
    const int per_view = P * C;
    for (int i = t; i < S * per_view; i += blockDim.x) {
        const int sv = i / per_view, rem = i % per_view, cc_ = rem % C;
        const int k = sv * C + cc_;
        const long long o = ((long long)sv * n + b) * per_view + rem;
        float gv = 0.f;
        if (s_A[k] != 0.f || s_D[k] != 0.f) {
            const float x = __ldg(ref + (long long)b * per_view + rem), y = __ldg(src + o);
            gv = s_A[k] * (x - s_ur[k]) + s_D[k] * (y - s_us[k]);
        }
        g_src[o] = gv;
    }
    // gradient w.r.t. the reference patch: sum over the selected views
    for (int i = t; i < per_view; i += blockDim.x) {
        const int cc_ = i % C;
        const float x = __ldg(ref + (long long)b * per_view + i);
        float gv = 0.f;
        for (int sv = 0; sv < S; ++sv) {
            const int k = sv * C + cc_;
            if (s_A[k] != 0.f || s_B[k] != 0.f) {
                const float y = __ldg(src + ((long long)sv * n + b) * per_view + i);
                gv += s_A[k] * (y - s_us[k]) + s_B[k] * (x - s_ur[k]);
            }
        }
        g_ref[(long long)b * per_view + i] = gv;
    }
}

}  // namespace

extern "C" int gens_lncc_fwd(const float* ref, const float* src, int n_rays, int n_src, int n_samples, int channels,
                             float* score, float* ncc_view, int* picked, void* stream) {
    if (n_rays == 0) return 0;
    GENS_CHECK_ARG(ref && src && score && n_rays > 0 && n_src > 0 && n_samples > 0 && channels > 0);
    if (n_src > kMaxSrc || channels > kMaxC || n_src * channels > 256) return GENS_E_UNSUPPORTED;
    lncc_fwd_kernel<<<n_rays, 256, 0, (cudaStream_t)stream>>>(ref, src, n_rays, n_src, n_samples, channels, score,
                                                             ncc_view, picked);
    return gens_launch_status();
}

extern "C" int gens_lncc_bwd(const float* ref, const float* src, const float* g_score, const int* picked, int n_rays,
                             int n_src, int n_samples, int channels, float* g_ref, float* g_src, void* stream) {
    if (n_rays == 0) return 0;
    GENS_CHECK_ARG(ref && src && g_score && picked && g_ref && g_src && n_rays > 0 && n_src > 0 && n_samples > 0 &&
                   channels > 0);
    if (n_src > kMaxSrc || channels > kMaxC || n_src * channels > 256) return GENS_E_UNSUPPORTED;
    lncc_bwd_kernel<<<n_rays, 256, 0, (cudaStream_t)stream>>>(ref, src, g_score, picked, n_rays, n_src, n_samples,
                                                             channels, g_ref, g_src);
    return gens_launch_status();
}
