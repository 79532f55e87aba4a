// K5: hierarchical up-sampling along rays, one warp per ray (sm_100a).
//
// Replaces one iteration of the loop in ImplicitSurface.render (reference models/modules/
// implicit_surface.py:378-393), i.e. up_sample (:60-109) + sample_pdf (:14-44) and the sort/gather half
// of cat_z_vals (:111-133): ~60 ATen launches over (B,M) temporaries per iteration in the reference.
//   upsample_kernel  nearest-mask test of the current samples, NeuS section alphas at a fixed inv_s,
//                    transmittance as a warp-shuffle prefix product, inverse-CDF sampling of 16 new depths
//   merge_kernel     stable merge of the (sorted) new depths into the sorted ray, SDF values following
// Nothing is materialised between the stages of a ray: the ray's samples live in shared memory
// (<= 128 + 16 floats per array), each lane owns four consecutive sections.
#include "common.cuh"

namespace {

constexpr int kMaxSamples = 160;     // 64 coarse + 4*16 importance (+ slack)
constexpr int kWarpsPerBlock = 8;

struct MaskPyr {
    const float* vol[GENS_MAX_SCALES];
    int dim[GENS_MAX_SCALES];
    int n;
};

__device__ __forceinline__ float unnorm_nearest(float c, int size, int fused) {
    const float t = __fadd_rn(c, 1.0f), s = (float)size;
    const float u = fused ? __fmaf_rn(t, s, -1.0f) : __fsub_rn(__fmul_rn(t, s), 1.0f);
    return __fmul_rn(u, 0.5f);
}

// the `.any()` over scales of the nearest mask look-up, bit-identical to K2 (sampling.cu)
__device__ __forceinline__ bool mask_any(const MaskPyr& m, float p0, float p1, float p2, int fused) {
    for (int s = 0; s < m.n; ++s) {
        const int D = m.dim[s];
        const float a = nearbyintf(unnorm_nearest(p0, D, fused));
        const float b = nearbyintf(unnorm_nearest(p1, D, fused));
        const float c = nearbyintf(unnorm_nearest(p2, D, fused));
        if (a >= 0.f && a < (float)D && b >= 0.f && b < (float)D && c >= 0.f && c < (float)D &&
            __ldg(m.vol[s] + ((long long)a * D + (long long)b) * D + (long long)c) != 0.f)
            return true;
    }
    return false;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_excl_scan_mul(float v, int lane) {
    // inclusive product scan, then shift by one lane
    float x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x *= y;
    }
    const float prev = __shfl_up_sync(0xffffffffu, x, 1);
    return lane == 0 ? 1.0f : prev;
}

__device__ __forceinline__ float warp_excl_scan_add(float v, int lane, float& total) {
    float x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    total = __shfl_sync(0xffffffffu, x, 31);
    const float prev = __shfl_up_sync(0xffffffffu, x, 1);
    return lane == 0 ? 0.0f : prev;
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock)
upsample_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ z_vals,
                const float* __restrict__ sdf, int B, int M, MaskPyr masks, int fused, float inv_s, int n_new,
                float* __restrict__ new_z) {
    __shared__ float s_z[kWarpsPerBlock][kMaxSamples], s_sdf[kWarpsPerBlock][kMaxSamples];
    __shared__ float s_rad[kWarpsPerBlock][kMaxSamples], s_cdf[kWarpsPerBlock][kMaxSamples];
    __shared__ unsigned char s_valid[kWarpsPerBlock][kMaxSamples];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ray = blockIdx.x * kWarpsPerBlock + warp;
    if (ray >= B) return;
    float *z = s_z[warp], *sd = s_sdf[warp], *rad = s_rad[warp], *cdf = s_cdf[warp];
    unsigned char* valid = s_valid[warp];
    const float ox = rays_o[3 * ray], oy = rays_o[3 * ray + 1], oz = rays_o[3 * ray + 2];
    const float dx = rays_d[3 * ray], dy = rays_d[3 * ray + 1], dz = rays_d[3 * ray + 2];

    // stage the ray: depths, SDF, per-sample mask validity and radius
    for (int j = lane; j < M; j += 32) {
        const float zj = z_vals[(long long)ray * M + j];
        // pts = o + d * z, product and sum rounded separately as torch does (decides the nearest voxel)
        const float px = __fadd_rn(ox, __fmul_rn(dx, zj)), py = __fadd_rn(oy, __fmul_rn(dy, zj));
        const float pz = __fadd_rn(oz, __fmul_rn(dz, zj));
        z[j] = zj;
        sd[j] = sdf[(long long)ray * M + j];
        valid[j] = mask_any(masks, px, py, pz, fused) ? 1 : 0;
        rad[j] = sqrtf(px * px + py * py + pz * pz);
    }
    __syncwarp();

    // sections j = 4*lane .. 4*lane+3 (M-1 sections): alpha, then weights via a prefix product
    const int S = M - 1;
    float alpha[4], one_minus[4];
    float local_prod = 1.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = 4 * lane + k;
        float a = 0.0f;
        if (j < S) {
            const float s0 = sd[j], s1 = sd[j + 1], z0 = z[j], z1 = z[j + 1];
            const bool inside = (rad[j] < 1.0f || rad[j + 1] < 1.0f) && valid[j] && valid[j + 1];
            float c = (s1 - s0) / (z1 - z0 + 1e-5f);
            const float cp = j > 0 ? (s0 - sd[j - 1]) / (z0 - z[j - 1] + 1e-5f) : 0.0f;
            c = fminf(fmaxf(fminf(cp, c), -1e3f), 0.0f);
            c = inside ? c : 0.0f;
            const float mid = (s0 + s1) * 0.5f, dist = z1 - z0;
            const float prev_cdf = sigmoidf_((mid - c * dist * 0.5f) * inv_s);
            const float next_cdf = sigmoidf_((mid + c * dist * 0.5f) * inv_s);
            a = (prev_cdf - next_cdf + 1e-5f) / (prev_cdf + 1e-5f);
        }
        alpha[k] = a;
        one_minus[k] = j < S ? (1.0f - a + 1e-7f) : 1.0f;
        local_prod *= one_minus[k];
    }
    float T = warp_excl_scan_mul(local_prod, lane);  // transmittance entering this lane's first section
    float w[4], local_sum = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = 4 * lane + k;
        w[k] = j < S ? alpha[k] * T + 1e-5f : 0.0f;  // sample_pdf: weights + 1e-5
        T *= one_minus[k];
        local_sum += w[k];
    }
    float total;
    float run = warp_excl_scan_add(local_sum, lane, total);
    run /= total;  // cdf entering this lane's first section
    // cdf[0] = 0, cdf[j+1] = sum_{i<=j} w_i / total
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = 4 * lane + k;
        if (j < S) {
            run += w[k] / total;
            cdf[j + 1] = run;
        }
    }
    if (lane == 0) cdf[0] = 0.0f;
    __syncwarp();

    // inverse CDF at u_k = (k + 0.5) / n_new (torch.linspace(0.5/n, 1-0.5/n, n) for n = 16 is exact)
    for (int k = lane; k < n_new; k += 32) {
        const float u = ((float)k + 0.5f) / (float)n_new;
        int lo = 0, hi = M;  // first index with cdf[idx] > u  (searchsorted right=True)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
        }
        const int below = max(lo - 1, 0), above = min(lo, M - 1);
        float den = cdf[above] - cdf[below];
        den = den < 1e-5f ? 1.0f : den;
        const float t = (u - cdf[below]) / den;
        new_z[(long long)ray * n_new + k] = z[below] + t * (z[above] - z[below]);
    }
}

// merge sorted z (B,M) with sorted new_z (B,K) -> (B,M+K); sdf follows when given
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
merge_kernel(const float* __restrict__ z_vals, const float* __restrict__ sdf, const float* __restrict__ new_z,
             const float* __restrict__ new_sdf, int B, int M, int K, float* __restrict__ z_out,
             float* __restrict__ sdf_out) {
    __shared__ float s_z[kWarpsPerBlock][kMaxSamples], s_n[kWarpsPerBlock][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ray = blockIdx.x * kWarpsPerBlock + warp;
    if (ray >= B) return;
    float *z = s_z[warp], *nz = s_n[warp];
    for (int j = lane; j < M; j += 32) z[j] = z_vals[(long long)ray * M + j];
    for (int k = lane; k < K; k += 32) nz[k] = new_z[(long long)ray * K + k];
    __syncwarp();
    const long long ob = (long long)ray * (M + K);
    for (int j = lane; j < M; j += 32) {  // old element j lands after the new ones strictly smaller
        const float v = z[j];
        int lo = 0, hi = K;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (nz[mid] < v) lo = mid + 1; else hi = mid; }
        z_out[ob + j + lo] = v;
        if (sdf_out) sdf_out[ob + j + lo] = sdf[(long long)ray * M + j];
    }
    for (int k = lane; k < K; k += 32) {  // new element k lands after the old ones <= it
        const float v = nz[k];
        int lo = 0, hi = M;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (z[mid] <= v) lo = mid + 1; else hi = mid; }
        z_out[ob + k + lo] = v;
        if (sdf_out) sdf_out[ob + k + lo] = new_sdf[(long long)ray * K + k];
    }
}

}  // namespace

extern "C" int gens_upsample_rays(const float* rays_o, const float* rays_d, const float* z_vals, const float* sdf,
                                  int n_rays, int n_samples, const gens_pyramid_t* masks, int aten_cuda_flavour,
                                  float inv_s, int n_new, float* new_z, void* stream) {
    if (n_rays == 0) return 0;
    GENS_CHECK_ARG(rays_o && rays_d && z_vals && sdf && masks && new_z && n_rays > 0 && n_samples > 1 && n_new > 0);
    if (n_samples > 128 || n_new > 32 || masks->n_scales <= 0 || masks->n_scales > GENS_MAX_SCALES)
        return GENS_E_UNSUPPORTED;
    MaskPyr m;
    m.n = masks->n_scales;
    for (int s = 0; s < m.n; ++s) {
        GENS_CHECK_ARG(masks->vol[s] && masks->dim[s] > 0);
        m.vol[s] = masks->vol[s];
        m.dim[s] = masks->dim[s];
    }
    upsample_kernel<<<ceil_div_i(n_rays, kWarpsPerBlock), 32 * kWarpsPerBlock, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, z_vals, sdf, n_rays, n_samples, m, aten_cuda_flavour, inv_s, n_new, new_z);
    return gens_launch_status();
}

extern "C" int gens_merge_samples(const float* z_vals, const float* sdf, const float* new_z, const float* new_sdf,
                                  int n_rays, int n_samples, int n_new, float* z_out, float* sdf_out, void* stream) {
    if (n_rays == 0) return 0;
    GENS_CHECK_ARG(z_vals && new_z && z_out && n_rays > 0 && n_samples > 0 && n_new > 0);
    GENS_CHECK_ARG((sdf_out == nullptr) || (sdf && new_sdf));
    if (n_samples > kMaxSamples || n_new > 32) return GENS_E_UNSUPPORTED;
    merge_kernel<<<ceil_div_i(n_rays, kWarpsPerBlock), 32 * kWarpsPerBlock, 0, (cudaStream_t)stream>>>(
        z_vals, sdf, new_z, new_sdf, n_rays, n_samples, n_new, z_out, sdf_out);
    return gens_launch_status();
}
