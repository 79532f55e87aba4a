// K6: source-view reprojection sampling for the colour branch of the ray marcher (sm_100a).
//
// Replaces projector.lookup_feature + compute_angle (reference models/modules/projector.py:278-349):
// per sample point and source view, project into the view at each of the S pyramid scales, test
// visibility (z>0, 0<=x<W_i, 0<=y<H_i, AND over scales), bilinearly sample the RGB image (scale 0) and the
// 4-channel feature map of every scale (align_corners=False, zeros padding) and compute the IBRNet
// ray-direction-difference features.  The reference issues 6 grid_sampler launches, 10 matmuls and a
// dozen permute/cat copies per call; here it is one launch, one thread per (point, view).
//
// Bit-exact contract: the visibility mask.  The projection follows the reference's two matmuls as
// k-ascending fma chains (what cuBLAS / MKL do for K = 4 and K = 3), IEEE division, no epsilon.
#include "common.cuh"

namespace {

struct SrcPyr {
    const float4* feat[GENS_MAX_SCALES];  // channels-last (ns, H_i, W_i, 4), source views only
    int h[GENS_MAX_SCALES], w[GENS_MAX_SCALES];
    int n;
};

struct GradSrcPyr {
    float4* feat[GENS_MAX_SCALES];
};

struct Bilin {
    int x0, y0;
    float w00, w01, w10, w11;  // (y0,x0) (y0,x1) (y1,x0) (y1,x1)
};

// grid_sample bilinear footprint, align_corners=False: ix = ((n+1)*size-1)/2
__device__ __forceinline__ Bilin footprint_unaligned(float nx, float ny, int W, int H, int fused) {
    const float tx = __fadd_rn(nx, 1.0f), ty = __fadd_rn(ny, 1.0f);
    float ix = fused ? __fmaf_rn(tx, (float)W, -1.0f) : __fsub_rn(__fmul_rn(tx, (float)W), 1.0f);
    float iy = fused ? __fmaf_rn(ty, (float)H, -1.0f) : __fsub_rn(__fmul_rn(ty, (float)H), 1.0f);
    ix *= 0.5f;
    iy *= 0.5f;
    // far-away / non-finite coordinates: everything is padding
    if (!(ix > -2.0f && ix < (float)W + 1.0f)) ix = -2.0f;
    if (!(iy > -2.0f && iy < (float)H + 1.0f)) iy = -2.0f;
    const float fx = floorf(ix), fy = floorf(iy);
    const float bx = ix - fx, by = iy - fy, ax = 1.0f - bx, ay = 1.0f - by;
    Bilin b;
    b.x0 = (int)fx; b.y0 = (int)fy;
    b.w00 = ax * ay; b.w01 = bx * ay; b.w10 = ax * by; b.w11 = bx * by;
    return b;
}

__device__ __forceinline__ float4 sample_zeros(const float4* __restrict__ map, int H, int W, const Bilin& b) {
    const bool x0 = (unsigned)b.x0 < (unsigned)W, x1 = (unsigned)(b.x0 + 1) < (unsigned)W;
    const bool y0 = (unsigned)b.y0 < (unsigned)H, y1 = (unsigned)(b.y0 + 1) < (unsigned)H;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long base = (long long)b.y0 * W + b.x0;
    const float4 v00 = (x0 && y0) ? __ldg(map + base) : z, v01 = (x1 && y0) ? __ldg(map + base + 1) : z;
    const float4 v10 = (x0 && y1) ? __ldg(map + base + W) : z, v11 = (x1 && y1) ? __ldg(map + base + W + 1) : z;
    float4 r;
    r.x = fmaf(v11.x, b.w11, fmaf(v10.x, b.w10, fmaf(v01.x, b.w01, v00.x * b.w00)));
    r.y = fmaf(v11.y, b.w11, fmaf(v10.y, b.w10, fmaf(v01.y, b.w01, v00.y * b.w00)));
    r.z = fmaf(v11.z, b.w11, fmaf(v10.z, b.w10, fmaf(v01.z, b.w01, v00.z * b.w00)));
    r.w = fmaf(v11.w, b.w11, fmaf(v10.w, b.w10, fmaf(v01.w, b.w01, v00.w * b.w00)));
    return r;
}

__device__ __forceinline__ void scatter_zeros(float4* __restrict__ gmap, int H, int W, const Bilin& b, const float4 g) {
    const bool x0 = (unsigned)b.x0 < (unsigned)W, x1 = (unsigned)(b.x0 + 1) < (unsigned)W;
    const bool y0 = (unsigned)b.y0 < (unsigned)H, y1 = (unsigned)(b.y0 + 1) < (unsigned)H;
    const long long base = (long long)b.y0 * W + b.x0;
    auto add = [&](long long i, float w) { atomicAdd(gmap + i, make_float4(g.x * w, g.y * w, g.z * w, g.w * w)); };
    if (x0 && y0) add(base, b.w00);
    if (x1 && y0) add(base + 1, b.w01);
    if (x0 && y1) add(base + W, b.w10);
    if (x1 && y1) add(base + W + 1, b.w11);
}

struct ViewProj {
    float q0[GENS_MAX_SCALES], q1[GENS_MAX_SCALES], q2;  // per-scale image-plane numerators, shared depth
};

// cam = (w2c @ [p,1])[:3] ; q_i = K_i[:3,:3] @ cam with rows 0-1 of K scaled by 0.5^i (exact)
__device__ __forceinline__ void project_src(const float* __restrict__ w2c, const float* __restrict__ K, float px,
                                            float py, float pz, int n_scales, ViewProj& o) {
    float cam[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) cam[r] = row_dot4(w2c + 4 * r, px, py, pz, 1.0f);
    float s = 1.0f;
    for (int i = 0; i < n_scales; ++i) {
        float t = __fmul_rn(__fmul_rn(K[0], s), cam[0]);
        t = __fmaf_rn(__fmul_rn(K[1], s), cam[1], t);
        o.q0[i] = __fmaf_rn(__fmul_rn(K[2], s), cam[2], t);
        t = __fmul_rn(__fmul_rn(K[4], s), cam[0]);
        t = __fmaf_rn(__fmul_rn(K[5], s), cam[1], t);
        o.q1[i] = __fmaf_rn(__fmul_rn(K[6], s), cam[2], t);
        s *= 0.5f;
    }
    float t = __fmul_rn(K[8], cam[0]);
    t = __fmaf_rn(K[9], cam[1], t);
    o.q2 = __fmaf_rn(K[10], cam[2], t);
}

__device__ __forceinline__ float div_scalar_flavour(float a, float b, int recip) {
    return recip ? __fmul_rn(a, __fdiv_rn(1.0f, b)) : __fdiv_rn(a, b);
}

template <bool BACKWARD>
__global__ void __launch_bounds__(256)
lookup_feature_kernel(const float* __restrict__ pts, long long n, int ns, const float* __restrict__ w2c_src,
                      const float* __restrict__ k_src, const float* __restrict__ c2w_ref, const float* __restrict__ c2w_src,
                      SrcPyr pyr, const float4* __restrict__ rgb, int flavour, float* __restrict__ feat_out,
                      float* __restrict__ raydiff_out, uint8_t* __restrict__ mask_out, const float* __restrict__ g_feat,
                      GradSrcPyr gpyr) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * ns) return;
    const long long p = i / ns;
    const int v = (int)(i % ns);
    const float px = __ldg(pts + 3 * p), py = __ldg(pts + 3 * p + 1), pz = __ldg(pts + 3 * p + 2);
    ViewProj pr;
    project_src(w2c_src + 16 * v, k_src + 16 * v, px, py, pz, pyr.n, pr);
    const int width = 3 + 4 * pyr.n;
    bool vis = pr.q2 > 0.0f;
    for (int s = 0; s < pyr.n; ++s) {
        const int H = pyr.h[s], W = pyr.w[s];
        const float x = __fdiv_rn(pr.q0[s], pr.q2), y = __fdiv_rn(pr.q1[s], pr.q2);
        vis = vis && (x >= 0.0f) && (x < (float)W) && (y >= 0.0f) && (y < (float)H);
        const float hx = (float)((double)(W - 1) / 2.0), hy = (float)((double)(H - 1) / 2.0);
        const float nx = __fsub_rn(div_scalar_flavour(x, hx, flavour), 1.0f);
        const float ny = __fsub_rn(div_scalar_flavour(y, hy, flavour), 1.0f);
        const Bilin b = footprint_unaligned(nx, ny, W, H, flavour);
        const long long vo = (long long)v * H * W;
        if (!BACKWARD) {
            const float4 f = sample_zeros(pyr.feat[s] + vo, H, W, b);
            float* o = feat_out + i * width + 3 + 4 * s;
            o[0] = f.x; o[1] = f.y; o[2] = f.z; o[3] = f.w;
            if (s == 0) {
                const float4 c = sample_zeros(rgb + vo, H, W, b);
                float* oc = feat_out + i * width;
                oc[0] = c.x; oc[1] = c.y; oc[2] = c.z;
            }
        } else if (gpyr.feat[s]) {
            const float* g = g_feat + i * width + 3 + 4 * s;
            scatter_zeros(gpyr.feat[s] + vo, H, W, b, make_float4(g[0], g[1], g[2], g[3]));
        }
    }
    if (BACKWARD) return;
    mask_out[i] = vis ? 1 : 0;
    // compute_angle (projector.py:278-291)
    float a[3], bvec[3], d[3];
    const float pp[3] = {px, py, pz};
    float na = 0.f, nb = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a[k] = __ldg(c2w_ref + 4 * k + 3) - pp[k];
        bvec[k] = __ldg(c2w_src + 16 * v + 4 * k + 3) - pp[k];
        na += a[k] * a[k];
        nb += bvec[k] * bvec[k];
    }
    const float ia = 1.0f / (sqrtf(na) + 1e-6f), ib = 1.0f / (sqrtf(nb) + 1e-6f);
    float nd = 0.f, dot = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a[k] *= ia;
        bvec[k] *= ib;
        d[k] = a[k] - bvec[k];
        nd += d[k] * d[k];
        dot += a[k] * bvec[k];
    }
    const float idn = 1.0f / fmaxf(sqrtf(nd), 1e-6f);
    float* r = raydiff_out + i * 4;
    r[0] = d[0] * idn; r[1] = d[1] * idn; r[2] = d[2] * idn; r[3] = dot;
}

// (n,C,h,w) NCHW, C in {3,4} -> channels-last (n,h,w,4) (unused channel = 0), and the inverse for gradients
__global__ void __launch_bounds__(256)
pack_nhwc4_kernel(const float* __restrict__ src, float4* __restrict__ dst, int C, long long hw, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long nidx = i / hw, p = i % hw;
    const float* s = src + nidx * C * hw + p;
    dst[i] = make_float4(__ldg(s), __ldg(s + hw), __ldg(s + 2 * hw), C > 3 ? __ldg(s + 3 * hw) : 0.f);
}
__global__ void __launch_bounds__(256)
unpack_nhwc4_kernel(const float4* __restrict__ src, float* __restrict__ dst, int C, long long hw, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long nidx = i / hw, p = i % hw;
    const float4 v = __ldg(src + i);
    float* d = dst + nidx * C * hw + p;
    d[0] = v.x; d[hw] = v.y; d[2 * hw] = v.z;
    if (C > 3) d[3 * hw] = v.w;
}

bool fill_src(const gens_image_pyramid_t* p, SrcPyr& o) {
    if (!p || p->n_scales <= 0 || p->n_scales > GENS_MAX_SCALES) return false;
    o.n = p->n_scales;
    for (int s = 0; s < o.n; ++s) {
        if (!p->map[s] || p->h[s] <= 0 || p->w[s] <= 0) return false;
        o.feat[s] = reinterpret_cast<const float4*>(p->map[s]);
        o.h[s] = p->h[s];
        o.w[s] = p->w[s];
    }
    return true;
}

}  // namespace

extern "C" int gens_pack_nhwc4(const float* src_nchw, float* dst_nhwc4, int n, int c, int h, int w, void* stream) {
    GENS_CHECK_ARG(src_nchw && dst_nhwc4 && n > 0 && h > 0 && w > 0);
    if (c != 3 && c != 4) return GENS_E_UNSUPPORTED;
    const long long hw = (long long)h * w, total = hw * n;
    pack_nhwc4_kernel<<<ceil_div_i(total, 256), 256, 0, (cudaStream_t)stream>>>(src_nchw, reinterpret_cast<float4*>(dst_nhwc4),
                                                                               c, hw, total);
    return gens_launch_status();
}

extern "C" int gens_unpack_nhwc4(const float* src_nhwc4, float* dst_nchw, int n, int c, int h, int w, void* stream) {
    GENS_CHECK_ARG(src_nhwc4 && dst_nchw && n > 0 && h > 0 && w > 0);
    if (c != 3 && c != 4) return GENS_E_UNSUPPORTED;
    const long long hw = (long long)h * w, total = hw * n;
    unpack_nhwc4_kernel<<<ceil_div_i(total, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(src_nhwc4),
                                                                                 dst_nchw, c, hw, total);
    return gens_launch_status();
}

extern "C" int gens_lookup_feature_fwd(const float* pts, long long n, int n_src, const float* w2c_src, const float* k_src,
                                       const float* c2w_ref, const float* c2w_src, const gens_image_pyramid_t* feats,
                                       const float* rgb_nhwc4, int aten_cuda_flavour, float* feat_out, float* raydiff_out,
                                       uint8_t* mask_out, void* stream) {
    if (n == 0) return 0;
    GENS_CHECK_ARG(pts && w2c_src && k_src && c2w_ref && c2w_src && rgb_nhwc4 && feat_out && raydiff_out && mask_out);
    GENS_CHECK_ARG(n > 0 && n_src > 0);
    SrcPyr p;
    if (!fill_src(feats, p)) return GENS_E_BADARG;
    GradSrcPyr g = {};
    lookup_feature_kernel<false><<<ceil_div_i(n * n_src, 256), 256, 0, (cudaStream_t)stream>>>(
        pts, n, n_src, w2c_src, k_src, c2w_ref, c2w_src, p, reinterpret_cast<const float4*>(rgb_nhwc4), aten_cuda_flavour,
        feat_out, raydiff_out, mask_out, nullptr, g);
    return gens_launch_status();
}

extern "C" int gens_lookup_feature_bwd(const float* pts, long long n, int n_src, const float* w2c_src, const float* k_src,
                                       const gens_image_pyramid_t* feats, int aten_cuda_flavour, const float* g_feat,
                                       const gens_image_pyramid_t* g_feats, void* stream) {
    if (n == 0) return 0;
    GENS_CHECK_ARG(pts && w2c_src && k_src && g_feat && g_feats && n > 0 && n_src > 0);
    SrcPyr p;
    if (!fill_src(feats, p)) return GENS_E_BADARG;
    GradSrcPyr g = {};
    for (int s = 0; s < p.n; ++s) g.feat[s] = reinterpret_cast<float4*>(const_cast<float*>(g_feats->map[s]));
    lookup_feature_kernel<true><<<ceil_div_i(n * n_src, 256), 256, 0, (cudaStream_t)stream>>>(
        pts, n, n_src, w2c_src, k_src, nullptr, nullptr, p, nullptr, aten_cuda_flavour, nullptr, nullptr, nullptr, g_feat, g);
    return gens_launch_status();
}
