// K2 / K3: multi-scale volume look-ups for the ray marcher (sm_100a).
//
// Replaces projector.lookup_volume (reference models/modules/projector.py:217-245) and the
// autograd triple behind it (cuda_gridsample.py:71-123 + the native gridsample_cuda.cu:212-533):
//   K2  mask_nearest     5 x F.grid_sample(mode='nearest', align_corners=False) + .any()
//   K3  trilinear_fwd    5 x grid_sample_3d(bilinear, zeros, align_corners=True) -> (n, 4*S)
//       trilinear_bwd    first-order backward: d/dpts (+ optional scatter into the volumes)
//       trilinear_bwd2   backward of the backward (what `grad2_3d` computes): mixed partials
// One launch covers every scale of the pyramid; the reference issues one ATen launch per scale
// plus reshape/permute/cat, 85 grid_sampler launches per render() call.
//
// Layout: feature volumes are re-packed once per volume version to channels-last (D,D,D,4) so a
// trilinear corner is one 16-byte load; a point's coordinate p = (p0,p1,p2) indexes tensor dims
// (2,3,4) directly -- the reference's pts.flip(-1) only exists to satisfy grid_sample's (x,y,z)
// = (W,H,D) convention.
#include "common.cuh"

namespace {

struct Pyr {
    const float4* vol[GENS_MAX_SCALES];  // channels-last (D,D,D,4)
    int dim[GENS_MAX_SCALES];
    int n;
};

struct MaskPyr {
    const float* vol[GENS_MAX_SCALES];  // (D,D,D) floats as the reference stores them
    int dim[GENS_MAX_SCALES];
    int n;
};

// ATen grid_sampler_unnormalize, align_corners=False: ((c+1)*size-1)/2.  ATen's CUDA build
// contracts the multiply-subtract into one fma; its CPU build does not.
__device__ __forceinline__ float unnorm_nearest(float c, int size, int fused) {
    const float t = __fadd_rn(c, 1.0f), s = (float)size;
    const float u = fused ? __fmaf_rn(t, s, -1.0f) : __fsub_rn(__fmul_rn(t, s), 1.0f);
    return __fmul_rn(u, 0.5f);
}

__global__ void __launch_bounds__(256)
mask_nearest_kernel(const float* __restrict__ pts, long long n, MaskPyr m, int fused, uint8_t* __restrict__ any_out,
                    float* __restrict__ each_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p0 = __ldg(pts + 3 * i), p1 = __ldg(pts + 3 * i + 1), p2 = __ldg(pts + 3 * i + 2);
    bool any = false;
#pragma unroll 1
    for (int s = 0; s < m.n; ++s) {
        const int D = m.dim[s];
        // nearbyint = round half to even, as ATen (std::nearbyint / ::nearbyintf)
        const float a = nearbyintf(unnorm_nearest(p0, D, fused));
        const float b = nearbyintf(unnorm_nearest(p1, D, fused));
        const float c = nearbyintf(unnorm_nearest(p2, D, fused));
        float v = 0.f;
        if (a >= 0.f && a < (float)D && b >= 0.f && b < (float)D && c >= 0.f && c < (float)D)
            v = __ldg(m.vol[s] + ((long long)a * D + (long long)b) * D + (long long)c);
        any |= (v != 0.f);
        if (each_out) each_out[i * m.n + s] = v;
    }
    if (any_out) any_out[i] = any ? 1 : 0;
}

// ---- trilinear ------------------------------------------------------------------------------
struct Cell {
    int i0[3];    // floor index per axis
    float t[3];   // fractional position in the cell
    float mult;   // d(unnormalised)/d(normalised) = (D-1)/2
};

__device__ __forceinline__ Cell locate(float p0, float p1, float p2, int D) {
    Cell c;
    const float p[3] = {p0, p1, p2};
    c.mult = 0.5f * (float)(D - 1);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        // ATen grid_sampler_unnormalize, align_corners=True: ((c+1)/2)*(size-1)
        float u = ((p[k] + 1.0f) * 0.5f) * (float)(D - 1);
        u = fminf(fmaxf(u, -2.0f), (float)D + 1.0f);  // far outside: every corner is padding anyway
        if (!(u == u)) u = -2.0f;
        const float f = floorf(u);
        c.i0[k] = (int)f;
        c.t[k] = u - f;
    }
    return c;
}

__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4lerp(const float4 a, const float4 b, float t) {
    return make_float4(fmaf(t, b.x - a.x, a.x), fmaf(t, b.y - a.y, a.y), fmaf(t, b.z - a.z, a.z), fmaf(t, b.w - a.w, a.w));
}
__device__ __forceinline__ float4 f4sub(const float4 a, const float4 b) {
    return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}
__device__ __forceinline__ float f4dot(const float4 a, const float4 b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float4 f4scale(const float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4axpy(float s, const float4 a, const float4 y) {
    return make_float4(fmaf(s, a.x, y.x), fmaf(s, a.y, y.y), fmaf(s, a.z, y.z), fmaf(s, a.w, y.w));
}

// the 8 corner values of the cell, zero where a corner lies outside the volume (zeros padding)
__device__ __forceinline__ void load_corners(const float4* __restrict__ vol, int D, const Cell& c, float4 (&v)[2][2][2]) {
#pragma unroll
    for (int da = 0; da < 2; ++da)
#pragma unroll
        for (int db = 0; db < 2; ++db)
#pragma unroll
            for (int dc = 0; dc < 2; ++dc) {
                const int a = c.i0[0] + da, b = c.i0[1] + db, cc = c.i0[2] + dc;
                const bool in = (unsigned)a < (unsigned)D && (unsigned)b < (unsigned)D && (unsigned)cc < (unsigned)D;
                v[da][db][dc] = in ? __ldg(vol + ((long long)a * D + b) * D + cc) : f4zero();
            }
}

// Value, first derivatives and mixed second derivatives of the trilinear interpolant w.r.t. the
// UNNORMALISED coordinates (multiply by `mult` per differentiation for normalised ones).
struct Jet {
    float4 f, d[3], dd[3];  // dd[0] = d2/da db, dd[1] = d2/da dc, dd[2] = d2/db dc
};

template <int ORDER>
__device__ __forceinline__ Jet interpolate(const float4 (&v)[2][2][2], const Cell& c) {
    const float ta = c.t[0], tb = c.t[1], tc = c.t[2];
    float4 l[2][2], lc[2][2];  // along axis 2 (fastest)
#pragma unroll
    for (int da = 0; da < 2; ++da)
#pragma unroll
        for (int db = 0; db < 2; ++db) {
            l[da][db] = f4lerp(v[da][db][0], v[da][db][1], tc);
            if (ORDER >= 1) lc[da][db] = f4sub(v[da][db][1], v[da][db][0]);
        }
    float4 m[2], mb[2], mc[2], mbc[2];  // along axis 1
#pragma unroll
    for (int da = 0; da < 2; ++da) {
        m[da] = f4lerp(l[da][0], l[da][1], tb);
        if (ORDER >= 1) {
            mb[da] = f4sub(l[da][1], l[da][0]);
            mc[da] = f4lerp(lc[da][0], lc[da][1], tb);
        }
        if (ORDER >= 2) mbc[da] = f4sub(lc[da][1], lc[da][0]);
    }
    Jet j;
    j.f = f4lerp(m[0], m[1], ta);
    if (ORDER >= 1) {
        j.d[0] = f4sub(m[1], m[0]);
        j.d[1] = f4lerp(mb[0], mb[1], ta);
        j.d[2] = f4lerp(mc[0], mc[1], ta);
    }
    if (ORDER >= 2) {
        j.dd[0] = f4sub(mb[1], mb[0]);
        j.dd[1] = f4sub(mc[1], mc[0]);
        j.dd[2] = f4lerp(mbc[0], mbc[1], ta);
    }
    return j;
}

__global__ void __launch_bounds__(256)
trilinear_fwd_kernel(const float* __restrict__ pts, long long n, Pyr pyr, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p0 = __ldg(pts + 3 * i), p1 = __ldg(pts + 3 * i + 1), p2 = __ldg(pts + 3 * i + 2);
    float4* o = reinterpret_cast<float4*>(out + i * 4 * pyr.n);
#pragma unroll 1
    for (int s = 0; s < pyr.n; ++s) {
        const Cell c = locate(p0, p1, p2, pyr.dim[s]);
        float4 v[2][2][2];
        load_corners(pyr.vol[s], pyr.dim[s], c, v);
        o[s] = interpolate<0>(v, c).f;
    }
}

// scatter  coef * w_corner  into the channels-last gradient volume, w = product of per-axis weights
__device__ __forceinline__ void scatter_corners(float4* __restrict__ gvol, int D, const Cell& c, const float (&wa)[2],
                                                const float (&wb)[2], const float (&wc)[2], const float4 g) {
#pragma unroll
    for (int da = 0; da < 2; ++da)
#pragma unroll
        for (int db = 0; db < 2; ++db)
#pragma unroll
            for (int dc = 0; dc < 2; ++dc) {
                const int a = c.i0[0] + da, b = c.i0[1] + db, cc = c.i0[2] + dc;
                if ((unsigned)a < (unsigned)D && (unsigned)b < (unsigned)D && (unsigned)cc < (unsigned)D) {
                    const float w = wa[da] * wb[db] * wc[dc];
                    if (w != 0.f) atomicAdd(gvol + ((long long)a * D + b) * D + cc, f4scale(g, w));
                }
            }
}

struct GradPyr {
    float4* vol[GENS_MAX_SCALES];  // channels-last gradient volumes (zero-initialised by the caller) or null
};

// first-order backward: g_pts[k] = sum_c g_out[c] * dfeat_c/dp_k ; g_vol += g_out * w
__global__ void __launch_bounds__(256)
trilinear_bwd_kernel(const float* __restrict__ pts, long long n, Pyr pyr, const float* __restrict__ g_out,
                     float* __restrict__ g_pts, GradPyr gv) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p0 = __ldg(pts + 3 * i), p1 = __ldg(pts + 3 * i + 1), p2 = __ldg(pts + 3 * i + 2);
    const float4* go = reinterpret_cast<const float4*>(g_out + i * 4 * pyr.n);
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
    for (int s = 0; s < pyr.n; ++s) {
        const int D = pyr.dim[s];
        const Cell c = locate(p0, p1, p2, D);
        const float4 g = __ldg(go + s);
        if (g_pts) {
            float4 v[2][2][2];
            load_corners(pyr.vol[s], D, c, v);
            const Jet j = interpolate<1>(v, c);
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[k] = fmaf(c.mult, f4dot(g, j.d[k]), acc[k]);
        }
        if (gv.vol[s]) {
            const float wa[2] = {1.f - c.t[0], c.t[0]}, wb[2] = {1.f - c.t[1], c.t[1]}, wc[2] = {1.f - c.t[2], c.t[2]};
            scatter_corners(gv.vol[s], D, c, wa, wb, wc, g);
        }
    }
    if (g_pts) {
        g_pts[3 * i] = acc[0];
        g_pts[3 * i + 1] = acc[1];
        g_pts[3 * i + 2] = acc[2];
    }
}

// backward of the backward, i.e. the gradient of  L2 = sum_k gg_pts[k] * g_pts[k]  where
// g_pts[k] = sum_c g_out[c] * dfeat_c/dp_k  (what the reference's grad2_3d returns for
// grad2_grad_input = 0, cuda_gridsample.py:110-123):
//   gg_out[c]  = sum_k gg_pts[k] * dfeat_c/dp_k
//   g2_pts[j]  = sum_{k != j} gg_pts[k] * sum_c g_out[c] * d2feat_c/dp_k dp_j   (pure second
//                derivatives of a trilinear interpolant vanish)
//   g2_vol    += g_out * sum_k gg_pts[k] * dw/dp_k
__global__ void __launch_bounds__(256)
trilinear_bwd2_kernel(const float* __restrict__ pts, long long n, Pyr pyr, const float* __restrict__ g_out,
                      const float* __restrict__ gg_pts, float* __restrict__ gg_out, float* __restrict__ g2_pts,
                      GradPyr gv) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p0 = __ldg(pts + 3 * i), p1 = __ldg(pts + 3 * i + 1), p2 = __ldg(pts + 3 * i + 2);
    const float q[3] = {__ldg(gg_pts + 3 * i), __ldg(gg_pts + 3 * i + 1), __ldg(gg_pts + 3 * i + 2)};
    const float4* go = reinterpret_cast<const float4*>(g_out + i * 4 * pyr.n);
    float4* ggo = reinterpret_cast<float4*>(gg_out + i * 4 * pyr.n);
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
    for (int s = 0; s < pyr.n; ++s) {
        const int D = pyr.dim[s];
        const Cell c = locate(p0, p1, p2, D);
        const float4 g = __ldg(go + s);
        float4 v[2][2][2];
        load_corners(pyr.vol[s], D, c, v);
        const Jet j = interpolate<2>(v, c);
        const float m = c.mult;
        float4 r = f4scale(j.d[0], q[0] * m);
        r = f4axpy(q[1] * m, j.d[1], r);
        r = f4axpy(q[2] * m, j.d[2], r);
        ggo[s] = r;
        const float hab = f4dot(g, j.dd[0]) * m * m, hac = f4dot(g, j.dd[1]) * m * m, hbc = f4dot(g, j.dd[2]) * m * m;
        acc[0] += q[1] * hab + q[2] * hac;
        acc[1] += q[0] * hab + q[2] * hbc;
        acc[2] += q[0] * hac + q[1] * hbc;
        if (gv.vol[s]) {
            // sum_k q_k * d(wa wb wc)/dp_k, per corner: scatter three rank-1 terms
            const float wa[2] = {1.f - c.t[0], c.t[0]}, wb[2] = {1.f - c.t[1], c.t[1]}, wc[2] = {1.f - c.t[2], c.t[2]};
            const float da_[2] = {-m * q[0], m * q[0]}, db_[2] = {-m * q[1], m * q[1]}, dc_[2] = {-m * q[2], m * q[2]};
            scatter_corners(gv.vol[s], D, c, da_, wb, wc, g);
            scatter_corners(gv.vol[s], D, c, wa, db_, wc, g);
            scatter_corners(gv.vol[s], D, c, wa, wb, dc_, g);
        }
    }
    g2_pts[3 * i] = acc[0];
    g2_pts[3 * i + 1] = acc[1];
    g2_pts[3 * i + 2] = acc[2];
}

// value + directional derivative along u (forward-mode): f (n,4S), df = J u (n,4S)
__global__ void __launch_bounds__(256)
trilinear_fwd_jvp_kernel(const float* __restrict__ pts, long long n, Pyr pyr, float u0, float u1, float u2,
                         float* __restrict__ out, float* __restrict__ dout) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p0 = __ldg(pts + 3 * i), p1 = __ldg(pts + 3 * i + 1), p2 = __ldg(pts + 3 * i + 2);
    float4* o = reinterpret_cast<float4*>(out + i * 4 * pyr.n);
    float4* od = reinterpret_cast<float4*>(dout + i * 4 * pyr.n);
#pragma unroll 1
    for (int s = 0; s < pyr.n; ++s) {
        const Cell c = locate(p0, p1, p2, pyr.dim[s]);
        float4 v[2][2][2];
        load_corners(pyr.vol[s], pyr.dim[s], c, v);
        const Jet j = interpolate<1>(v, c);
        o[s] = j.f;
        float4 r = f4scale(j.d[0], u0 * c.mult);
        r = f4axpy(u1 * c.mult, j.d[1], r);
        od[s] = f4axpy(u2 * c.mult, j.d[2], r);
    }
}

// reverse-mode through the look-up and its tangent, accumulated onto grad / smooth (n,3):
//   grad   += J^T g          smooth += (H u)^T g + J^T dg
__global__ void __launch_bounds__(256)
trilinear_vjp2_kernel(const float* __restrict__ pts, long long n, Pyr pyr, float u0, float u1, float u2,
                      const float* __restrict__ g_f, const float* __restrict__ dg_f, float* __restrict__ grad,
                      float* __restrict__ smooth) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p0 = __ldg(pts + 3 * i), p1 = __ldg(pts + 3 * i + 1), p2 = __ldg(pts + 3 * i + 2);
    const float4* g4 = reinterpret_cast<const float4*>(g_f + i * 4 * pyr.n);
    const float4* dg4 = reinterpret_cast<const float4*>(dg_f + i * 4 * pyr.n);
    float ga[3] = {0.f, 0.f, 0.f}, sa[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
    for (int s = 0; s < pyr.n; ++s) {
        const Cell c = locate(p0, p1, p2, pyr.dim[s]);
        float4 v[2][2][2];
        load_corners(pyr.vol[s], pyr.dim[s], c, v);
        const Jet j = interpolate<2>(v, c);
        const float4 g = __ldg(g4 + s), dg = __ldg(dg4 + s);
        const float m = c.mult, mm = c.mult * c.mult;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            ga[k] = fmaf(m, f4dot(g, j.d[k]), ga[k]);
            sa[k] = fmaf(m, f4dot(dg, j.d[k]), sa[k]);
        }
        const float hab = f4dot(g, j.dd[0]) * mm, hac = f4dot(g, j.dd[1]) * mm, hbc = f4dot(g, j.dd[2]) * mm;
        sa[0] += u1 * hab + u2 * hac;
        sa[1] += u0 * hab + u2 * hbc;
        sa[2] += u0 * hac + u1 * hbc;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        grad[3 * i + k] += ga[k];
        if (smooth) smooth[3 * i + k] += sa[k];
    }
}

// (1,4,D,D,D) NCDHW <-> channels-last (D,D,D,4)
__global__ void __launch_bounds__(256)
pack_volume_kernel(const float* __restrict__ src, float4* __restrict__ dst, long long d3) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d3) return;
    dst[i] = make_float4(__ldg(src + i), __ldg(src + d3 + i), __ldg(src + 2 * d3 + i), __ldg(src + 3 * d3 + i));
}
__global__ void __launch_bounds__(256)
unpack_volume_kernel(const float4* __restrict__ src, float* __restrict__ dst, long long d3) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d3) return;
    const float4 v = __ldg(src + i);
    dst[i] = v.x; dst[d3 + i] = v.y; dst[2 * d3 + i] = v.z; dst[3 * d3 + i] = v.w;
}

bool fill_pyr(const gens_pyramid_t* p, Pyr& out) {
    if (!p || p->n_scales <= 0 || p->n_scales > GENS_MAX_SCALES) return false;
    out.n = p->n_scales;
    for (int s = 0; s < p->n_scales; ++s) {
        if (!p->vol[s] || p->dim[s] <= 0) return false;
        out.vol[s] = reinterpret_cast<const float4*>(p->vol[s]);
        out.dim[s] = p->dim[s];
    }
    return true;
}

void fill_grad(const gens_pyramid_t* g, int n_scales, GradPyr& out) {
    for (int s = 0; s < GENS_MAX_SCALES; ++s)
        out.vol[s] = (g && s < n_scales) ? reinterpret_cast<float4*>(const_cast<float*>(g->vol[s])) : nullptr;
}

}  // namespace

extern "C" int gens_pack_volume(const float* src_ncdhw, float* dst_channels_last, int D, void* stream) {
    GENS_CHECK_ARG(src_ncdhw && dst_channels_last && D > 0);
    const long long d3 = (long long)D * D * D;
    pack_volume_kernel<<<ceil_div_i(d3, 256), 256, 0, (cudaStream_t)stream>>>(
        src_ncdhw, reinterpret_cast<float4*>(dst_channels_last), d3);
    return gens_launch_status();
}

extern "C" int gens_unpack_volume(const float* src_channels_last, float* dst_ncdhw, int D, void* stream) {
    GENS_CHECK_ARG(src_channels_last && dst_ncdhw && D > 0);
    const long long d3 = (long long)D * D * D;
    unpack_volume_kernel<<<ceil_div_i(d3, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(src_channels_last), dst_ncdhw, d3);
    return gens_launch_status();
}

extern "C" int gens_mask_nearest(const float* pts, long long n, const gens_pyramid_t* masks, int aten_cuda_flavour,
                                 uint8_t* any_out, float* each_out, void* stream) {
    if (n == 0) return 0;
    GENS_CHECK_ARG(pts && masks && n >= 0 && (any_out || each_out));
    if (masks->n_scales <= 0 || masks->n_scales > GENS_MAX_SCALES) return GENS_E_UNSUPPORTED;
    if (n == 0) return 0;
    MaskPyr m;
    m.n = masks->n_scales;
    for (int s = 0; s < m.n; ++s) {
        GENS_CHECK_ARG(masks->vol[s] && masks->dim[s] > 0);
        m.vol[s] = masks->vol[s];
        m.dim[s] = masks->dim[s];
    }
    mask_nearest_kernel<<<ceil_div_i(n, 256), 256, 0, (cudaStream_t)stream>>>(pts, n, m, aten_cuda_flavour, any_out,
                                                                            each_out);
    return gens_launch_status();
}

extern "C" int gens_trilinear_fwd(const float* pts, long long n, const gens_pyramid_t* vols, float* out, void* stream) {
    if (n == 0) return 0;
    GENS_CHECK_ARG(pts && out && n >= 0);
    Pyr p;
    if (!fill_pyr(vols, p)) return GENS_E_BADARG;
    if (n == 0) return 0;
    trilinear_fwd_kernel<<<ceil_div_i(n, 256), 256, 0, (cudaStream_t)stream>>>(pts, n, p, out);
    return gens_launch_status();
}

extern "C" int gens_trilinear_bwd(const float* pts, long long n, const gens_pyramid_t* vols, const float* g_out,
                                  float* g_pts, const gens_pyramid_t* g_vols, void* stream) {
    if (n == 0) return 0;
    GENS_CHECK_ARG(pts && g_out && n >= 0 && (g_pts || g_vols));
    Pyr p;
    if (!fill_pyr(vols, p)) return GENS_E_BADARG;
    if (n == 0) return 0;
    GradPyr gv;
    fill_grad(g_vols, p.n, gv);
    trilinear_bwd_kernel<<<ceil_div_i(n, 256), 256, 0, (cudaStream_t)stream>>>(pts, n, p, g_out, g_pts, gv);
    return gens_launch_status();
}

extern "C" int gens_trilinear_bwd2(const float* pts, long long n, const gens_pyramid_t* vols, const float* g_out,
                                   const float* gg_pts, float* gg_out, float* g2_pts, const gens_pyramid_t* g2_vols,
                                   void* stream) {
    if (n == 0) return 0;
    GENS_CHECK_ARG(pts && g_out && gg_pts && gg_out && g2_pts && n >= 0);
    Pyr p;
    if (!fill_pyr(vols, p)) return GENS_E_BADARG;
    if (n == 0) return 0;
    GradPyr gv;
    fill_grad(g2_vols, p.n, gv);
    trilinear_bwd2_kernel<<<ceil_div_i(n, 256), 256, 0, (cudaStream_t)stream>>>(pts, n, p, g_out, gg_pts, gg_out, g2_pts,
                                                                              gv);
    return gens_launch_status();
}

extern "C" int gens_trilinear_fwd_jvp(const float* pts, long long n, const gens_pyramid_t* vols, const float* u3,
                                      float* out, float* dout, void* stream) {
    if (n == 0) return 0;
    GENS_CHECK_ARG(pts && out && dout && u3 && n >= 0);
    Pyr p;
    if (!fill_pyr(vols, p)) return GENS_E_BADARG;
    if (n == 0) return 0;
    trilinear_fwd_jvp_kernel<<<ceil_div_i(n, 256), 256, 0, (cudaStream_t)stream>>>(pts, n, p, u3[0], u3[1], u3[2], out,
                                                                                  dout);
    return gens_launch_status();
}

extern "C" int gens_trilinear_vjp2(const float* pts, long long n, const gens_pyramid_t* vols, const float* u3,
                                   const float* g_f, const float* dg_f, float* grad, float* smooth, void* stream) {
    if (n == 0) return 0;
    GENS_CHECK_ARG(pts && g_f && dg_f && grad && u3 && n >= 0);
    Pyr p;
    if (!fill_pyr(vols, p)) return GENS_E_BADARG;
    if (n == 0) return 0;
    trilinear_vjp2_kernel<<<ceil_div_i(n, 256), 256, 0, (cudaStream_t)stream>>>(pts, n, p, u3[0], u3[1], u3[2], g_f, dg_f,
                                                                               grad, smooth);
    return gens_launch_status();
}
