// K1: fused multi-view feature-volume aggregation for sm_100a.
//
// Replaces one scale of Volume.agg_mean_var (reference models/modules/volume.py:21-58), which
// runs ~320 ATen ops and >= 15 full (nv,.,D^3) temporaries per scale, with ONE launch that
// reads each feature map once (through L2) and writes the 8+1 output channels once.
//
// Mapping: one thread owns VEC consecutive voxels along the fastest tensor dim (world z), so
// each of the 9 channel planes is written with one coalesced 16-byte streaming store per
// thread (512 B per warp per plane).  Camera matrices live in shared memory; the x/y part
// of the k-ascending projection chain is shared by the VEC voxels of a thread.  Feature
// maps are channels-last (nv,H,W,4): one bilinear corner = one 16-byte read-only load, and
// neighbouring voxels hit neighbouring pixels so the gathers are served by L1/L2.
//
// Arithmetic is the reference's, step for step (see oracle/gens_oracle.c for the CPU
// restatement it is tested against): two-stage projection, IEEE division, the two-moment
// variance E[x^2]-E[x]^2 (NOT Welford: parity with the reference's cancellation behaviour
// is the contract), every rounding spelled with _rn intrinsics.
#include "common.cuh"

namespace {

struct Cam {
    float w2c[16];
    float k[12];
};

struct Proj {
    float ix, iy;
    bool valid;
};

template <bool RECIP>
__device__ __forceinline__ float div_scalar(float a, float b, float inv_b) {
    return RECIP ? __fmul_rn(a, inv_b) : __fdiv_rn(a, b);
}

// Finish the projection of one voxel given the x/y partial sums of the camera transform.
template <bool RECIP>
__device__ __forceinline__ Proj project_finish(const Cam& cam, const float pre[4], float z, float hx,
                                                float hy, float inv_hx, float inv_hy, int W, int H) {
    float c[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float t = __fmaf_rn(cam.w2c[4 * r + 2], z, pre[r]);
        c[r] = __fmaf_rn(cam.w2c[4 * r + 3], 1.0f, t);
    }
    float img[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) img[r] = row_dot4(cam.k + 4 * r, c[0], c[1], c[2], c[3]);
    float den = __fadd_rn(img[2], 1e-8f);
    float x = __fdiv_rn(img[0], den);
    float y = __fdiv_rn(img[1], den);
    float nx = __fsub_rn(div_scalar<RECIP>(x, hx, inv_hx), 1.0f);
    float ny = __fsub_rn(div_scalar<RECIP>(y, hy, inv_hy), 1.0f);
    Proj p;
    p.valid = (fabsf(nx) <= 1.0f) && (fabsf(ny) <= 1.0f) && (img[2] > 0.0f);
    // ATen grid_sampler_unnormalize, align_corners=True
    p.ix = __fmul_rn(__fmul_rn(__fadd_rn(nx, 1.0f), 0.5f), (float)(W - 1));
    p.iy = __fmul_rn(__fmul_rn(__fadd_rn(ny, 1.0f), 0.5f), (float)(H - 1));
    return p;
}

struct Footprint {
    int x0, y0;
    float w_nw, w_ne, w_sw, w_se;
};

__device__ __forceinline__ Footprint footprint(float ix, float iy) {
    float fx0 = floorf(ix), fy0 = floorf(iy);
    float fx1 = __fadd_rn(fx0, 1.0f), fy1 = __fadd_rn(fy0, 1.0f);
    float ax = __fsub_rn(fx1, ix), bx = __fsub_rn(ix, fx0);
    float ay = __fsub_rn(fy1, iy), by = __fsub_rn(iy, fy0);
    Footprint f;
    f.x0 = (int)fx0;
    f.y0 = (int)fy0;
    f.w_nw = __fmul_rn(ax, ay);
    f.w_ne = __fmul_rn(bx, ay);
    f.w_sw = __fmul_rn(ax, by);
    f.w_se = __fmul_rn(bx, by);
    return f;
}

__device__ __forceinline__ void fma4(float4& acc, const float4 v, float w) {
    acc.x = __fmaf_rn(v.x, w, acc.x);
    acc.y = __fmaf_rn(v.y, w, acc.y);
    acc.z = __fmaf_rn(v.z, w, acc.z);
    acc.w = __fmaf_rn(v.w, w, acc.w);
}

// Bilinear sample with zeros padding from a channels-last (H,W,4) map: corners outside the
// map contribute nothing (their weight is still computed from the un-clamped coordinate).
__device__ __forceinline__ float4 sample4(const float4* __restrict__ map, int H, int W, const Footprint& f) {
    const bool x0_in = (unsigned)f.x0 < (unsigned)W, x1_in = (unsigned)(f.x0 + 1) < (unsigned)W;
    const bool y0_in = (unsigned)f.y0 < (unsigned)H, y1_in = (unsigned)(f.y0 + 1) < (unsigned)H;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long base = (long long)f.y0 * W + f.x0;
    // issue all four loads before the first use
    const float4 v_nw = (x0_in && y0_in) ? ldg4(map + base) : zero;
    const float4 v_ne = (x1_in && y0_in) ? ldg4(map + base + 1) : zero;
    const float4 v_sw = (x0_in && y1_in) ? ldg4(map + base + W) : zero;
    const float4 v_se = (x1_in && y1_in) ? ldg4(map + base + W + 1) : zero;
    float4 acc = zero;
    if (x0_in && y0_in) fma4(acc, v_nw, f.w_nw);
    if (x1_in && y0_in) fma4(acc, v_ne, f.w_ne);
    if (x0_in && y1_in) fma4(acc, v_sw, f.w_sw);
    if (x1_in && y1_in) fma4(acc, v_se, f.w_se);
    return acc;
}

__device__ __forceinline__ void load_cams(Cam* s_cam, const float* w2c, const float* k_stage, int nv) {
    for (int i = threadIdx.x; i < nv * 28; i += blockDim.x) {
        int v = i / 28, j = i % 28;
        float val = j < 16 ? w2c[v * 16 + j] : k_stage[v * 16 + (j - 16)];
        (j < 16 ? s_cam[v].w2c[j] : s_cam[v].k[j - 16]) = val;
    }
    __syncthreads();
}

template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) __stcs(p + j, v[j]);
    }
}

template <int VEC, bool RECIP>
__global__ void __launch_bounds__(256)
volume_agg_fwd_kernel(const float4* __restrict__ feat, int nv, int H, int W, const float* __restrict__ w2c,
                      const float* __restrict__ k_stage, const float* __restrict__ grid, int D, int a0,
                      long long n_groups, long long out_off, long long channel_stride, int min_vis_view,
                      float hx, float hy, float inv_hx, float inv_hy, float* __restrict__ volume,
                      float* __restrict__ mask_volume) {
    __shared__ Cam s_cam[GENS_MAX_VIEWS];
    load_cams(s_cam, w2c, k_stage, nv);

    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const long long n0 = g * VEC;  // flat voxel offset inside the slab [a0, a1)
    const int DD = D * D;
    const int a = a0 + (int)(n0 / DD);
    const int rem = (int)(n0 % DD);
    const int b = rem / D, c0 = rem % D;

    const float X = __ldg(grid + a), Y = __ldg(grid + b);
    float Z[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) Z[j] = __ldg(grid + c0 + j);

    float4 s[VEC], q[VEC];
    int cnt[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        s[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        q[j] = s[j];
        cnt[j] = 0;
    }

    const long long map_stride = (long long)H * W;
    for (int v = 0; v < nv; ++v) {
        const Cam& cam = s_cam[v];
        float pre[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) pre[r] = __fmaf_rn(cam.w2c[4 * r + 1], Y, __fmul_rn(cam.w2c[4 * r], X));
        Proj p[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) p[j] = project_finish<RECIP>(cam, pre, Z[j], hx, hy, inv_hx, inv_hy, W, H);
        const float4* map = feat + v * map_stride;
        float4 f[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            f[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p[j].valid) f[j] = sample4(map, H, W, footprint(p[j].ix, p[j].iy));
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            if (p[j].valid) {
                cnt[j] += 1;
                s[j].x = __fadd_rn(s[j].x, f[j].x);
                s[j].y = __fadd_rn(s[j].y, f[j].y);
                s[j].z = __fadd_rn(s[j].z, f[j].z);
                s[j].w = __fadd_rn(s[j].w, f[j].w);
                q[j].x = __fadd_rn(q[j].x, __fmul_rn(f[j].x, f[j].x));
                q[j].y = __fadd_rn(q[j].y, __fmul_rn(f[j].y, f[j].y));
                q[j].z = __fadd_rn(q[j].z, __fmul_rn(f[j].z, f[j].z));
                q[j].w = __fadd_rn(q[j].w, __fmul_rn(f[j].w, f[j].w));
            }
        }
    }

    float out[9][VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const float den = cnt[j] <= 0 ? 1e-8f : (float)cnt[j];
        const float sv[4] = {s[j].x, s[j].y, s[j].z, s[j].w};
        const float qv[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float mean = __fdiv_rn(sv[k], den);
            out[k][j] = mean;
            out[4 + k][j] = __fsub_rn(__fdiv_rn(qv[k], den), __fmul_rn(mean, mean));
        }
        out[8][j] = cnt[j] > min_vis_view ? 1.0f : 0.0f;
    }
    const long long o = out_off + n0;
#pragma unroll
    for (int k = 0; k < 8; ++k) store_vec<VEC>(volume + k * channel_stride + o, out[k]);
    store_vec<VEC>(mask_volume + o, out[8]);
}

template <bool RECIP>
__global__ void __launch_bounds__(256)
volume_project_debug_kernel(int nv, int H, int W, const float* __restrict__ w2c, const float* __restrict__ k_stage,
                            const float* __restrict__ grid, int D, float hx, float hy, float inv_hx, float inv_hy,
                            int32_t* __restrict__ ix0, int32_t* __restrict__ iy0, uint8_t* __restrict__ valid) {
    __shared__ Cam s_cam[GENS_MAX_VIEWS];
    load_cams(s_cam, w2c, k_stage, nv);
    const long long D3 = (long long)D * D * D;
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= D3) return;
    const int a = (int)(n / ((long long)D * D)), rem = (int)(n % ((long long)D * D));
    const float X = grid[a], Y = grid[rem / D], Z = grid[rem % D];
    for (int v = 0; v < nv; ++v) {
        const Cam& cam = s_cam[v];
        float pre[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) pre[r] = __fmaf_rn(cam.w2c[4 * r + 1], Y, __fmul_rn(cam.w2c[4 * r], X));
        Proj p = project_finish<RECIP>(cam, pre, Z, hx, hy, inv_hx, inv_hy, W, H);
        Footprint f = footprint(p.valid ? p.ix : 0.f, p.valid ? p.iy : 0.f);
        ix0[v * D3 + n] = p.valid ? f.x0 : 0;
        iy0[v * D3 + n] = p.valid ? f.y0 : 0;
        valid[v * D3 + n] = p.valid ? 1 : 0;
    }
}

// Backward w.r.t. the feature maps.  With m_v the view validity, n' the clamped count,
//   mean_k = sum_v m_v f_vk / n',  var_k = sum_v m_v f_vk^2 / n' - mean_k^2
//   d/df_vk = m_v / n' * ( g_mean_k + 2 g_var_k (f_vk - mean_k) )
// scattered to the four bilinear corners with 16-byte vector atomics.
template <bool RECIP>
__global__ void __launch_bounds__(256)
volume_agg_bwd_kernel(const float4* __restrict__ feat, int nv, int H, int W, const float* __restrict__ w2c,
                      const float* __restrict__ k_stage, const float* __restrict__ grid, int D, int a0,
                      long long n_vox, long long out_off, long long channel_stride, float hx, float hy,
                      float inv_hx, float inv_hy, const float* __restrict__ grad_volume,
                      float4* __restrict__ grad_feat) {
    __shared__ Cam s_cam[GENS_MAX_VIEWS];
    load_cams(s_cam, w2c, k_stage, nv);
    const long long n0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n0 >= n_vox) return;
    const int DD = D * D;
    const int a = a0 + (int)(n0 / DD);
    const int rem = (int)(n0 % DD);
    const float X = __ldg(grid + a), Y = __ldg(grid + rem / D), Z = __ldg(grid + rem % D);

    const long long o = out_off + n0;
    float gm[4], gv[4];
    bool any = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        gm[k] = __ldg(grad_volume + k * channel_stride + o);
        gv[k] = __ldg(grad_volume + (4 + k) * channel_stride + o);
        any |= (gm[k] != 0.f) | (gv[k] != 0.f);
    }
    if (!any) return;

    const long long map_stride = (long long)H * W;
    float4 f[GENS_MAX_VIEWS];
    Footprint fp[GENS_MAX_VIEWS];
    unsigned valid_bits = 0;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt = 0;
#pragma unroll 1
    for (int v = 0; v < nv; ++v) {
        const Cam& cam = s_cam[v];
        float pre[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) pre[r] = __fmaf_rn(cam.w2c[4 * r + 1], Y, __fmul_rn(cam.w2c[4 * r], X));
        Proj p = project_finish<RECIP>(cam, pre, Z, hx, hy, inv_hx, inv_hy, W, H);
        if (p.valid) {
            fp[v] = footprint(p.ix, p.iy);
            f[v] = sample4(feat + v * map_stride, H, W, fp[v]);
            valid_bits |= 1u << v;
            cnt += 1;
            s.x += f[v].x; s.y += f[v].y; s.z += f[v].z; s.w += f[v].w;
        }
    }
    if (cnt == 0) return;
    const float inv_n = 1.0f / (float)cnt;
    const float mean[4] = {s.x * inv_n, s.y * inv_n, s.z * inv_n, s.w * inv_n};
#pragma unroll 1
    for (int v = 0; v < nv; ++v) {
        if (!((valid_bits >> v) & 1u)) continue;
        const float fv[4] = {f[v].x, f[v].y, f[v].z, f[v].w};
        float gf[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) gf[k] = inv_n * (gm[k] + 2.0f * gv[k] * (fv[k] - mean[k]));
        float4* gmap = grad_feat + v * map_stride;
        const Footprint& t = fp[v];
        const bool x1_in = t.x0 + 1 < W, y1_in = t.y0 + 1 < H;
        const long long base = (long long)t.y0 * W + t.x0;
        auto scat = [&](long long idx, float w) {
            atomicAdd(gmap + idx, make_float4(gf[0] * w, gf[1] * w, gf[2] * w, gf[3] * w));
        };
        scat(base, t.w_nw);
        if (x1_in) scat(base + 1, t.w_ne);
        if (y1_in) scat(base + W, t.w_sw);
        if (x1_in && y1_in) scat(base + W + 1, t.w_se);
    }
}

__global__ void __launch_bounds__(256)
nchw4_to_nhwc4_kernel(const float* __restrict__ src, float4* __restrict__ dst, long long hw, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long n = i / hw, p = i % hw;
    const float* s = src + n * 4 * hw + p;
    dst[i] = make_float4(__ldg(s), __ldg(s + hw), __ldg(s + 2 * hw), __ldg(s + 3 * hw));
}

struct HalfExtent {
    float hx, hy, inv_hx, inv_hy;
};
inline HalfExtent half_extent(int W, int H) {
    HalfExtent e;
    e.hx = (float)((double)(W - 1) / 2.0);  // python: (width - 1) / 2, then cast to the tensor dtype
    e.hy = (float)((double)(H - 1) / 2.0);
    e.inv_hx = 1.0f / e.hx;  // ATen CUDA div_true: opmath_t(1.0) / scalar
    e.inv_hy = 1.0f / e.hy;
    return e;
}

}  // namespace

extern "C" int gens_nchw4_to_nhwc4(const float* src, float* dst, int n, int h, int w, void* stream) {
    GENS_CHECK_ARG(src && dst && n > 0 && h > 0 && w > 0);
    const long long hw = (long long)h * w, total = hw * n;
    nchw4_to_nhwc4_kernel<<<ceil_div_i(total, 256), 256, 0, (cudaStream_t)stream>>>(
        src, reinterpret_cast<float4*>(dst), hw, total);
    return gens_launch_status();
}

extern "C" int gens_volume_agg_fwd(const float* feat_nhwc, int nv, int H, int W, const float* w2c,
                                   const float* k_stage, const float* grid, int D, int a0, int a1, int a_base,
                                   long long channel_stride, int min_vis_view, int div_mode, float* volume,
                                   float* mask_volume, void* stream) {
    GENS_CHECK_ARG(feat_nhwc && w2c && k_stage && grid && volume && mask_volume);
    GENS_CHECK_ARG(nv > 0 && H > 0 && W > 0 && D > 0 && a0 >= 0 && a1 <= D && a_base >= 0 && a_base <= a0);
    if (nv > GENS_MAX_VIEWS) return GENS_E_UNSUPPORTED;
    if (a1 <= a0) return 0;
    const long long n_vox = (long long)(a1 - a0) * D * D;
    const long long out_off = (long long)(a0 - a_base) * D * D;
    const HalfExtent e = half_extent(W, H);
    const bool vec4 = (D % 4 == 0) && (channel_stride % 4 == 0) &&
                      (((uintptr_t)volume | (uintptr_t)mask_volume) % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const float4* feat = reinterpret_cast<const float4*>(feat_nhwc);
#define GENS_LAUNCH_AGG(VEC, RECIP)                                                                          \
    volume_agg_fwd_kernel<VEC, RECIP><<<ceil_div_i(n_vox / VEC, 256), 256, 0, st>>>(                          \
        feat, nv, H, W, w2c, k_stage, grid, D, a0, n_vox / VEC, out_off, channel_stride, min_vis_view, e.hx, \
        e.hy, e.inv_hx, e.inv_hy, volume, mask_volume)
    if (vec4) {
        if (div_mode == GENS_DIV_RECIP) GENS_LAUNCH_AGG(4, true); else GENS_LAUNCH_AGG(4, false);
    } else {
        if (div_mode == GENS_DIV_RECIP) GENS_LAUNCH_AGG(1, true); else GENS_LAUNCH_AGG(1, false);
    }
#undef GENS_LAUNCH_AGG
    return gens_launch_status();
}

extern "C" int gens_volume_project_debug(int nv, int H, int W, const float* w2c, const float* k_stage,
                                         const float* grid, int D, int div_mode, int32_t* ix0, int32_t* iy0,
                                         uint8_t* valid, void* stream) {
    GENS_CHECK_ARG(w2c && k_stage && grid && ix0 && iy0 && valid && nv > 0 && D > 0 && H > 0 && W > 0);
    if (nv > GENS_MAX_VIEWS) return GENS_E_UNSUPPORTED;
    const HalfExtent e = half_extent(W, H);
    const long long D3 = (long long)D * D * D;
    cudaStream_t st = (cudaStream_t)stream;
    if (div_mode == GENS_DIV_RECIP)
        volume_project_debug_kernel<true><<<ceil_div_i(D3, 256), 256, 0, st>>>(
            nv, H, W, w2c, k_stage, grid, D, e.hx, e.hy, e.inv_hx, e.inv_hy, ix0, iy0, valid);
    else
        volume_project_debug_kernel<false><<<ceil_div_i(D3, 256), 256, 0, st>>>(
            nv, H, W, w2c, k_stage, grid, D, e.hx, e.hy, e.inv_hx, e.inv_hy, ix0, iy0, valid);
    return gens_launch_status();
}

extern "C" int gens_volume_agg_bwd(const float* feat_nhwc, int nv, int H, int W, const float* w2c,
                                   const float* k_stage, const float* grid, int D, int a0, int a1, int a_base,
                                   long long channel_stride, int div_mode, const float* grad_volume,
                                   float* grad_feat_nhwc, void* stream) {
    GENS_CHECK_ARG(feat_nhwc && w2c && k_stage && grid && grad_volume && grad_feat_nhwc);
    GENS_CHECK_ARG(nv > 0 && H > 0 && W > 0 && D > 0 && a0 >= 0 && a1 <= D && a_base >= 0 && a_base <= a0);
    if (nv > GENS_MAX_VIEWS) return GENS_E_UNSUPPORTED;
    if (a1 <= a0) return 0;
    const long long n_vox = (long long)(a1 - a0) * D * D;
    const long long out_off = (long long)(a0 - a_base) * D * D;
    const HalfExtent e = half_extent(W, H);
    cudaStream_t st = (cudaStream_t)stream;
    const float4* feat = reinterpret_cast<const float4*>(feat_nhwc);
    float4* gfeat = reinterpret_cast<float4*>(grad_feat_nhwc);
    if (div_mode == GENS_DIV_RECIP)
        volume_agg_bwd_kernel<true><<<ceil_div_i(n_vox, 256), 256, 0, st>>>(
            feat, nv, H, W, w2c, k_stage, grid, D, a0, n_vox, out_off, channel_stride, e.hx, e.hy, e.inv_hx,
            e.inv_hy, grad_volume, gfeat);
    else
        volume_agg_bwd_kernel<false><<<ceil_div_i(n_vox, 256), 256, 0, st>>>(
            feat, nv, H, W, w2c, k_stage, grid, D, a0, n_vox, out_off, channel_stride, e.hx, e.hy, e.inv_hx,
            e.inv_hy, grad_volume, gfeat);
    return gens_launch_status();
}
