// K8: plane-induced homography patch warp of the feature-metric consistency term (sm_100a, inference).
//
// Replaces surface_patch_warp + patch_homography (reference models/modules/projector.py:353-437): per ray the
// surface point X (reference camera frame) and its normal n define the plane n.X = d; a patch of p x p
// reference pixels around the projection of X is mapped into every source view by
//     H_s = K_s (R_s^T R_0 + R_s^T (c_0 - c_s) n^T / (n.X + 1e-10)) K_0^-1
// and the C-channel feature image is sampled bilinearly (zeros padding, align_corners = True) at the warped
// pixels (sources) and at the patch pixels themselves (reference view).  The reference does this with an
// einsum, a dozen small ATen ops and two cuDNN grid_sampler launches over (B p^2 (ns+1)) x C gathers; here one
// thread owns one (view, ray, patch pixel), rebuilds the 3x3 chain from the per-view matrices staged in shared
// memory, and gathers its C channels from the NCHW planes.
// Bound: L2 gather traffic, 16 B x C per sample (images stay NCHW: re-packing 3 x 12 x H x W floats per call would
// cost more than it saves).
#include "common.cuh"

namespace {

constexpr int kMaxViews = GENS_MAX_VIEWS;

struct ViewMats {
    float k[9];      // K_s (3x3 of the 4x4 intrinsics)
    float rrel[9];   // R_s^T R_0
    float trel[3];   // R_s^T (c_0 - c_s)
};

__device__ __forceinline__ void mat3_mul(const float* a, const float* b, float* o) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}

__global__ void __launch_bounds__(256)
patch_warp_kernel(const float* __restrict__ pts, const float* __restrict__ nrm, const float* __restrict__ images,
                  const float* __restrict__ intrinsics, const float* __restrict__ poses, const float* __restrict__ k0_inv4,
                  int n_rays, int nv, int C, int H, int W, int patch, float* __restrict__ ref_out,
                  float* __restrict__ src_out) {
    __shared__ ViewMats s_view[kMaxViews];
    __shared__ float s_r0[9], s_c0[3], s_k0[9], s_k0inv[9];
    if (threadIdx.x < nv) {
        const int v = threadIdx.x;
        const float* P = poses + 16 * v;
        const float* K = intrinsics + 16 * v;
        ViewMats m;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                m.k[3 * i + j] = K[4 * i + j];
                // (R_v^T R_0)[i][j] = sum_k R_v[k][i] R_0[k][j]
                m.rrel[3 * i + j] = P[i] * poses[j] + P[4 + i] * poses[4 + j] + P[8 + i] * poses[8 + j];
            }
#pragma unroll
        for (int i = 0; i < 3; ++i)
            m.trel[i] = P[i] * (poses[3] - P[3]) + P[4 + i] * (poses[7] - P[7]) + P[8 + i] * (poses[11] - P[11]);
        s_view[v] = m;
    }
    if (threadIdx.x < 9) {
        const int i = threadIdx.x / 3, j = threadIdx.x % 3;
        s_r0[threadIdx.x] = poses[4 * i + j];
        s_k0[threadIdx.x] = intrinsics[4 * i + j];
        s_k0inv[threadIdx.x] = k0_inv4[4 * i + j];
        if (threadIdx.x < 3) s_c0[threadIdx.x] = poses[4 * threadIdx.x + 3];
    }
    __syncthreads();

    const int pp = patch * patch, half = patch / 2;
    const long long total = (long long)nv * n_rays * pp;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int q = (int)(t % pp);
    const long long vb = t / pp;
    const int b = (int)(vb % n_rays), v = (int)(vb / n_rays);

    // surface point in the reference camera frame, its projection, the patch pixel
    const float px = pts[3 * b], py = pts[3 * b + 1], pz = pts[3 * b + 2];
    float X[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
        X[j] = (px * s_r0[j] + py * s_r0[3 + j] + pz * s_r0[6 + j]) -
               (s_c0[0] * s_r0[j] + s_c0[1] * s_r0[3 + j] + s_c0[2] * s_r0[6 + j]);
    float proj[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) proj[i] = X[0] * s_k0[3 * i] + X[1] * s_k0[3 * i + 1] + X[2] * s_k0[3 * i + 2];
    const float u = proj[0] / (proj[2] + 1e-8f) + (float)(q % patch - half);
    const float w_ = proj[1] / (proj[2] + 1e-8f) + (float)(q / patch - half);

    float gx, gy;  // pixel coordinates in view v
    if (v == 0) {
        gx = u;
        gy = w_;
    } else {
        const float nx = nrm[3 * b], ny = nrm[3 * b + 1], nz = nrm[3 * b + 2];
        const float disp = nx * X[0] + ny * X[1] + nz * X[2] + 1e-10f;
        const ViewMats& m = s_view[v];
        float M[9], KM[9], Hm[9];
        const float n3[3] = {nx, ny, nz};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) M[3 * i + j] = m.rrel[3 * i + j] + m.trel[i] * n3[j] / disp;
        mat3_mul(m.k, M, KM);
        mat3_mul(KM, s_k0inv, Hm);
        const float wx = Hm[0] * u + Hm[1] * w_ + Hm[2];
        const float wy = Hm[3] * u + Hm[4] * w_ + Hm[5];
        const float wz = Hm[6] * u + Hm[7] * w_ + Hm[8];
        gx = wx / (wz + 1e-8f);
        gy = wy / (wz + 1e-8f);
    }
    // the reference normalises to [-1,1] and grid_sample un-normalises again (align_corners = True)
    const float ix = ((2.0f * gx / (float)(W - 1) - 1.0f) + 1.0f) * 0.5f * (float)(W - 1);
    const float iy = ((2.0f * gy / (float)(H - 1) - 1.0f) + 1.0f) * 0.5f * (float)(H - 1);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float bx = ix - fx0, by = iy - fy0, ax = 1.0f - bx, ay = 1.0f - by;
    // a NaN / huge coordinate fails every bounds test below: zeros, as ATen's within_bounds_2d
    const bool in_x0 = fx0 >= 0.f && fx0 <= (float)(W - 1), in_x1 = fx0 + 1.f >= 0.f && fx0 + 1.f <= (float)(W - 1);
    const bool in_y0 = fy0 >= 0.f && fy0 <= (float)(H - 1), in_y1 = fy0 + 1.f >= 0.f && fy0 + 1.f <= (float)(H - 1);
    const int x0 = in_x0 ? (int)fx0 : 0, x1 = in_x1 ? (int)fx0 + 1 : 0, y0 = in_y0 ? (int)fy0 : 0, y1 = in_y1 ? (int)fy0 + 1 : 0;
    const float w_nw = (in_x0 && in_y0) ? ax * ay : 0.f, w_ne = (in_x1 && in_y0) ? bx * ay : 0.f;
    const float w_sw = (in_x0 && in_y1) ? ax * by : 0.f, w_se = (in_x1 && in_y1) ? bx * by : 0.f;
    const long long hw = (long long)H * W;
    const float* img = images + (long long)v * C * hw;
    float* out = (v == 0 ? ref_out : src_out + (long long)(v - 1) * n_rays * pp * C) + ((long long)b * pp + q) * C;
    const int o_nw = y0 * W + x0, o_ne = y0 * W + x1, o_sw = y1 * W + x0, o_se = y1 * W + x1;
    for (int c = 0; c < C; ++c) {
        const float* pl = img + c * hw;
        out[c] = __ldg(pl + o_nw) * w_nw + __ldg(pl + o_ne) * w_ne + __ldg(pl + o_sw) * w_sw + __ldg(pl + o_se) * w_se;
    }
}

}  // namespace

// pts / nrm (n_rays,3): surface points (world) and unit normals (reference camera frame); images (nv,C,H,W);
// intrinsics / poses (nv,4,4); k0_inv4 (4,4) = inverse(intrinsics[0]); ref_out (n_rays,p*p,C),
// src_out (nv-1,n_rays,p*p,C).
extern "C" int gens_patch_warp(const float* pts, const float* nrm, const float* images, const float* intrinsics,
                               const float* poses, const float* k0_inv4, int n_rays, int nv, int channels, int H, int W,
                               int patch, float* ref_out, float* src_out, void* stream) {
    if (n_rays == 0) return 0;
    GENS_CHECK_ARG(pts && nrm && images && intrinsics && poses && k0_inv4 && ref_out && (src_out || nv == 1));
    GENS_CHECK_ARG(n_rays > 0 && nv > 0 && channels > 0 && H > 1 && W > 1 && patch > 0 && (patch & 1));
    if (nv > kMaxViews) return GENS_E_UNSUPPORTED;
    const long long total = (long long)nv * n_rays * patch * patch;
    patch_warp_kernel<<<ceil_div_i(total, 256), 256, 0, (cudaStream_t)stream>>>(
        pts, nrm, images, intrinsics, poses, k0_inv4, n_rays, nv, channels, H, W, patch, ref_out, src_out);
    return gens_launch_status();
}
