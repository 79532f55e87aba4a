/*
 * gens_oracle.c -- CPU restatement of the GenS hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  Nothing under
 * gens_b200/ imports, links or executes it.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit (indices, masks)
 * and to 1e-5 (values) against golden vectors produced by importing the unmodified
 * reference from /root/reference on CPU (tests/golden/make_golden.py, fixtures in
 * tests/golden/ (npz files), test in tests/test_oracle_golden.py).  The reference itself has
 * no tests or golden vectors.
 *
 * The arithmetic that lives in a third-party dependency (PyTorch ATen 2.11:
 * grid_sampler 2-D/3-D, matmul with K=4) is restated from its observable behaviour:
 *   - matmul (.,4,4)@(.,4,N), N>=64: k-ascending chain  t=a0*x0; t=fma(a1,x1,t); ...
 *     (found empirically against torch CPU, bit-exact on 3e6 values)
 *   - grid_sampler un-normalise, align_corners=True :  ((c+1)*0.5)*(size-1)
 *     align_corners=False:  ((c+1)*size-1)*0.5      (ATen GridSampler.cuh)
 *   - bilinear / trilinear: floor corners, weights as products of (corner+1-c) and
 *     (c-corner), zero padding = out-of-range corners contribute nothing.
 *
 * Compile with -ffp-contract=off so that only the explicit fmaf() calls fuse.
 * All entry points are plain C, no global state; OpenMP over the outermost loop.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GENS_DIV_TRUE 0   /* tensor / python_scalar as on the reference's CPU path  */
#define GENS_DIV_RECIP 1  /* ... as ATen's CUDA div_true does it: a * (1.0f / b)      */

/* (4x4) row r times column vector, k-ascending fma chain (see header). */
static inline float row_dot4(const float *a, float x0, float x1, float x2, float x3) {
    float t = a[0] * x0;
    t = fmaf(a[1], x1, t);
    t = fmaf(a[2], x2, t);
    t = fmaf(a[3], x3, t);
    return t;
}

static inline float div_by_scalar(float a, float b, int div_mode) {
    if (div_mode == GENS_DIV_RECIP) {
        float inv = 1.0f / b;
        return a * inv;
    }
    return a / b;
}

/*
 * Project one world point into one view.  Follows volume.py:31-43 of the reference:
 *   cam = inverse(c2w) @ [X,1];  img = K_stage @ cam;  xy = img[:2] / (img[2] + 1e-8)
 *   n = xy / ((size-1)/2) - 1;   mask = |nx|<=1 & |ny|<=1 & img_z>0
 * w2c and K are the 4x4 matrices the host already prepared (torch.inverse / row scaling).
 */
static inline int project_voxel(const float *w2c, const float *K, float X, float Y, float Z,
                                int W, int H, int div_mode, float *nx, float *ny) {
    float cam[4], img[3];
    for (int r = 0; r < 4; ++r) cam[r] = row_dot4(w2c + 4 * r, X, Y, Z, 1.0f);
    for (int r = 0; r < 3; ++r) img[r] = row_dot4(K + 4 * r, cam[0], cam[1], cam[2], cam[3]);
    float den = img[2] + 1e-8f;
    float x = img[0] / den;
    float y = img[1] / den;
    float hx = (float)((double)(W - 1) / 2.0);
    float hy = (float)((double)(H - 1) / 2.0);
    *nx = div_by_scalar(x, hx, div_mode) - 1.0f;
    *ny = div_by_scalar(y, hy, div_mode) - 1.0f;
    return (fabsf(*nx) <= 1.0f) && (fabsf(*ny) <= 1.0f) && (img[2] > 0.0f);
}

/* Bilinear sample of C channel planes (NCHW, one view) at pixel coords (ix,iy), zeros pad. */
static inline void bilinear_zeros(const float *feat, int C, int H, int W, float ix, float iy,
                                  float *out, int *ix0_out, int *iy0_out) {
    float fx0 = floorf(ix), fy0 = floorf(iy);
    float fx1 = fx0 + 1.0f, fy1 = fy0 + 1.0f;
    float w_nw = (fx1 - ix) * (fy1 - iy);
    float w_ne = (ix - fx0) * (fy1 - iy);
    float w_sw = (fx1 - ix) * (iy - fy0);
    float w_se = (ix - fx0) * (iy - fy0);
    /* int conversion is safe: callers only use indices when |coord| is modest */
    long x0 = (long)fx0, y0 = (long)fy0, x1 = x0 + 1, y1 = y0 + 1;
    if (ix0_out) *ix0_out = (int)x0;
    if (iy0_out) *iy0_out = (int)y0;
    int in_x0 = x0 >= 0 && x0 < W, in_x1 = x1 >= 0 && x1 < W;
    int in_y0 = y0 >= 0 && y0 < H, in_y1 = y1 >= 0 && y1 < H;
    for (int c = 0; c < C; ++c) {
        const float *p = feat + (size_t)c * H * W;
        float acc = 0.0f;
        if (in_x0 && in_y0) acc = fmaf(p[y0 * W + x0], w_nw, acc);
        if (in_x1 && in_y0) acc = fmaf(p[y0 * W + x1], w_ne, acc);
        if (in_x0 && in_y1) acc = fmaf(p[y1 * W + x0], w_sw, acc);
        if (in_x1 && in_y1) acc = fmaf(p[y1 * W + x1], w_se, acc);
        out[c] = acc;
    }
}

/*
 * Volume.agg_mean_var for ONE scale (reference volume.py:21-58).
 *   feat   (nv, C, H, W)      K (nv,4,4) already stage-scaled      w2c (nv,4,4)
 *   grid   (D) = torch.linspace(-1,1,D)
 *   planes [a0, a1) of tensor dim 2 (world x) are produced (slab sharding); outputs are
 *   indexed as full (2C, D, D, D) / (D, D, D) arrays.
 *   dbg_*  optional (nv, D,D,D): floor corner indices and per-view validity.
 */
int gens_oracle_volume_agg(const float *feat, int nv, int C, int H, int W, const float *w2c,
                           const float *K, const float *grid, int D, int a0, int a1,
                           int min_vis_view, int div_mode, float *volume, float *mask_volume,
                           int32_t *dbg_ix0, int32_t *dbg_iy0, uint8_t *dbg_mask) {
    if (C > 16 || nv > 16) return -1;
    const size_t D3 = (size_t)D * D * D;
#pragma omp parallel for collapse(2) schedule(static)
    for (int a = a0; a < a1; ++a) {
        for (int b = 0; b < D; ++b) {
            for (int c = 0; c < D; ++c) {
                size_t n = ((size_t)a * D + b) * D + c;
                float s[16], q[16], f[16];
                for (int k = 0; k < C; ++k) s[k] = q[k] = 0.0f;
                int cnt = 0;
                for (int v = 0; v < nv; ++v) {
                    float nx, ny;
                    int m = project_voxel(w2c + 16 * v, K + 16 * v, grid[a], grid[b], grid[c], W, H,
                                          div_mode, &nx, &ny);
                    float ix = ((nx + 1.0f) * 0.5f) * (float)(W - 1);
                    float iy = ((ny + 1.0f) * 0.5f) * (float)(H - 1);
                    int ix0 = 0, iy0 = 0; /* corner indices are only defined where the view is valid */
                    if (m) bilinear_zeros(feat + (size_t)v * C * H * W, C, H, W, ix, iy, f, &ix0, &iy0);
                    if (dbg_ix0) dbg_ix0[v * D3 + n] = ix0;
                    if (dbg_iy0) dbg_iy0[v * D3 + n] = iy0;
                    if (dbg_mask) dbg_mask[v * D3 + n] = (uint8_t)m;
                    if (m) {
                        cnt += 1;
                        for (int k = 0; k < C; ++k) {
                            float sq = f[k] * f[k];
                            s[k] = s[k] + f[k];
                            q[k] = q[k] + sq;
                        }
                    }
                }
                float den = cnt <= 0 ? 1e-8f : (float)cnt;
                for (int k = 0; k < C; ++k) {
                    float mean = s[k] / den;
                    float msq = mean * mean;
                    volume[(size_t)k * D3 + n] = mean;
                    volume[(size_t)(C + k) * D3 + n] = q[k] / den - msq;
                }
                mask_volume[n] = cnt > min_vis_view ? 1.0f : 0.0f;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * lookup_volume (reference projector.py:217-245).  A point p=(p0,p1,p2) is flipped to the grid
 * (x,y,z) = (p2,p1,p0), so p0 indexes tensor dim 2 (D), p1 dim 3 (H), p2 dim 4 (W).
 * ------------------------------------------------------------------------------------------ */

/* F.grid_sample(mode='nearest', padding_mode='zeros', align_corners=False) of one (D,D,D) volume.
 * fused = 1 reproduces ATen's CUDA un-normalise (fma contraction of (c+1)*size-1). */
int gens_oracle_nearest(const float *pts, long n, const float *vol, int D, int fused, float *out) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        float idx[3];
        for (int k = 0; k < 3; ++k) {
            float t = pts[3 * i + k] + 1.0f, s = (float)D;
            float u = fused ? fmaf(t, s, -1.0f) : (t * s - 1.0f);
            idx[k] = nearbyintf(u * 0.5f); /* round half to even */
        }
        float v = 0.0f;
        if (idx[0] >= 0 && idx[0] < D && idx[1] >= 0 && idx[1] < D && idx[2] >= 0 && idx[2] < D)
            v = vol[((long)idx[0] * D + (long)idx[1]) * D + (long)idx[2]];
        out[i] = v;
    }
    return 0;
}

/* grid_sample 3-D, bilinear, zeros padding, align_corners=True of one (C,D,D,D) volume -> out (n,C).
 * Weights and accumulation order as ATen's grid_sampler_3d kernel (tnw..bse, out += val*w). */
int gens_oracle_trilinear(const float *pts, long n, const float *vol, int C, int D, float *out) {
    const long D3 = (long)D * D * D;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        /* grid x <- p2 (W), y <- p1 (H), z <- p0 (D) */
        float ix = ((pts[3 * i + 2] + 1.0f) * 0.5f) * (float)(D - 1);
        float iy = ((pts[3 * i + 1] + 1.0f) * 0.5f) * (float)(D - 1);
        float iz = ((pts[3 * i + 0] + 1.0f) * 0.5f) * (float)(D - 1);
        float x0 = floorf(ix), y0 = floorf(iy), z0 = floorf(iz);
        float x1 = x0 + 1.0f, y1 = y0 + 1.0f, z1 = z0 + 1.0f;
        float w[8];
        w[0] = (x1 - ix) * (y1 - iy) * (z1 - iz); /* tnw */
        w[1] = (ix - x0) * (y1 - iy) * (z1 - iz); /* tne */
        w[2] = (x1 - ix) * (iy - y0) * (z1 - iz); /* tsw */
        w[3] = (ix - x0) * (iy - y0) * (z1 - iz); /* tse */
        w[4] = (x1 - ix) * (y1 - iy) * (iz - z0); /* bnw */
        w[5] = (ix - x0) * (y1 - iy) * (iz - z0); /* bne */
        w[6] = (x1 - ix) * (iy - y0) * (iz - z0); /* bsw */
        w[7] = (ix - x0) * (iy - y0) * (iz - z0); /* bse */
        for (int c = 0; c < C; ++c) {
            float acc = 0.0f;
            for (int k = 0; k < 8; ++k) {
                float xf = (k & 1) ? x1 : x0, yf = (k & 2) ? y1 : y0, zf = (k & 4) ? z1 : z0;
                if (xf >= 0 && xf < D && yf >= 0 && yf < D && zf >= 0 && zf < D)
                    acc = fmaf(vol[c * D3 + ((long)zf * D + (long)yf) * D + (long)xf], w[k], acc);
            }
            out[i * C + c] = acc;
        }
    }
    return 0;
}
