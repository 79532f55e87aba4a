/*
 * gens_b200.h -- C ABI of the B200-native GenS hot path (libgens_b200.so).
 *
 * The reference (prstrive/GenS) has no FFI table: its boundary is a set of Python
 * callables plus one pybind module (`gridsample_grad2`, reference
 * models/modules/grid_sample_cuda/gridsample_cuda.cpp:53-56).  The entry points below are
 * what a maintainer binds instead (ctypes stub in INTEGRATION.md); each one cites the
 * reference interface it replaces.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer (fp32
 *     unless noted) that the caller allocated -- the library never allocates or frees;
 *   - work is enqueued on the cudaStream_t passed as `stream` (void*), never synchronised;
 *   - returns 0 on success, a negative GENS_E_* code for argument errors, or a positive
 *     cudaError_t if the launch failed; gens_error_string() explains either;
 *   - no global state, callable from any host thread.
 */
#ifndef GENS_B200_H_
#define GENS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GENS_ABI_VERSION 1

#define GENS_E_BADARG (-1)      /* null pointer / non-positive size            */
#define GENS_E_UNSUPPORTED (-2) /* shape outside what the kernels are built for */

/* tensor / python-scalar division flavour of `norm_x = x / ((W-1)/2) - 1`
 * (reference models/modules/volume.py:38-39).  ATen's CPU kernel divides, its CUDA kernel
 * multiplies by the fp32 reciprocal; masks are bit-exact against whichever is selected. */
#define GENS_DIV_TRUE 0
#define GENS_DIV_RECIP 1

int gens_abi_version(void);
const char *gens_error_string(int code);

/* ---- layout helper ---------------------------------------------------------------- */
/* (n,4,h,w) NCHW -> (n,h,w,4) NHWC so that one bilinear corner is a single 16-byte load. */
int gens_nchw4_to_nhwc4(const float *src, float *dst, int n, int h, int w, void *stream);

/* ---- K1: multi-view feature-volume aggregation ---------------------------------------
 * Replaces one scale of Volume.agg_mean_var (reference models/modules/volume.py:21-58):
 * project every voxel centre into every view, bilinear-sample the 4-channel feature map
 * (zeros padding, align_corners=True), masked sum / sum-of-squares / count over views,
 * write [mean(4), var(4)] and the visibility mask (count > min_vis_view).
 *
 *   feat_nhwc (nv,H,W,4)   w2c (nv,4,4) = inverse(c2ws)   k_stage (nv,4,4) rows 0-1 scaled
 *   grid (D) = linspace(-1,1,D)
 *   planes [a0,a1) of tensor dim 2 (world x) are produced -- slab sharding.
 *   volume: channel c, voxel (a,b,c') is written at
 *           volume[c*channel_stride + ((a-a_base)*D + b)*D + c'], same for mask (1 channel)
 *           (a_base = a0, channel_stride = (a1-a0)*D*D for a slab buffer;
 *            a_base = 0,  channel_stride = D*D*D       for the full tensor).
 */
int gens_volume_agg_fwd(const float *feat_nhwc, int nv, int H, int W, const float *w2c,
                        const float *k_stage, const float *grid, int D, int a0, int a1,
                        int a_base, long long channel_stride, int min_vis_view, int div_mode,
                        float *volume, float *mask_volume, void *stream);

/* Debug/parity view of K1's projection stage: per (view, voxel) the floor corner index of
 * the bilinear footprint and the validity bit (volume.py:43).  Outputs are (nv, D,D,D);
 * ix0/iy0 are 0 where the view is invalid. */
int gens_volume_project_debug(int nv, int H, int W, const float *w2c, const float *k_stage,
                              const float *grid, int D, int div_mode, int32_t *ix0,
                              int32_t *iy0, uint8_t *valid, void *stream);

/* Backward of K1 w.r.t. the feature maps (the voxel grid is under no_grad in the
 * reference, volume.py:27-44).  grad_volume addressed like `volume` above; grad_feat_nhwc
 * (nv,H,W,4) must be zero-initialised by the caller (atomic scatter). */
int gens_volume_agg_bwd(const float *feat_nhwc, int nv, int H, int W, const float *w2c,
                        const float *k_stage, const float *grid, int D, int a0, int a1,
                        int a_base, long long channel_stride, int div_mode,
                        const float *grad_volume, float *grad_feat_nhwc, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GENS_B200_H_ */
