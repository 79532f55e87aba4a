// K9: masked total variation of the volume pyramid in ONE pass (sm_100a).
//
// Replaces ImplicitSurface.tv_regularization (reference models/modules/implicit_surface.py:135-150),
// which every render_core call runs (:260) as ~12 full-volume ATen passes per scale (three shifted
// products of the mask, three shifted differences of the volume, squares, masked sums): 307 MB x ~10 of
// traffic per call, 75-1200 calls per validated image.  Here each voxel is visited once: a thread reads the
// voxel, its +1 neighbour along each tensor axis (the +z one is the next lane's value, the +y / +x ones hit
// L1 / L2) for every channel, and the block reduces four sums per scale:
//     tx, ty, tz = sum over channels and voxels of (v[+1] - v)^2 where mask[+1] * mask > 0
//     cnt        = number of x-axis pairs with mask[+1] * mask > 0   (the reference's mx.sum(), which
//                  normalises all three axes)
// Accumulation is fp32 per thread (at most C terms), fp64 from the warp reduction upwards.
// Bound: HBM, D^3 * (C+1) * 4 bytes read once per scale.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

struct TvArgs {
    const float* vol[GENS_MAX_SCALES];
    const float* mask[GENS_MAX_SCALES];
    int dim[GENS_MAX_SCALES];
    int n;
    int channels;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kThreads) tv_reduce_kernel(TvArgs a, double* __restrict__ out) {
    const int s = blockIdx.y;
    const int D = a.dim[s];
    const long long D2 = (long long)D * D, D3 = D2 * D;
    const float* __restrict__ vol = a.vol[s];
    const float* __restrict__ msk = a.mask[s];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    // D <= 1024 (checked by the entry point): 32-bit index arithmetic
    for (unsigned i = blockIdx.x * kThreads + threadIdx.x; i < (unsigned)D3; i += gridDim.x * kThreads) {
        const unsigned row = i / (unsigned)D;
        const int z = (int)(i - row * (unsigned)D), x = (int)(row / (unsigned)D), y = (int)(row - (unsigned)x * D);
        const float m0 = msk ? __ldg(msk + i) : 1.0f;
        const bool px = x + 1 < D && (msk ? __ldg(msk + i + D2) : 1.0f) * m0 > 0.0f;
        const bool py = y + 1 < D && (msk ? __ldg(msk + i + D) : 1.0f) * m0 > 0.0f;
        const bool pz = z + 1 < D && (msk ? __ldg(msk + i + 1) : 1.0f) * m0 > 0.0f;
        if (!(px || py || pz)) continue;
        float tx = 0.f, ty = 0.f, tz = 0.f;
        for (int c = 0; c < a.channels; ++c) {
            const float* v = vol + c * D3 + i;
            const float v0 = __ldg(v);
            if (px) { const float d = __ldg(v + D2) - v0; tx = fmaf(d, d, tx); }
            if (py) { const float d = __ldg(v + D) - v0; ty = fmaf(d, d, ty); }
            if (pz) { const float d = __ldg(v + 1) - v0; tz = fmaf(d, d, tz); }
        }
        acc[0] += tx; acc[1] += ty; acc[2] += tz; acc[3] += px ? 1.0 : 0.0;
    }
    __shared__ double red[kThreads / 32][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double w = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = w;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) t += red[w][threadIdx.x];
        if (t != 0.0) atomicAdd(out + 4 * s + threadIdx.x, t);
    }
}

}  // namespace

// out (n_scales, 4) fp64 = [tx, ty, tz, cnt] per scale, ACCUMULATED (the caller zeroes it).
// vols->vol[s] = (channels, D, D, D) NCDHW as the reference stores them; masks == NULL or masks->vol[s] ==
// NULL means "all ones" (tv_regularization's volume_mask_cas=None default).
extern "C" int gens_tv_reduce(const gens_pyramid_t* vols, const gens_pyramid_t* masks, int channels, int n_blocks,
                              double* out, void* stream) {
    GENS_CHECK_ARG(vols && out && channels > 0 && n_blocks > 0);
    GENS_CHECK_ARG(vols->n_scales >= 0 && vols->n_scales <= GENS_MAX_SCALES);
    if (masks) GENS_CHECK_ARG(masks->n_scales == vols->n_scales);
    if (vols->n_scales == 0) return 0;
    TvArgs a;
    a.n = vols->n_scales;
    a.channels = channels;
    for (int s = 0; s < a.n; ++s) {
        GENS_CHECK_ARG(vols->vol[s] && vols->dim[s] > 0 && vols->dim[s] <= 1024);
        if (masks) GENS_CHECK_ARG(masks->dim[s] == vols->dim[s]);
        a.vol[s] = vols->vol[s];
        a.mask[s] = masks ? masks->vol[s] : nullptr;
        a.dim[s] = vols->dim[s];
    }
    tv_reduce_kernel<<<dim3(n_blocks, a.n), kThreads, 0, (cudaStream_t)stream>>>(a, out);
    return gens_launch_status();
}
