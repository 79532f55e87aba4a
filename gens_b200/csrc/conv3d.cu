// K13: 3x3x3 convolution (stride 1, zero padding 1) with few channels, and the InstanceNorm that follows it (sm_100a).
//
// Where it sits: the volume regulariser that consumes K1's output (reference models/modules/reg_network.py:105-166,
// SURVEY 8f-4).  Its two 256^3 layers (conv0 8 -> 8, out_layers[0] 8 -> 4) take 13.9 / 14.1 ms each in cuDNN's generic
// implicit_convolveNd_sgemm (4 TFLOP/s) and every InstanceNorm3d over 8 x 16.7 M values 30 ms in ATen's batch-norm
// kernels (8 thread blocks reduce one channel each) -- 98 of the 112 ms the whole network takes on one B200
// (gpurun_out/r2q_regnet.txt).  fp32 FMA work, not a GEMM worth reshaping for tensor cores (N = 4..16).
//
//   conv3d_k3_kernel<COUT>   one CTA = 4 x 8 x 32 output voxels (d, h, w), thread = one (h, w) column of 4 voxels x all
//                            COUT channels.  Per chunk of 8 input channels the 6 x 10 x 34 input tile and the chunk's
//                            weights are staged in shared memory; per (cin, kh, kw) a thread reads its 6 inputs along d
//                            once and feeds 3 (kd) x 4 (voxels) x COUT FMAs, outputs as f32x2 pairs on FFMA2 with
//                            the weight pair broadcast from shared memory.
//                            The input may be an x-slab: `below` / `above` are the neighbouring planes (NULL = zeros).
//                            Epilogue: optional bias, coalesced stores, and the per-channel sum / sum of squares of the
//                            tile reduced in the block and added to `stats` (2 x COUT doubles) for the norm.
//   norm_relu_kernel         y = relu((x - mean_c) * rstd_c) [+ skip], in place, mean / rstd from `stats`.
// HBM traffic per conv: input once (+ 2x halo re-reads through L2) + output once.
#include "common.cuh"
#include "f32x2.cuh"

namespace {

constexpr int kTD = 4, kTH = 8, kTW = 32;          // output tile
constexpr int kID = kTD + 2, kIH = kTH + 2, kIW = kTW + 2;
constexpr int kCinChunk = 8;
constexpr int kTileFloats = kCinChunk * kID * kIH * kIW;  // 16320

template <int COUT>
__global__ void __launch_bounds__(kTH * kTW)
conv3d_k3_kernel(const float* __restrict__ x, const float* __restrict__ below, const float* __restrict__ above,
                 const float* __restrict__ wpk, const float* __restrict__ bias, int cin, int D, int H, int W,
                 float* __restrict__ y, double* __restrict__ stats) {
    extern __shared__ __align__(16) float smem[];
    float* s_in = smem;                    // [8][6][10][34]
    float* s_w = smem + kTileFloats;       // [8][3 kh][3 kw][3 kd][COUT]
    __shared__ float s_red[2 * COUT];
    constexpr int NP = COUT / 2;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * kTW + tx;
    const int w0 = blockIdx.x * kTW, h0 = blockIdx.y * kTH, d0 = blockIdx.z * kTD;
    const long long plane = (long long)H * W, vol = plane * D;
    if (tid < 2 * COUT) s_red[tid] = 0.f;

    f32x2 acc[kTD][NP];
#pragma unroll
    for (int od = 0; od < kTD; ++od)
#pragma unroll
        for (int p = 0; p < NP; ++p) acc[od][p] = bias ? pk(__ldg(bias + 2 * p), __ldg(bias + 2 * p + 1)) : pk(0.f, 0.f);

    for (int c0 = 0; c0 < cin; c0 += kCinChunk) {
        __syncthreads();  // the previous chunk's readers are done
        // ---- stage the input tile (zero outside the volume; planes -1 / D from the neighbouring slabs)
        for (int i = tid; i < kTileFloats; i += kTH * kTW) {
            const int iw = i % kIW, ih = (i / kIW) % kIH, id = (i / (kIW * kIH)) % kID, c = i / (kIW * kIH * kID);
            const int gw = w0 + iw - 1, gh = h0 + ih - 1, gd = d0 + id - 1;
            float v = 0.f;
            if (gw >= 0 && gw < W && gh >= 0 && gh < H) {
                const long long at = (long long)gh * W + gw;
                if (gd >= 0 && gd < D) v = __ldg(x + (long long)(c0 + c) * vol + (long long)gd * plane + at);
                else if (gd == -1 && below) v = __ldg(below + (long long)(c0 + c) * plane + at);
                else if (gd == D && above) v = __ldg(above + (long long)(c0 + c) * plane + at);
            }
            s_in[i] = v;
        }
        for (int i = tid; i < kCinChunk * 27 * COUT; i += kTH * kTW) s_w[i] = __ldg(wpk + (long long)c0 * 27 * COUT + i);
        __syncthreads();
        // ---- 8 input channels x 9 (kh, kw) columns: 6 inputs along d feed 3 x 4 x COUT FMAs
#pragma unroll 1
        for (int c = 0; c < kCinChunk; ++c) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float* col = s_in + ((c * kID) * kIH + (ty + kh)) * kIW + tx + kw;
                    float v[kID];
#pragma unroll
                    for (int id = 0; id < kID; ++id) v[id] = col[id * kIH * kIW];
                    const float* wk = s_w + ((c * 3 + kh) * 3 + kw) * 3 * COUT;
#pragma unroll
                    for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
                        for (int p = 0; p < NP; ++p) {
                            const float2 w2 = *reinterpret_cast<const float2*>(wk + kd * COUT + 2 * p);
                            const f32x2 wp = pk(w2.x, w2.y);
#pragma unroll
                            for (int od = 0; od < kTD; ++od) acc[od][p] = fma2(wp, bc(v[od + kd]), acc[od][p]);
                        }
                    }
                }
            }
        }
    }
    // ---- stores + InstanceNorm moments
    const int gw = w0 + tx, gh = h0 + ty;
    const bool inside = gw < W && gh < H;
    float s1[COUT], s2[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) s1[co] = s2[co] = 0.f;
#pragma unroll
    for (int od = 0; od < kTD; ++od) {
        const int gd = d0 + od;
        if (!inside || gd >= D) continue;
        float* o = y + (long long)gd * plane + (long long)gh * W + gw;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float a = lo(acc[od][p]), b = hi(acc[od][p]);
            o[(long long)(2 * p) * vol] = a;
            o[(long long)(2 * p + 1) * vol] = b;
            s1[2 * p] += a; s2[2 * p] = fmaf(a, a, s2[2 * p]);
            s1[2 * p + 1] += b; s2[2 * p + 1] = fmaf(b, b, s2[2 * p + 1]);
        }
    }
    if (stats) {
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            float a = s1[co], b = s2[co];
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, m);
                b += __shfl_xor_sync(0xffffffffu, b, m);
            }
            if (tx == 0) {
                atomicAdd(&s_red[co], a);
                atomicAdd(&s_red[COUT + co], b);
            }
        }
        __syncthreads();
        if (tid < 2 * COUT) atomicAdd(stats + tid, (double)s_red[tid]);
    }
}

// y = relu((x - mean_c) * rstd_c) (+ skip), in place; stats = [sum_c | sumsq_c] over `count` values per channel
__global__ void __launch_bounds__(256)
norm_relu_kernel(float* __restrict__ x, const double* __restrict__ stats, int channels, long long per_channel,
                 double count, float eps, const float* __restrict__ skip) {
    const long long n4 = per_channel / 4;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (i >= n4) return;
    const double mean = stats[c] / count;
    double var = stats[channels + c] / count - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const float m = (float)mean, r = (float)(1.0 / sqrt(var + (double)eps));
    float4* p = reinterpret_cast<float4*>(x + (long long)c * per_channel) + i;
    float4 v = *p;
    v.x = fmaxf((v.x - m) * r, 0.f);
    v.y = fmaxf((v.y - m) * r, 0.f);
    v.z = fmaxf((v.z - m) * r, 0.f);
    v.w = fmaxf((v.w - m) * r, 0.f);
    if (skip) {
        const float4 s = __ldg(reinterpret_cast<const float4*>(skip + (long long)c * per_channel) + i);
        v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
    }
    *p = v;
}

template <int COUT>
int launch_conv(const float* x, const float* below, const float* above, const float* wpk, const float* bias, int cin, int D,
                int H, int W, float* y, double* stats, cudaStream_t st) {
    const int smem = (kTileFloats + kCinChunk * 27 * COUT) * (int)sizeof(float);
    const cudaError_t e = cudaFuncSetAttribute(conv3d_k3_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    const dim3 grid(ceil_div_i(W, kTW), ceil_div_i(H, kTH), ceil_div_i(D, kTD)), block(kTW, kTH);
    conv3d_k3_kernel<COUT><<<grid, block, smem, st>>>(x, below, above, wpk, bias, cin, D, H, W, y, stats);
    return gens_launch_status();
}

}  // namespace

extern "C" int gens_conv3d_k3(const float* x, const float* lo_plane, const float* hi_plane, const float* w_packed,
                              const float* bias, int c_in, int c_out, int d, int h, int w, float* y, double* stats,
                              void* stream) {
    GENS_CHECK_ARG(x && w_packed && y && d > 0 && h > 0 && w > 0 && c_in > 0);
    if (c_in % kCinChunk != 0) return GENS_E_UNSUPPORTED;
    const cudaStream_t st = (cudaStream_t)stream;
    switch (c_out) {
        case 4: return launch_conv<4>(x, lo_plane, hi_plane, w_packed, bias, c_in, d, h, w, y, stats, st);
        case 8: return launch_conv<8>(x, lo_plane, hi_plane, w_packed, bias, c_in, d, h, w, y, stats, st);
        case 16: return launch_conv<16>(x, lo_plane, hi_plane, w_packed, bias, c_in, d, h, w, y, stats, st);
        default: return GENS_E_UNSUPPORTED;
    }
}

extern "C" int gens_instnorm_relu(float* x, const double* stats, int channels, long long per_channel, double count,
                                  float eps, const float* skip, void* stream) {
    GENS_CHECK_ARG(x && stats && channels > 0 && per_channel > 0 && count > 0);
    if (per_channel % 4 != 0) return GENS_E_UNSUPPORTED;
    const dim3 grid(ceil_div_i(per_channel / 4, 256), channels);
    norm_relu_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, stats, channels, per_channel, count, eps, skip);
    return gens_launch_status();
}
