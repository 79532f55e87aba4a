// K13: 3x3x3 convolution (stride 1, zero padding 1) with few channels, and the InstanceNorm that follows it (sm_100a).
//
// Where it sits: the volume regulariser that consumes K1's output (reference models/modules/reg_network.py:105-166,
// SURVEY 8f-4).  Its two 256^3 layers (conv0 8 -> 8, out_layers[0] 8 -> 4) take 13.9 / 14.1 ms each in cuDNN's generic
// implicit_convolveNd_sgemm (4 TFLOP/s) and every InstanceNorm3d over 8 x 16.7 M values 30 ms in ATen's batch-norm
// kernels (8 thread blocks reduce one channel each) -- 98 of the 112 ms the whole network takes on one B200
// (gpurun_out/r2q_regnet.txt).  fp32 FMA work, not a GEMM worth reshaping for tensor cores (N = 4..16).
//
//   conv3d_k3_kernel<COUT>   one CTA = 8 x 8 x 32 output voxels (d, h, w), thread = one (h, w) column of 8 voxels x all
//                            COUT channels.  Per chunk of 4 input channels the 10 x 10 x 34 input tile and the chunk's
//                            weights are staged in shared memory; per (cin, kh, kw) a thread reads its 10 inputs along d
//                            once and feeds 3 (kd) x 8 (voxels) x COUT FMAs, outputs as f32x2 pairs on FFMA2 with
//                            two weight pairs per broadcast LDS.128 (a 4-voxel column with LDS.64 weights kept the
//                            shared-memory pipe ~75 % busy: 2.11 ms for 8 -> 8 channels at 256^3).
//                            The input may be an x-slab: `below` / `above` are the neighbouring planes (NULL = zeros).
//                            Epilogue: optional bias, coalesced stores, and the per-channel sum / sum of squares of the
//                            tile reduced in the block and added to `stats` (2 x COUT doubles) for the norm.
//   conv3d_k3s2_kernel       the stride-2 layers, deconv3d_k3s2_kernel the transposed stride-2 layers (see there).
//   norm_relu_kernel         y = relu((x - mean_c) * rstd_c) [+ skip], in place, mean / rstd from `stats`.
// HBM traffic per conv: input once (+ 2x halo re-reads through L2) + output once.
#include "common.cuh"
#include "f32x2.cuh"

namespace {

constexpr int kTH = 8, kTW = 32;                   // output tile in h, w; kTD (template) planes in d
constexpr int kIH = kTH + 2, kIW = kTW + 2;
constexpr int kCinChunk = 4;

template <int COUT, int kTD>
__global__ void __launch_bounds__(kTH * kTW)
conv3d_k3_kernel(const float* __restrict__ x, const float* __restrict__ below, const float* __restrict__ above,
                 const float* __restrict__ wpk, const float* __restrict__ bias, int cin, int D, int H, int W,
                 float* __restrict__ y, double* __restrict__ stats) {
    constexpr int kID = kTD + 2, kTileFloats = kCinChunk * kID * kIH * kIW;
    extern __shared__ __align__(16) float smem[];
    float* s_in = smem;                    // [4][kTD + 2][10][34]
    float* s_w = smem + kTileFloats;       // [4][3 kh][3 kw][3 kd][COUT]
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
        // (per-element index decomposition; a warp-per-row variant with warp-uniform addresses was slower, 2.66 vs
        // 2.11 ms at 8 -> 8 channels and 256^3: fewer independent loads in flight per thread)
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
        // ---- 4 input channels x 9 (kh, kw) columns: 10 inputs along d feed 3 x 8 x COUT FMAs
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
                        for (int q = 0; q < COUT / 4; ++q) {  // one LDS.128 = two weight pairs, each feeding kTD FFMA2
                            const float4 w4 = *reinterpret_cast<const float4*>(wk + kd * COUT + 4 * q);
                            const f32x2 wa = pk(w4.x, w4.y), wb = pk(w4.z, w4.w);
#pragma unroll
                            for (int od = 0; od < kTD; ++od) {
                                acc[od][2 * q] = fma2(wa, bc(v[od + kd]), acc[od][2 * q]);
                                acc[od][2 * q + 1] = fma2(wb, bc(v[od + kd]), acc[od][2 * q + 1]);
                            }
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

// Stride-2 variant (the first layer of every encoder stage): out (D/2, H/2, W/2), out voxel o <- in 2o-1 .. 2o+1.
// 1/8 of the stride-1 layer's outputs and FMAs; bound by the input read.  One thread = one output voxel x all COUT
// channels, inputs straight from global memory (neighbouring threads share lines through L1), weights of all input
// channels in shared memory.  `below` = plane -1 of an x-slab (NULL = zeros); D, H, W (input) even.
template <int COUT>
__global__ void __launch_bounds__(256)
conv3d_k3s2_kernel(const float* __restrict__ x, const float* __restrict__ below, const float* __restrict__ wpk, int cin,
                   int D, int H, int W, float* __restrict__ y, double* __restrict__ stats) {
    extern __shared__ __align__(16) float s_w[];  // [cin][kh][kw][kd][COUT]
    __shared__ float s_red[2 * COUT];
    constexpr int NP = COUT / 2;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < cin * 27 * COUT; i += 256) s_w[i] = __ldg(wpk + i);
    if (tid < 2 * COUT) s_red[tid] = 0.f;
    __syncthreads();
    const int Do = D / 2, Ho = H / 2, Wo = W / 2;
    const int ow = blockIdx.x * 32 + threadIdx.x, oh = blockIdx.y * 8 + threadIdx.y, od = blockIdx.z;
    const bool inside = ow < Wo && oh < Ho;
    const long long plane = (long long)H * W, vol = plane * D;
    f32x2 acc[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] = pk(0.f, 0.f);
    if (inside) {
#pragma unroll 1
        for (int c = 0; c < cin; ++c) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int gh = 2 * oh - 1 + kh;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int gw = 2 * ow - 1 + kw;
                    const bool ok = gh >= 0 && gw >= 0;  // the upper ends never leave an even-sized input
                    const float* wk = s_w + ((c * 3 + kh) * 3 + kw) * 3 * COUT;
#pragma unroll
                    for (int kd = 0; kd < 3; ++kd) {
                        const int gd = 2 * od - 1 + kd;
                        float v = 0.f;
                        if (ok) {
                            const long long at = (long long)gh * W + gw;
                            if (gd >= 0) v = __ldg(x + (long long)c * vol + (long long)gd * plane + at);
                            else if (below) v = __ldg(below + (long long)c * plane + at);
                        }
#pragma unroll
                        for (int p = 0; p < NP; ++p) {
                            const float2 w2 = *reinterpret_cast<const float2*>(wk + kd * COUT + 2 * p);
                            acc[p] = fma2(pk(w2.x, w2.y), bc(v), acc[p]);
                        }
                    }
                }
            }
        }
        const long long oplane = (long long)Ho * Wo;
        float* o = y + (long long)od * oplane + (long long)oh * Wo + ow;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            o[(long long)(2 * p) * oplane * Do] = lo(acc[p]);
            o[(long long)(2 * p + 1) * oplane * Do] = hi(acc[p]);
        }
    }
    if (stats) {
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            const float val = inside ? ((co & 1) ? hi(acc[co / 2]) : lo(acc[co / 2])) : 0.f;
            float a = val, b = val * val;
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, m);
                b += __shfl_xor_sync(0xffffffffu, b, m);
            }
            if (threadIdx.x == 0) {
                atomicAdd(&s_red[co], a);
                atomicAdd(&s_red[COUT + co], b);
            }
        }
        __syncthreads();
        if (tid < 2 * COUT) atomicAdd(stats + tid, (double)s_red[tid]);
    }
}

// Transposed stride-2 variant (every decoder stage: ConvTranspose3d(k 3, stride 2, padding 1, output_padding 1)):
// out (2D, 2H, 2W), out voxel o <- in i with o = 2i - 1 + k, i.e. per axis an even output 2i takes (i, k = 1), an odd
// output 2i+1 takes (i, k = 2) and (i+1, k = 0).  One thread = one INPUT voxel: it reads its 2 x 2 x 2 neighbourhood and
// produces the 2 x 2 x 2 outputs of its cell x all COUT channels -- 27 taps in total instead of 8 x 27, no zero-stuffed
// intermediate (cuDNN runs this layer as a data-gradient convolution after a 0.28 ms fill: 3.2 ms at 128^3 -> 256^3).
// `above` = plane D of an x-slab (NULL = zeros).  wpk = [cin][kd][kh][kw][COUT].
template <int COUT>
__global__ void __launch_bounds__(256, 2)
deconv3d_k3s2_kernel(const float* __restrict__ x, const float* __restrict__ above, const float* __restrict__ wpk, int cin,
                     int D, int H, int W, float* __restrict__ y, double* __restrict__ stats) {
    extern __shared__ __align__(16) float s_w[];  // [cin][kd][kh][kw][COUT]
    __shared__ float s_red[2 * COUT];
    constexpr int NP = COUT / 2;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < cin * 27 * COUT; i += 256) s_w[i] = __ldg(wpk + i);
    if (tid < 2 * COUT) s_red[tid] = 0.f;
    __syncthreads();
    const int iw = blockIdx.x * 32 + threadIdx.x, ih = blockIdx.y * 8 + threadIdx.y, id = blockIdx.z;
    const bool inside = iw < W && ih < H;
    const long long plane = (long long)H * W, vol = plane * D;
    f32x2 acc[8][NP];  // [parity: 4 a_d + 2 a_h + a_w]
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int p = 0; p < NP; ++p) acc[a][p] = pk(0.f, 0.f);
    if (inside) {
#pragma unroll 1
        for (int c = 0; c < cin; ++c) {
            float v[8];  // neighbourhood [4 dd + 2 dh + dw]
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int dd = n >> 2, dh = (n >> 1) & 1, dw = n & 1;
                const int gd = id + dd, gh = ih + dh, gw = iw + dw;
                float t = 0.f;
                if (gh < H && gw < W) {
                    const long long at = (long long)gh * W + gw;
                    if (gd < D) t = __ldg(x + (long long)c * vol + (long long)gd * plane + at);
                    else if (above) t = __ldg(above + (long long)c * plane + at);
                }
                v[n] = t;
            }
            const float* wc = s_w + c * 27 * COUT;
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int ad = a >> 2, ah = (a >> 1) & 1, aw = a & 1;
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const int dd = n >> 2, dh = (n >> 1) & 1, dw = n & 1;
                    if (dd > ad || dh > ah || dw > aw) continue;  // compile-time: 27 (a, n) pairs survive
                    const int kd = ad ? (dd ? 0 : 2) : 1, kh = ah ? (dh ? 0 : 2) : 1, kw = aw ? (dw ? 0 : 2) : 1;
                    const float* wk = wc + ((kd * 3 + kh) * 3 + kw) * COUT;
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        const float2 w2 = *reinterpret_cast<const float2*>(wk + 2 * p);
                        acc[a][p] = fma2(pk(w2.x, w2.y), bc(v[n]), acc[a][p]);
                    }
                }
            }
        }
        const long long orow = 2LL * W, oplane = 4LL * plane, ovol = oplane * (2LL * D);
#pragma unroll
        for (int a = 0; a < 8; a += 2) {  // the two w-parities of one (a_d, a_h) leave as one 8-byte store per channel
            const int ad = a >> 2, ah = (a >> 1) & 1;
            float* o = y + (long long)(2 * id + ad) * oplane + (long long)(2 * ih + ah) * orow + 2 * iw;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                *reinterpret_cast<float2*>(o + (long long)(2 * p) * ovol) = make_float2(lo(acc[a][p]), lo(acc[a + 1][p]));
                *reinterpret_cast<float2*>(o + (long long)(2 * p + 1) * ovol) = make_float2(hi(acc[a][p]), hi(acc[a + 1][p]));
            }
        }
    }
    if (stats) {
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            float a = 0.f, b = 0.f;
            if (inside) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float val = (co & 1) ? hi(acc[q][co / 2]) : lo(acc[q][co / 2]);
                    a += val;
                    b = fmaf(val, val, b);
                }
            }
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, m);
                b += __shfl_xor_sync(0xffffffffu, b, m);
            }
            if (threadIdx.x == 0) {
                atomicAdd(&s_red[co], a);
                atomicAdd(&s_red[COUT + co], b);
            }
        }
        __syncthreads();
        if (tid < 2 * COUT) atomicAdd(stats + tid, (double)s_red[tid]);
    }
}

template <int COUT, int kTD>
int launch_conv(const float* x, const float* below, const float* above, const float* wpk, const float* bias, int cin, int D,
                int H, int W, float* y, double* stats, cudaStream_t st) {
    const int smem = (kCinChunk * (kTD + 2) * kIH * kIW + kCinChunk * 27 * COUT) * (int)sizeof(float);
    const cudaError_t e = cudaFuncSetAttribute(conv3d_k3_kernel<COUT, kTD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    const dim3 grid(ceil_div_i(W, kTW), ceil_div_i(H, kTH), ceil_div_i(D, kTD)), block(kTW, kTH);
    conv3d_k3_kernel<COUT, kTD><<<grid, block, smem, st>>>(x, below, above, wpk, bias, cin, D, H, W, y, stats);
    return gens_launch_status();
}

}  // namespace

namespace {
int g_conv_td8 = 0;
}
// measurement knob (tools/prof_regnet.py): 8-voxel columns for the 8-output-channel kernel as well
extern "C" int gens_debug_conv_td8(int on) {
    g_conv_td8 = on ? 1 : 0;
    return 0;
}

extern "C" int gens_conv3d_k3(const float* x, const float* lo_plane, const float* hi_plane, const float* w_packed,
                              const float* bias, int c_in, int c_out, int d, int h, int w, float* y, double* stats,
                              void* stream) {
    GENS_CHECK_ARG(x && w_packed && y && d > 0 && h > 0 && w > 0 && c_in > 0);
    if (c_in % kCinChunk != 0) return GENS_E_UNSUPPORTED;
    const cudaStream_t st = (cudaStream_t)stream;
    switch (c_out) {
        // column height per thread: 8 voxels at 4 output channels (64 registers), 4 at 8 / 16 (8 voxels x 8 channels
        // needs 118 registers and was slower: 2.50 vs 2.11 ms at 256^3)
        case 4: return launch_conv<4, 8>(x, lo_plane, hi_plane, w_packed, bias, c_in, d, h, w, y, stats, st);
        case 8: return g_conv_td8 ? launch_conv<8, 8>(x, lo_plane, hi_plane, w_packed, bias, c_in, d, h, w, y, stats, st)
                                  : launch_conv<8, 4>(x, lo_plane, hi_plane, w_packed, bias, c_in, d, h, w, y, stats, st);
        case 16: return launch_conv<16, 4>(x, lo_plane, hi_plane, w_packed, bias, c_in, d, h, w, y, stats, st);
        default: return GENS_E_UNSUPPORTED;
    }
}

namespace {
template <int COUT>
int launch_strided(bool transposed, const float* x, const float* halo, const float* wpk, int cin, int D, int H, int W,
                   float* y, double* stats, cudaStream_t st) {
    const int smem = cin * 27 * COUT * (int)sizeof(float);
    if (smem > 96 * 1024) return GENS_E_UNSUPPORTED;
    auto kern = transposed ? deconv3d_k3s2_kernel<COUT> : conv3d_k3s2_kernel<COUT>;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    const dim3 block(32, 8);
    const dim3 grid = transposed ? dim3(ceil_div_i(W, 32), ceil_div_i(H, 8), D) : dim3(ceil_div_i(W / 2, 32), ceil_div_i(H / 2, 8), D / 2);
    kern<<<grid, block, smem, st>>>(x, halo, wpk, cin, D, H, W, y, stats);
    return gens_launch_status();
}
}  // namespace

extern "C" int gens_conv3d_k3s2(const float* x, const float* lo_plane, const float* w_packed, int c_in, int c_out, int d,
                                int h, int w, float* y, double* stats, void* stream) {
    GENS_CHECK_ARG(x && w_packed && y && d > 0 && h > 0 && w > 0 && c_in > 0);
    if (d % 2 || h % 2 || w % 2) return GENS_E_UNSUPPORTED;
    const cudaStream_t st = (cudaStream_t)stream;
    switch (c_out) {
        case 8: return launch_strided<8>(false, x, lo_plane, w_packed, c_in, d, h, w, y, stats, st);
        case 16: return launch_strided<16>(false, x, lo_plane, w_packed, c_in, d, h, w, y, stats, st);
        default: return GENS_E_UNSUPPORTED;
    }
}

extern "C" int gens_deconv3d_k3s2(const float* x, const float* hi_plane, const float* w_packed, int c_in, int c_out,
                                  int d, int h, int w, float* y, double* stats, void* stream) {
    GENS_CHECK_ARG(x && w_packed && y && d > 0 && h > 0 && w > 0 && c_in > 0);
    const cudaStream_t st = (cudaStream_t)stream;
    switch (c_out) {
        case 8: return launch_strided<8>(true, x, hi_plane, w_packed, c_in, d, h, w, y, stats, st);
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
