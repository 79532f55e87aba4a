// Fused element-wise stages of the analytic SDF pass (value, gradient and the second-order
// "smooth" term in one sweep, no autograd graph).
//
// Replaces what the reference obtains with two nested torch.autograd.grad(create_graph=True) calls
// through the SDF MLP (reference models/modules/sdf_network.py:131-153): ~9k ATen launches and a
// double-backward graph per render() call.  Here the MLP is differentiated by hand, forward-over-
// reverse, along the fixed direction u = (1,1,1) the reference uses (d_output2 = ones):
//     rows [0,n)  of every work matrix carry the primal quantity,
//     rows [n,2n) its directional derivative along u ("tangent"),
// so each layer is ONE cuBLAS SGEMM over 2n rows (plain library GEMM, fp32) in each direction, and
// the kernels below do everything between the GEMMs:
//   encode      p, feats, dfeats -> positional encodings + tangents          (embedder.py:11-36)
//   act_fwd     a = y + featpart + b ; h = softplus(a) ; dh = sp'(a) da ; keeps sp'(a), sp''(a) da
//   act_bwd     g_a = sp' g_h ;  dg_a = sp'' da g_h + sp' dg_h
//   decode      cotangents of the encodings -> cotangents of p (PE part) and of the volume features
// softplus is torch.nn.Softplus(beta=100, threshold=20) as the reference builds it (sdf_network.py:95).
#include "common.cuh"

namespace {

__device__ __forceinline__ void softplus3(float a, float beta, float& sp, float& d1, float& d2) {
    const float t = a * beta;
    if (t > 20.0f) {  // torch's linear region
        sp = a; d1 = 1.0f; d2 = 0.0f;
        return;
    }
    const float z = expf(t);
    sp = log1pf(z) / beta;
    d1 = z / (1.0f + z);
    d2 = beta * d1 * (1.0f - d1);
}

// enc (2n, d*(1+2L)): [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)], tangent rows below
__device__ __forceinline__ void encode_one(float x, float dx, int L, int d, int i, float* __restrict__ row,
                                           float* __restrict__ drow) {
    row[i] = x;
    if (drow) drow[i] = dx;
    float f = 1.0f;
    for (int k = 0; k < L; ++k) {
        float s, c;
        sincosf(x * f, &s, &c);
        row[(1 + 2 * k) * d + i] = s;
        row[(2 + 2 * k) * d + i] = c;
        if (drow) {
            drow[(1 + 2 * k) * d + i] = f * c * dx;
            drow[(2 + 2 * k) * d + i] = -f * s * dx;
        }
        f *= 2.0f;
    }
}

__global__ void __launch_bounds__(256)
encode_kernel(const float* __restrict__ pts, const float* __restrict__ feats, const float* __restrict__ dfeats,
              long long n, float scale, float u0, float u1, float u2, int L_pos, int L_feat, int n_feat,
              float* __restrict__ pos, float* __restrict__ fe) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = 3 + n_feat;
    if (i >= n * per) return;
    const long long p = i / per;
    const int j = (int)(i % per);
    if (j < 3) {
        const int w = 3 * (1 + 2 * L_pos);
        const float u = j == 0 ? u0 : (j == 1 ? u1 : u2);
        encode_one(pts[3 * p + j] * scale, u * scale, L_pos, 3, j, pos + p * w, dfeats ? pos + (n + p) * w : nullptr);
    } else {
        const int c = j - 3, w = n_feat * (1 + 2 * L_feat);
        encode_one(feats[p * n_feat + c], dfeats ? dfeats[p * n_feat + c] : 0.f, L_feat, n_feat, c, fe + p * w,
                   dfeats ? fe + (n + p) * w : nullptr);
    }
}

// y (2n,fo) GEMM result; fp = feature-part slice (2n rows, leading dim ldfp) or null; x_out (2n rows, ld ldx)
__global__ void __launch_bounds__(256)
act_fwd_kernel(const float* __restrict__ y, const float* __restrict__ fp, int ldfp, const float* __restrict__ bias,
               long long n, int fo, float beta, float out_scale, float* __restrict__ x_out, int ldx,
               float* __restrict__ s1, float* __restrict__ t2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * fo) return;
    const long long r = i / fo;
    const int c = (int)(i % fo);
    float a = y[r * fo + c] + bias[c];
    if (fp) a += fp[r * ldfp + c];
    float sp, d1, d2;
    softplus3(a, beta, sp, d1, d2);
    x_out[r * ldx + c] = sp * out_scale;
    if (s1) {  // tangent rows present
        float da = y[(n + r) * fo + c];
        if (fp) da += fp[(n + r) * ldfp + c];
        x_out[(n + r) * ldx + c] = d1 * da * out_scale;
        s1[i] = d1;
        t2[i] = d2 * da;
    }
}

// copy a (2n, w) block scaled into columns [col, col+w) of x_out (the skip connection [h, pos]/sqrt 2)
__global__ void __launch_bounds__(256)
copy_scaled_kernel(const float* __restrict__ src, int w, long long rows, float s, float* __restrict__ dst, int ld,
                   int col) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * w) return;
    dst[(i / w) * ld + col + (int)(i % w)] = src[i] * s;
}

// g (2n rows, ld ldg, scaled by in_scale) = [g_h; dg_h]  ->  ga (2n rows, ld ldga) = [g_a; dg_a]
__global__ void __launch_bounds__(256)
act_bwd_kernel(const float* __restrict__ g, int ldg, float in_scale, const float* __restrict__ s1,
               const float* __restrict__ t2, long long n, int fo, float* __restrict__ ga, int ldga) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * fo) return;
    const long long r = i / fo;
    const int c = (int)(i % fo);
    const float gh = g[r * ldg + c] * in_scale, dgh = g[(n + r) * ldg + c] * in_scale;
    const float d1 = s1[i];
    ga[r * ldga + c] = d1 * gh;
    ga[(n + r) * ldga + c] = t2[i] * gh + d1 * dgh;
}

// cotangent of one encoded block back to its argument: returns (g, dg) for x given cotangents of enc
__device__ __forceinline__ void decode_one(float x, float dx, int L, int d, int i, const float* __restrict__ grow,
                                           const float* __restrict__ dgrow, float& g, float& dg) {
    g = grow[i];
    dg = dgrow[i];
    float f = 1.0f;
    for (int k = 0; k < L; ++k) {
        float s, c;
        sincosf(x * f, &s, &c);
        const float gs = grow[(1 + 2 * k) * d + i], gc = grow[(2 + 2 * k) * d + i];
        const float dgs = dgrow[(1 + 2 * k) * d + i], dgc = dgrow[(2 + 2 * k) * d + i];
        g += f * (c * gs - s * gc);
        // d/de [ f cos(fx) gs - f sin(fx) gc ] = -f^2 dx (sin gs + cos gc) + f (cos dgs - sin dgc)
        dg += -f * f * dx * (s * gs + c * gc) + f * (c * dgs - s * dgc);
        f *= 2.0f;
    }
}

__global__ void __launch_bounds__(256)
decode_kernel(const float* __restrict__ pts, const float* __restrict__ feats, const float* __restrict__ dfeats,
              const float* __restrict__ g_pos, const float* __restrict__ g_fe, long long n, float scale, float u0,
              float u1, float u2, int L_pos, int L_feat, int n_feat, float* __restrict__ g_f,
              float* __restrict__ dg_f, float* __restrict__ grad, float* __restrict__ smooth) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = 3 + n_feat;
    if (i >= n * per) return;
    const long long p = i / per;
    const int j = (int)(i % per);
    float g, dg;
    if (j < 3) {
        const int w = 3 * (1 + 2 * L_pos);
        const float u = j == 0 ? u0 : (j == 1 ? u1 : u2);
        decode_one(pts[3 * p + j] * scale, u * scale, L_pos, 3, j, g_pos + p * w, g_pos + (n + p) * w, g, dg);
        grad[3 * p + j] = g * scale;      // chain through x = p * scale
        smooth[3 * p + j] = dg * scale;
    } else {
        const int c = j - 3, w = n_feat * (1 + 2 * L_feat);
        decode_one(feats[p * n_feat + c], dfeats[p * n_feat + c], L_feat, n_feat, c, g_fe + p * w, g_fe + (n + p) * w,
                   g, dg);
        g_f[p * n_feat + c] = g;
        dg_f[p * n_feat + c] = dg;
    }
}

}  // namespace

extern "C" int gens_sdf_encode(const float* pts, const float* feats, const float* dfeats, long long n, float scale,
                               const float* u3, int multires, int feat_multires, int n_feat, float* pos, float* fe,
                               void* stream) {
    if (n == 0) return 0;
    GENS_CHECK_ARG(pts && feats && u3 && pos && fe && n >= 0 && n_feat > 0);  // dfeats NULL: value only (n rows)
    if (n == 0) return 0;
    const long long total = n * (3 + n_feat);
    encode_kernel<<<ceil_div_i(total, 256), 256, 0, (cudaStream_t)stream>>>(pts, feats, dfeats, n, scale, u3[0], u3[1],
                                                                          u3[2], multires, feat_multires, n_feat, pos, fe);
    return gens_launch_status();
}

extern "C" int gens_sdf_act_fwd(const float* y, const float* featpart, int ld_featpart, const float* bias, long long n,
                                int fan_out, float beta, float out_scale, float* x_out, int ld_x, float* sp1, float* sp2da,
                                void* stream) {
    if (n == 0) return 0;
    GENS_CHECK_ARG(y && bias && x_out && n >= 0 && fan_out > 0 && ld_x >= fan_out && (!sp1 == !sp2da));
    if (n == 0) return 0;
    act_fwd_kernel<<<ceil_div_i(n * fan_out, 256), 256, 0, (cudaStream_t)stream>>>(
        y, featpart, ld_featpart, bias, n, fan_out, beta, out_scale, x_out, ld_x, sp1, sp2da);
    return gens_launch_status();
}

extern "C" int gens_copy_scaled(const float* src, int width, long long rows, float s, float* dst, int ld_dst, int col,
                                void* stream) {
    GENS_CHECK_ARG(src && dst && width > 0 && rows >= 0 && ld_dst >= col + width);
    if (rows == 0) return 0;
    copy_scaled_kernel<<<ceil_div_i(rows * width, 256), 256, 0, (cudaStream_t)stream>>>(src, width, rows, s, dst, ld_dst, col);
    return gens_launch_status();
}

extern "C" int gens_sdf_act_bwd(const float* g, int ld_g, float in_scale, const float* sp1, const float* sp2da,
                                long long n, int fan_out, float* ga, int ld_ga, void* stream) {
    if (n == 0) return 0;
    GENS_CHECK_ARG(g && sp1 && sp2da && ga && n >= 0 && fan_out > 0 && ld_g >= fan_out && ld_ga >= fan_out);
    if (n == 0) return 0;
    act_bwd_kernel<<<ceil_div_i(n * fan_out, 256), 256, 0, (cudaStream_t)stream>>>(g, ld_g, in_scale, sp1, sp2da, n,
                                                                                  fan_out, ga, ld_ga);
    return gens_launch_status();
}

extern "C" int gens_sdf_decode(const float* pts, const float* feats, const float* dfeats, const float* g_pos,
                               const float* g_fe, long long n, float scale, const float* u3, int multires,
                               int feat_multires, int n_feat, float* g_f, float* dg_f, float* grad, float* smooth,
                               void* stream) {
    if (n == 0) return 0;
    GENS_CHECK_ARG(pts && feats && dfeats && g_pos && g_fe && u3 && g_f && dg_f && grad && smooth && n >= 0);
    if (n == 0) return 0;
    const long long total = n * (3 + n_feat);
    decode_kernel<<<ceil_div_i(total, 256), 256, 0, (cudaStream_t)stream>>>(
        pts, feats, dfeats, g_pos, g_fe, n, scale, u3[0], u3[1], u3[2], multires, feat_multires, n_feat, g_f, dg_f, grad,
        smooth);
    return gens_launch_status();
}
