// K7: NeuS alpha compositing of one batch of rays, one warp per ray (sm_100a).
//
// Replaces, for inference (torch.no_grad), the ~100 ATen launches between the network evaluations and the
// output dictionary of ImplicitSurface.render_core (reference models/modules/implicit_surface.py:179-326):
// masking of the evaluated samples (:179-190, :199-200), cosine annealing and the section alphas (:206-226),
// the transmittance cumprod and weights (:235-236), every weighted ray sum (colour, normal, depth: :238-247),
// the eikonal / smoothness partial sums (:249-257), the visibility count behind valid_mask (:202-203) and the
// first SDF zero crossing with its interpolated depth (:262-300).  A lane owns samples lane, lane+32, ...
// (coalesced loads); the transmittance is a warp-shuffle prefix product per 32-sample row with the row total
// carried to the next row; nothing is materialised between the stages.
// Bound: HBM, ~75 bytes read + ~24 bytes written per ray sample.
#include "common.cuh"

namespace {

constexpr int kMaxSamples = 160;
constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock) composite_kernel(gens_composite_args_t a) {
    __shared__ float s_z[kWarpsPerBlock][kMaxSamples + 1], s_sdf[kWarpsPerBlock][kMaxSamples + 1];
    __shared__ unsigned char s_vm[kWarpsPerBlock][kMaxSamples + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ray = blockIdx.x * kWarpsPerBlock + warp;
    if (ray >= a.n_rays) return;
    const int N = a.n_samples;
    const long long base = (long long)ray * N;
    float *z = s_z[warp], *sd = s_sdf[warp];
    unsigned char* vmask = s_vm[warp];

    // stage depths, masked SDF and voxel validity (neighbour access for the sections and the crossing search)
    for (int j = lane; j < N; j += 32) {
        z[j] = a.z_vals[base + j];
        const bool ev = a.evaluated[base + j] != 0;
        const float s = ev ? a.sdf_raw[base + j] : 100.0f;
        sd[j] = s;
        vmask[j] = a.voxel_mask[base + j] != 0;
        a.sdf_out[base + j] = s;
    }
    __syncwarp();

    const float ox = a.rays_o[3 * ray], oy = a.rays_o[3 * ray + 1], oz = a.rays_o[3 * ray + 2];
    const float dx = a.rays_d[3 * ray], dy = a.rays_d[3 * ray + 1], dz = a.rays_d[3 * ray + 2];
    const float inv_s = fminf(fmaxf(*a.inv_s, 1e-6f), 1e6f);
    const float car = a.cos_anneal_ratio;

    float T_carry = 1.0f;
    float w_sum = 0.f, w_max = 0.f, col[3] = {0.f, 0.f, 0.f}, nrm[3] = {0.f, 0.f, 0.f}, depth = 0.f;
    float sm[3] = {0.f, 0.f, 0.f}, ge_num = 0.f, ge_den = 0.f;
    int n_visible = 0, first_cross = N;  // N = "none"
    for (int r = 0; r * 32 < N; ++r) {
        const int j = r * 32 + lane;
        const bool live = j < N;
        float alpha = 0.f, mid = 0.f, inside = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
        float cr = 0.f, cg = 0.f, cb = 0.f, hx = 0.f, hy = 0.f, hz = 0.f;
        if (live) {
            const long long i = base + j;
            const float dist = j + 1 < N ? z[j + 1] - z[j] : a.sample_dist;
            mid = z[j] + dist * 0.5f;
            const bool ev = a.evaluated[i] != 0;
            const float vm = vmask[j] ? 1.0f : 0.0f;
            if (ev) {
                gx = a.grad_raw[3 * i]; gy = a.grad_raw[3 * i + 1]; gz = a.grad_raw[3 * i + 2];
                hx = a.smooth_raw[3 * i]; hy = a.smooth_raw[3 * i + 1]; hz = a.smooth_raw[3 * i + 2];
                cr = a.colour_raw[3 * i]; cg = a.colour_raw[3 * i + 1]; cb = a.colour_raw[3 * i + 2];
            }
            a.gradients_out[3 * i] = gx; a.gradients_out[3 * i + 1] = gy; a.gradients_out[3 * i + 2] = gz;
            const float px = a.pts[3 * i], py = a.pts[3 * i + 1], pz = a.pts[3 * i + 2];
            const float pn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
            inside = pn < 1.0f ? vm : 0.0f;
            const float relax = pn < 1.2f ? vm : 0.0f;
            a.inside_out[i] = inside;
            const float true_cos = dx * gx + dy * gy + dz * gz;
            float iter_cos = -(fmaxf(-true_cos * 0.5f + 0.5f, 0.f) * (1.0f - car) + fmaxf(-true_cos, 0.f) * car);
            iter_cos *= vm;
            const float half = fminf(fmaxf(iter_cos, -10.f), 10.f) * dist * 0.5f;
            const float s = sd[j];
            const float prev_cdf = sigmoidf_((s - half) * inv_s), next_cdf = sigmoidf_((s + half) * inv_s);
            alpha = fminf(fmaxf((prev_cdf - next_cdf + 1e-5f) / (prev_cdf + 1e-5f), 0.f), 1.f) * vm;
            const float gn = sqrtf(gx * gx + gy * gy + gz * gz) - 1.0f;
            ge_num += relax * gn * gn;
            ge_den += relax;
            int vis = 0;
            for (int v = 0; v < a.n_src; ++v) vis += a.mask_views[i * a.n_src + v] != 0;
            n_visible += vis > 1;
            if (j + 1 < N && s * sd[j + 1] <= 0.f && vmask[j] && vmask[j + 1]) first_cross = min(first_cross, j);
        }
        // exclusive prefix product of (1 - alpha + 1e-7) over this row, times the carry of the rows before
        const float om = live ? 1.0f - alpha + 1e-7f : 1.0f;
        float x = om;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x *= y;
        }
        const float row_total = __shfl_sync(0xffffffffu, x, 31);
        float excl = __shfl_up_sync(0xffffffffu, x, 1);
        excl = lane == 0 ? 1.0f : excl;
        const float w = alpha * (T_carry * excl);
        T_carry *= row_total;
        if (live) {
            a.weights_out[base + j] = w;
            w_sum += w;
            w_max = fmaxf(w_max, w);
            col[0] += cr * w; col[1] += cg * w; col[2] += cb * w;
            nrm[0] += gx * w; nrm[1] += gy * w; nrm[2] += gz * w;
            depth += mid * w;
            const float wi = w * inside;
            sm[0] += hx * wi; sm[1] += hy * wi; sm[2] += hz * wi;
        }
    }
    w_sum = warp_sum(w_sum);
    w_max = warp_max(w_max);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        col[k] = warp_sum(col[k]);
        nrm[k] = warp_sum(nrm[k]);
        sm[k] = warp_sum(sm[k]);
    }
    depth = warp_sum(depth);
    ge_num = warp_sum(ge_num);
    ge_den = warp_sum(ge_den);
    n_visible = (int)(warp_sum((float)n_visible) + 0.5f);
    first_cross = warp_min_i(first_cross);

    if (lane == 0) {
        const float* R = a.rot;  // inverse(c2w[0,:3,:3]), row-major
        const float cam_dz = R[6] * dx + R[7] * dy + R[8] * dz;
        a.weight_sum_out[ray] = w_sum;
        a.weight_max_out[ray] = w_max;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            a.color_out[3 * ray + k] = col[k];
            a.normal_out[3 * ray + k] = R[3 * k] * nrm[0] + R[3 * k + 1] * nrm[1] + R[3 * k + 2] * nrm[2];
        }
        a.depth_out[ray] = depth * cam_dz;
        a.valid_out[ray] = n_visible > 8;
        a.ge_num_out[ray] = ge_num;
        a.ge_den_out[ray] = ge_den;
        a.smooth_norm_out[ray] = sqrtf(sm[0] * sm[0] + sm[1] * sm[1] + sm[2] * sm[2]);

        // first zero crossing (argmax of crossing * (N-1-j) * pair_valid: earliest; index 0 when none)
        const bool has = first_cross < N;
        const int j0 = has ? first_cross : 0, j1 = j0 + 1;
        float g0[3], g1[3], in0, in1, mid0, mid1;
        const int jj[2] = {j0, j1};
        float* gs[2] = {g0, g1};
        float ins[2], mids[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int j = jj[t];
            const long long i = base + j;
            const bool ev = a.evaluated[i] != 0;
            gs[t][0] = ev ? a.grad_raw[3 * i] : 0.f;
            gs[t][1] = ev ? a.grad_raw[3 * i + 1] : 0.f;
            gs[t][2] = ev ? a.grad_raw[3 * i + 2] : 0.f;
            const float px = a.pts[3 * i], py = a.pts[3 * i + 1], pz = a.pts[3 * i + 2];
            const float pn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
            ins[t] = (pn < 1.0f && vmask[j]) ? 1.0f : 0.0f;
            const float dist = j + 1 < N ? z[j + 1] - z[j] : a.sample_dist;
            mids[t] = z[j] + dist * 0.5f;
        }
        in0 = ins[0]; in1 = ins[1]; mid0 = mids[0]; mid1 = mids[1];
        float mid_inside = (0.5f * (in0 + in1) > 0.5f) ? 1.0f : 0.0f;
        mid_inside *= has ? 1.0f : 0.0f;
        const float n0 = sqrtf(g0[0] * g0[0] + g0[1] * g0[1] + g0[2] * g0[2]);
        const float n1 = sqrtf(g1[0] * g1[0] + g1[1] * g1[1] + g1[2] * g1[2]);
        const float cos_d = (g0[0] * g1[0] + g0[1] * g1[1] + g0[2] * g1[2]) / (n0 * n1 + 1e-8f);
        mid_inside *= cos_d > 0.5f ? 1.0f : 0.0f;
        const float sa = sd[j0], sb = sd[j1];
        float z0 = (sa * mid1 - sb * mid0) / (sa - sb + 1e-10f);
        a.mid_inside_out[ray] = mid_inside;
        a.sdf_depth_out[ray] = z0 * cam_dz * mid_inside;
        if (z0 < 0.f) z0 = 0.f;
        if (z0 > *a.z_max) z0 = 0.f;
        a.pts_sdf0_out[3 * ray] = ox + dx * z0;
        a.pts_sdf0_out[3 * ray + 1] = oy + dy * z0;
        a.pts_sdf0_out[3 * ray + 2] = oz + dz * z0;
    }
}

}  // namespace

extern "C" int gens_composite_rays(const gens_composite_args_t* args, void* stream) {
    GENS_CHECK_ARG(args != nullptr);
    const gens_composite_args_t& a = *args;
    if (a.n_rays == 0) return 0;
    GENS_CHECK_ARG(a.n_rays > 0 && a.n_samples > 1 && a.n_src >= 0);
    GENS_CHECK_ARG(a.rays_o && a.rays_d && a.z_vals && a.pts && a.sdf_raw && a.grad_raw && a.smooth_raw && a.colour_raw &&
                   a.voxel_mask && a.evaluated && (a.mask_views || a.n_src == 0) && a.inv_s && a.z_max && a.rot);
    GENS_CHECK_ARG(a.weights_out && a.weight_sum_out && a.weight_max_out && a.color_out && a.normal_out && a.depth_out &&
                   a.inside_out && a.valid_out && a.sdf_out && a.gradients_out && a.mid_inside_out && a.sdf_depth_out &&
                   a.pts_sdf0_out && a.ge_num_out && a.ge_den_out && a.smooth_norm_out);
    if (a.n_samples > kMaxSamples) return GENS_E_UNSUPPORTED;
    composite_kernel<<<ceil_div_i(a.n_rays, kWarpsPerBlock), 32 * kWarpsPerBlock, 0, (cudaStream_t)stream>>>(a);
    return gens_launch_status();
}
