// K1: fused multi-view feature-volume aggregation for sm_100a.
//
// Replaces one scale of Volume.agg_mean_var (reference models/modules/volume.py:21-58), which
// runs ~320 ATen ops and >= 15 full (nv,.,D^3) temporaries per scale, with ONE launch that
// reads each feature map once (through L1/L2) and writes the 8+1 output channels once.
//
// Data layout.  Feature maps are re-packed once per call (one launch for all scales) to "pixel
// pairs": texel (x,y) of view v holds [f(x,y,0..3), f(x+1,y,0..3)] = 32 bytes, with one zero row
// below and zeros for x+1 == W, shape (nv, H+1, W, 8).  A bilinear footprint is then TWO 256-bit
// read-only loads (LDG.E.256: top and bottom pixel pair) instead of four 128-bit ones -- half
// the L1 requests over the same cache lines, which is what bounds this kernel -- and the +1
// corners of any valid sample exist in memory, so the gathers are unconditional.  Outputs are
// the reference's NCDHW planes, written with coalesced streaming stores.
//
// Mapping (packed kernel, D % 128 == 0).  A warp owns 32 CONSECUTIVE voxels along the fastest
// tensor dim (world z); a thread owns four voxels 32 apart.  Consecutive voxels project
// about one pixel apart, so a warp-wide gather touches 4-5 cache lines instead of the 16 a
// "4 consecutive voxels per thread" layout would; every store instruction is one full
// 128-byte line.  The kernel is bound by instruction issue and L1 gather bandwidth, not HBM
// (see DESIGN.md), so the arithmetic runs on Blackwell's packed fp32x2 pipe: two voxels per
// FFMA2 in the projection, two channels per FFMA2 in the bilinear accumulation.
//
// Arithmetic is the reference's, rounding for rounding (oracle/gens_oracle.c is the CPU
// restatement it is tested against bit-for-bit): two-stage projection as k-ascending fma
// chains, IEEE division, two-moment variance E[x^2]-E[x]^2 (NOT Welford -- parity with the
// reference's cancellation behaviour is the contract).
#include <mutex>

#include "common.cuh"
#include "f32x2.cuh"

namespace {

struct __align__(16) Cam {
    float w2c[16];
    float k[12];
    int affine;  // w2c row 3 == (0,0,0,1), K == [[fx,0,cx,0],[0,fy,cy,0],[0,0,1,0]]
    int pad[3];
};

struct Extent {
    float hx, hy, inv_hx, inv_hy;
};

// Multi-GPU builds: every result is stored into the FINAL tensors of all ranks of the node (own memory and
// NVLink peer mappings of the same symmetric allocation), so the slab exchange rides on the kernel's own
// stores instead of an all-gather + scatter afterwards.  n == 0: single destination (volume / mask_volume).
struct PeerOut {
    float* vol[GENS_MAX_PEERS];
    float* msk[GENS_MAX_PEERS];
    int n;
    int self;  // index of this rank's own buffer among vol[] / msk[]
};

// Camera sets in the constant bank.  A set = the staged matrices of up to kCamViews views for ONE scale (K rows
// already multiplied by 0.5^scale); the per-view coefficients then reach the FFMA2s as uniform-register operands
// instead of shared-memory loads through the SM's L1 data pipe (6.3 M of the 37.9 M data-pipe wavefronts of the
// 256^3 launch, profiles/r01_k1_256_ncu_summary.txt).  The sets are filled on the device (the pack launch, or
// gens_stage_cameras) and brought into the bank by a stream-ordered device-to-device cudaMemcpyToSymbolAsync; a ring
// of kCamGroups builds x GENS_MAX_SCALES scales.  Public slot ids are 1-based (0 = stage in shared memory).
constexpr int kCamViews = 8;
constexpr int kCamGroups = 4;
constexpr int kCamSets = kCamGroups * GENS_MAX_SCALES;
__constant__ Cam c_cam[kCamSets * kCamViews];

// Stage the per-view matrices in shared memory.  `k_row_scale` = 0.5^scale multiplies rows 0-1 of
// the intrinsics exactly as the reference's `intrs_stage[:, :2] *= 0.5**i` (volume.py:25); a
// power-of-two factor, so the product is exact and bit-identical to the torch op.
__device__ __forceinline__ int cam_is_affine(const float* w, const float* k) {
    return w[12] == 0.f && w[13] == 0.f && w[14] == 0.f && w[15] == 1.f && k[1] == 0.f && k[3] == 0.f && k[4] == 0.f &&
           k[7] == 0.f && k[8] == 0.f && k[9] == 0.f && k[10] == 1.f && k[11] == 0.f;
}

__device__ __forceinline__ void load_cams(Cam* s_cam, const float* w2c, const float* intrs, float k_row_scale,
                                          int nv) {
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthreads = blockDim.x * blockDim.y;
    for (int i = tid; i < nv * 32; i += nthreads) {
        const int v = i >> 5, j = i & 31;
        if (j < 16) s_cam[v].w2c[j] = w2c[v * 16 + j];
        else if (j < 28) s_cam[v].k[j - 16] = __fmul_rn(intrs[v * 16 + (j - 16)], j < 24 ? k_row_scale : 1.0f);
    }
    __syncthreads();
    if (tid < nv) s_cam[tid].affine = cam_is_affine(s_cam[tid].w2c, s_cam[tid].k);
    __syncthreads();
}

// One view of one camera set, written by ONE thread into the global staging copy of the constant bank
// (same values as load_cams: K rows 0-1 times the exact power of two).
__device__ __forceinline__ void stage_cam(Cam* dst, const float* w2c, const float* intrs, float k_row_scale) {
    Cam c;
#pragma unroll
    for (int j = 0; j < 16; ++j) c.w2c[j] = w2c[j];
#pragma unroll
    for (int j = 0; j < 12; ++j) c.k[j] = __fmul_rn(intrs[j], j < 8 ? k_row_scale : 1.0f);
    c.affine = cam_is_affine(c.w2c, c.k);
    c.pad[0] = c.pad[1] = c.pad[2] = 0;
    *dst = c;
}

// ------------------------------------------------------------------------------------------
// scalar path (any D): one voxel per thread.  Also the parity view of the projection stage.
// ------------------------------------------------------------------------------------------
struct Proj {
    float ix, iy;
    bool valid;
};

template <bool RECIP>
__device__ __forceinline__ Proj project_scalar(const Cam& cam, float X, float Y, float Z, const Extent& e) {
    float c[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) c[r] = row_dot4(cam.w2c + 4 * r, X, Y, Z, 1.0f);
    float img[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) img[r] = row_dot4(cam.k + 4 * r, c[0], c[1], c[2], c[3]);
    const float den = __fadd_rn(img[2], 1e-8f);
    const float x = __fdiv_rn(img[0], den), y = __fdiv_rn(img[1], den);
    const float nx = __fsub_rn(RECIP ? __fmul_rn(x, e.inv_hx) : __fdiv_rn(x, e.hx), 1.0f);
    const float ny = __fsub_rn(RECIP ? __fmul_rn(y, e.inv_hy) : __fdiv_rn(y, e.hy), 1.0f);
    Proj p;
    p.valid = (fabsf(nx) <= 1.0f) && (fabsf(ny) <= 1.0f) && (img[2] > 0.0f);
    // ATen grid_sampler_unnormalize (align_corners=True): ((n+1)*0.5)*(size-1) == (n+1)*((size-1)/2)
    // bit for bit, the halving being exact.
    p.ix = __fmul_rn(__fadd_rn(nx, 1.0f), e.hx);
    p.iy = __fmul_rn(__fadd_rn(ny, 1.0f), e.hy);
    return p;
}

struct Footprint {
    int x0, y0;
    float w_nw, w_ne, w_sw, w_se;
};

// Valid samples have ix in [0, W-1], so ix - floor(ix) is exact and 1 - frac == (floor+1) - ix.
__device__ __forceinline__ Footprint footprint(float ix, float iy) {
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float bx = __fsub_rn(ix, fx0), by = __fsub_rn(iy, fy0);
    const float ax = __fsub_rn(1.0f, bx), ay = __fsub_rn(1.0f, by);
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

// Bilinear sample of a VALID footprint from the pixel-pair map (pitch = W texels of 32 bytes).
__device__ __forceinline__ float4 sample_padded(const Pair* __restrict__ map, int pitch, const Footprint& f) {
    const Pair* p = map + (f.y0 * pitch + f.x0);
    const Pair top = ldg256(p), bot = ldg256(p + pitch);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    fma4(acc, top.a, f.w_nw);
    fma4(acc, top.b, f.w_ne);
    fma4(acc, bot.a, f.w_sw);
    fma4(acc, bot.b, f.w_se);
    return acc;
}

// Exact s/n for the small integer n = number of valid views (Markstein: r = RN(1/n) is
// precomputed, one residual correction gives the correctly rounded quotient).
__device__ __forceinline__ f32x2 div_count2(f32x2 s, float n, float r) {
    const f32x2 q = mul2(s, bc(r));
    const f32x2 rem = fma2(bc(-n), q, s);
    return fma2(rem, bc(r), q);
}

template <bool RECIP>
__global__ void __launch_bounds__(256)
volume_agg_scalar_kernel(const Pair* __restrict__ feat, int nv, int H, int W, const float* __restrict__ w2c,
                         const float* __restrict__ k_stage, float k_row_scale, const float* __restrict__ grid, int D, int a0,
                         long long out_off, long long channel_stride, int min_vis_view, Extent e,
                         float* __restrict__ volume, float* __restrict__ mask_volume, const __grid_constant__ PeerOut peers) {
    __shared__ Cam s_cam[GENS_MAX_VIEWS];
    load_cams(s_cam, w2c, k_stage, k_row_scale, nv);
    const int c = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y * 8 + threadIdx.y, a = a0 + blockIdx.z;
    if (c >= D || b >= D) return;
    const float X = __ldg(grid + a), Y = __ldg(grid + b), Z = __ldg(grid + c);
    const int pitch = W;
    const long long map_stride = (long long)(H + 1) * pitch;

    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    int cnt = 0;
    for (int v = 0; v < nv; ++v) {
        const Proj p = project_scalar<RECIP>(s_cam[v], X, Y, Z, e);
        if (p.valid) {
            const float4 f = sample_padded(feat + v * map_stride, pitch, footprint(p.ix, p.iy));
            cnt += 1;
            s.x = __fadd_rn(s.x, f.x); s.y = __fadd_rn(s.y, f.y);
            s.z = __fadd_rn(s.z, f.z); s.w = __fadd_rn(s.w, f.w);
            q.x = __fadd_rn(q.x, __fmul_rn(f.x, f.x)); q.y = __fadd_rn(q.y, __fmul_rn(f.y, f.y));
            q.z = __fadd_rn(q.z, __fmul_rn(f.z, f.z)); q.w = __fadd_rn(q.w, __fmul_rn(f.w, f.w));
        }
    }
    const float den = cnt <= 0 ? 1e-8f : (float)cnt;
    const float sv[4] = {s.x, s.y, s.z, s.w}, qv[4] = {q.x, q.y, q.z, q.w};
    const long long o = out_off + ((long long)blockIdx.z * D + b) * D + c;
    float res[9];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        res[k] = __fdiv_rn(sv[k], den);
        res[4 + k] = __fsub_rn(__fdiv_rn(qv[k], den), __fmul_rn(res[k], res[k]));
    }
    res[8] = cnt > min_vis_view ? 1.0f : 0.0f;
    const int n_dst = peers.n > 0 ? peers.n : 1;
    for (int d = 0; d < n_dst; ++d) {
        float* vol = peers.n > 0 ? peers.vol[d] : volume;
        float* msk = peers.n > 0 ? peers.msk[d] : mask_volume;
#pragma unroll
        for (int k = 0; k < 8; ++k) __stcs(vol + k * channel_stride + o, res[k]);
        __stcs(msk + o, res[8]);
    }
}

template <bool RECIP>
__global__ void __launch_bounds__(256)
volume_project_debug_kernel(int nv, const float* __restrict__ w2c, const float* __restrict__ k_stage,
                            float k_row_scale, const float* __restrict__ grid, int D, Extent e, int32_t* __restrict__ ix0,
                            int32_t* __restrict__ iy0, uint8_t* __restrict__ valid) {
    __shared__ Cam s_cam[GENS_MAX_VIEWS];
    load_cams(s_cam, w2c, k_stage, k_row_scale, nv);
    const int c = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y * 8 + threadIdx.y, a = blockIdx.z;
    if (c >= D || b >= D) return;
    const long long D3 = (long long)D * D * D, n = ((long long)a * D + b) * D + c;
    const float X = grid[a], Y = grid[b], Z = grid[c];
    for (int v = 0; v < nv; ++v) {
        const Proj p = project_scalar<RECIP>(s_cam[v], X, Y, Z, e);
        const Footprint f = footprint(p.valid ? p.ix : 0.f, p.valid ? p.iy : 0.f);
        ix0[v * D3 + n] = p.valid ? f.x0 : 0;
        iy0[v * D3 + n] = p.valid ? f.y0 : 0;
        valid[v * D3 + n] = p.valid ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------
// packed path (D % (64*PAIRS) == 0): PAIRS voxel pairs per thread, two voxels per FFMA2.
// ------------------------------------------------------------------------------------------
struct Proj2 {
    f32x2 ix, iy;
    bool valid_lo, valid_hi;
};

// Projection of a voxel pair sharing (X,Y); `pre[r]` = fma(w[r][1], Y, w[r][0]*X).
template <bool RECIP, bool AFFINE>
__device__ __forceinline__ Proj2 project_pair(const Cam& cam, const float (&pre)[4], f32x2 Z, const Extent& e) {
    f32x2 img0, img1, depth;
    if (AFFINE) {
        // Same chain with the structural zeros/ones of a rigid pose and a pinhole K removed:
        // fma(0, x, t) == t and fma(1, x, +-0) == x for finite x, so every surviving rounding
        // is identical.
        f32x2 c[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            c[r] = add2(fma2(bc(cam.w2c[4 * r + 2]), Z, bc(pre[r])), bc(cam.w2c[4 * r + 3]));
        img0 = fma2(bc(cam.k[2]), c[2], mul2(bc(cam.k[0]), c[0]));
        img1 = fma2(bc(cam.k[6]), c[2], mul2(bc(cam.k[5]), c[1]));
        depth = c[2];
    } else {
        f32x2 c[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
            c[r] = add2(fma2(bc(cam.w2c[4 * r + 2]), Z, bc(pre[r])), bc(cam.w2c[4 * r + 3]));
        f32x2 img[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            f32x2 t = mul2(bc(cam.k[4 * r]), c[0]);
            t = fma2(bc(cam.k[4 * r + 1]), c[1], t);
            t = fma2(bc(cam.k[4 * r + 2]), c[2], t);
            img[r] = fma2(bc(cam.k[4 * r + 3]), c[3], t);
        }
        img0 = img[0]; img1 = img[1]; depth = img[2];
    }
    const float d_lo = lo(depth), d_hi = hi(depth);
    const f32x2 den = add2(depth, bc(1e-8f));
    // depth <= 0 is invalid whatever x/y are; depth in (0, 2^100) puts den in [1e-8, 2^100], the
    // range div2() is exact on.  Anything else (never, for real cameras) takes IEEE division.
    const Recip2 rc = recip2(den);
    f32x2 x = div2(img0, rc), y = div2(img1, rc);
    if (__builtin_expect(d_lo >= 1e30f || d_hi >= 1e30f, 0)) {
        x = pk(__fdiv_rn(lo(img0), lo(den)), __fdiv_rn(hi(img0), hi(den)));
        y = pk(__fdiv_rn(lo(img1), lo(den)), __fdiv_rn(hi(img1), hi(den)));
    }
    f32x2 nx, ny;
    if (RECIP) {
        nx = add2(mul2_rounded(x, bc(e.inv_hx)), bc(-1.0f));
        ny = add2(mul2_rounded(y, bc(e.inv_hy)), bc(-1.0f));
    } else {
        nx = add2(pk(__fdiv_rn(lo(x), e.hx), __fdiv_rn(hi(x), e.hx)), bc(-1.0f));
        ny = add2(pk(__fdiv_rn(lo(y), e.hy), __fdiv_rn(hi(y), e.hy)), bc(-1.0f));
    }
    Proj2 p;
    p.valid_lo = (fabsf(lo(nx)) <= 1.0f) && (fabsf(lo(ny)) <= 1.0f) && (d_lo > 0.0f);
    p.valid_hi = (fabsf(hi(nx)) <= 1.0f) && (fabsf(hi(ny)) <= 1.0f) && (d_hi > 0.0f);
    p.ix = mul2(add2(nx, bc(1.0f)), bc(e.hx));
    p.iy = mul2(add2(ny, bc(1.0f)), bc(e.hy));
    return p;
}

struct Acc {
    f32x2 s_xy, s_zw, q_xy, q_zw;
    int cnt;
};

struct Corner4 {
    float4 nw, ne, sw, se;
};

__device__ __forceinline__ Corner4 gather4(const Pair* __restrict__ map, int pitch, int x0, int y0) {
    // 32-bit texel indices scaled onto a per-view base held in a register pair: one IMAD.WIDE per address
    // instead of a sign extension, a 64-bit add of the view offset and a 64-bit shift-add (launch_agg_fwd
    // bounds the map size so that the index fits).
    const unsigned i = (unsigned)(y0 * pitch + x0);
    const Pair top = ldg256(map + i), bot = ldg256(map + (i + (unsigned)pitch));
    Corner4 c;
    c.nw = top.a;
    c.ne = top.b;
    c.sw = bot.a;
    c.se = bot.b;
    return c;
}

__device__ __forceinline__ void blend_accumulate(const Corner4& v, float w_nw, float w_ne, float w_sw, float w_se,
                                                 Acc& a) {
    f32x2 f_xy = mul2(pk(v.nw.x, v.nw.y), bc(w_nw));  // fma(v, w, 0) == v*w
    f32x2 f_zw = mul2(pk(v.nw.z, v.nw.w), bc(w_nw));
    f_xy = fma2(pk(v.ne.x, v.ne.y), bc(w_ne), f_xy);
    f_zw = fma2(pk(v.ne.z, v.ne.w), bc(w_ne), f_zw);
    f_xy = fma2(pk(v.sw.x, v.sw.y), bc(w_sw), f_xy);
    f_zw = fma2(pk(v.sw.z, v.sw.w), bc(w_sw), f_zw);
    f_xy = fma2(pk(v.se.x, v.se.y), bc(w_se), f_xy);
    f_zw = fma2(pk(v.se.z, v.se.w), bc(w_se), f_zw);
    a.cnt += 1;
    a.s_xy = add2(a.s_xy, f_xy);
    a.s_zw = add2(a.s_zw, f_zw);
    a.q_xy = add2(a.q_xy, mul2_rounded(f_xy, f_xy));
    a.q_zw = add2(a.q_zw, mul2_rounded(f_zw, f_zw));
}

// One view for all of a thread's voxel pairs.  GATHER = 0: both voxels of a pair issue their
// four gathers before the first blend (8 loads in flight per thread); GATHER = 1: one voxel at a
// time (4 in flight, 16 fewer live registers).
template <int PAIRS, bool RECIP, bool AFFINE, int GATHER>
__device__ __forceinline__ void accumulate_view(const Cam& cam, const Pair* __restrict__ map, int pitch,
                                                float X, float Y, const f32x2 (&Z)[PAIRS], const Extent& e,
                                                Acc (&acc)[2 * PAIRS]) {
    // keep the view's base pointer in a register pair: otherwise ptxas folds the (uniform) view offset into every
    // address as a 64-bit add + shift-add, six integer instructions per gather instead of two
    asm("" : "+l"(map));
    float pre[4];
#pragma unroll
    for (int r = 0; r < (AFFINE ? 3 : 4); ++r)
        pre[r] = __fmaf_rn(cam.w2c[4 * r + 1], Y, __fmul_rn(cam.w2c[4 * r], X));
#pragma unroll
    for (int h = 0; h < PAIRS; ++h) {
        const Proj2 p = project_pair<RECIP, AFFINE>(cam, pre, Z[h], e);
        if (!(p.valid_lo || p.valid_hi)) continue;
        const float fx_lo = floorf(lo(p.ix)), fx_hi = floorf(hi(p.ix));
        const float fy_lo = floorf(lo(p.iy)), fy_hi = floorf(hi(p.iy));
        const f32x2 bx = sub2(p.ix, pk(fx_lo, fx_hi)), by = sub2(p.iy, pk(fy_lo, fy_hi));
        const f32x2 ax = sub2(bc(1.0f), bx), ay = sub2(bc(1.0f), by);
        const f32x2 w_nw = mul2(ax, ay), w_ne = mul2(bx, ay), w_sw = mul2(ax, by), w_se = mul2(bx, by);
        if (GATHER == 0) {
            Corner4 v_lo, v_hi;
            if (p.valid_lo) v_lo = gather4(map, pitch, (int)fx_lo, (int)fy_lo);
            if (p.valid_hi) v_hi = gather4(map, pitch, (int)fx_hi, (int)fy_hi);
            if (p.valid_lo) blend_accumulate(v_lo, lo(w_nw), lo(w_ne), lo(w_sw), lo(w_se), acc[2 * h]);
            if (p.valid_hi) blend_accumulate(v_hi, hi(w_nw), hi(w_ne), hi(w_sw), hi(w_se), acc[2 * h + 1]);
        } else {
            if (p.valid_lo)
                blend_accumulate(gather4(map, pitch, (int)fx_lo, (int)fy_lo), lo(w_nw), lo(w_ne), lo(w_sw),
                                 lo(w_se), acc[2 * h]);
            if (p.valid_hi)
                blend_accumulate(gather4(map, pitch, (int)fx_hi, (int)fy_hi), hi(w_nw), hi(w_ne), hi(w_sw),
                                 hi(w_se), acc[2 * h + 1]);
        }
    }
}

// ROWS: row-groups of 8 rows a block walks through, amortising the camera staging and its barriers.
template <int PAIRS, bool RECIP, int MIN_BLOCKS, int GATHER, int ROWS, bool PEERS = false>
__global__ void __launch_bounds__(256, MIN_BLOCKS)
volume_agg_packed_kernel(const Pair* __restrict__ feat, int nv, int H, int W, const float* __restrict__ w2c,
                         const float* __restrict__ k_stage, float k_row_scale, const float* __restrict__ grid, int D, int a0,
                         long long out_off, long long channel_stride, int min_vis_view, Extent e,
                         float* __restrict__ volume, float* __restrict__ mask_volume, const __grid_constant__ PeerOut peers) {
    __shared__ Cam s_cam[GENS_MAX_VIEWS];
    __shared__ float s_inv_count[GENS_MAX_VIEWS + 1];  // RN(1/n) for the exact count division
    if (threadIdx.y == 0 && threadIdx.x <= GENS_MAX_VIEWS)
        s_inv_count[threadIdx.x] = threadIdx.x == 0 ? 1e8f : __fdiv_rn(1.0f, (float)threadIdx.x);
    load_cams(s_cam, w2c, k_stage, k_row_scale, nv);
    const int c0 = blockIdx.x * (64 * PAIRS) + threadIdx.x, a = a0 + blockIdx.z;
    const float X = __ldg(grid + a);
    f32x2 Z[PAIRS];
#pragma unroll
    for (int h = 0; h < PAIRS; ++h) Z[h] = pk(__ldg(grid + c0 + 64 * h), __ldg(grid + c0 + 64 * h + 32));
    const int pitch = W;
    const long long map_stride = (long long)(H + 1) * pitch;

#pragma unroll 1
    for (int rr = 0; rr < ROWS; ++rr) {
        const int b = (blockIdx.y * ROWS + rr) * 8 + threadIdx.y;
        if (b >= D) return;
        const float Y = __ldg(grid + b);
        Acc acc[2 * PAIRS];
#pragma unroll
        for (int j = 0; j < 2 * PAIRS; ++j) {
            acc[j].s_xy = acc[j].s_zw = acc[j].q_xy = acc[j].q_zw = bc(0.0f);
            acc[j].cnt = 0;
        }

#pragma unroll 1
        for (int v = 0; v < nv; ++v) {
            const Pair* map = feat + v * map_stride;
            if (s_cam[v].affine) accumulate_view<PAIRS, RECIP, true, GATHER>(s_cam[v], map, pitch, X, Y, Z, e, acc);
            else accumulate_view<PAIRS, RECIP, false, GATHER>(s_cam[v], map, pitch, X, Y, Z, e, acc);
        }

        const long long row = out_off + ((long long)blockIdx.z * D + b) * D + c0;
#pragma unroll
        for (int j = 0; j < 2 * PAIRS; ++j) {
            const int cnt = acc[j].cnt;
            const float n = cnt <= 0 ? 1e-8f : (float)cnt;
            const float r = s_inv_count[cnt];
            const f32x2 m_xy = div_count2(acc[j].s_xy, n, r), m_zw = div_count2(acc[j].s_zw, n, r);
            const f32x2 v_xy = sub2(div_count2(acc[j].q_xy, n, r), mul2_rounded(m_xy, m_xy));
            const f32x2 v_zw = sub2(div_count2(acc[j].q_zw, n, r), mul2_rounded(m_zw, m_zw));
            if (PEERS) {
                // one store per channel and destination: the own tensor and every NVLink peer's
                const float mk = cnt > min_vis_view ? 1.0f : 0.0f;
                for (int d = 0; d < peers.n; ++d) {
                    float* o = peers.vol[d] + row + 32 * j;
                    __stcs(o, lo(m_xy));
                    __stcs(o + channel_stride, hi(m_xy));
                    __stcs(o + 2 * channel_stride, lo(m_zw));
                    __stcs(o + 3 * channel_stride, hi(m_zw));
                    __stcs(o + 4 * channel_stride, lo(v_xy));
                    __stcs(o + 5 * channel_stride, hi(v_xy));
                    __stcs(o + 6 * channel_stride, lo(v_zw));
                    __stcs(o + 7 * channel_stride, hi(v_zw));
                    __stcs(peers.msk[d] + row + 32 * j, mk);
                }
                continue;
            }
            float* o = volume + row + 32 * j;
            __stcs(o, lo(m_xy));
            __stcs(o + channel_stride, hi(m_xy));
            __stcs(o + 2 * channel_stride, lo(m_zw));
            __stcs(o + 3 * channel_stride, hi(m_zw));
            __stcs(o + 4 * channel_stride, lo(v_xy));
            __stcs(o + 5 * channel_stride, hi(v_xy));
            __stcs(o + 6 * channel_stride, lo(v_zw));
            __stcs(o + 7 * channel_stride, hi(v_zw));
            __stcs(mask_volume + row + 32 * j, cnt > min_vis_view ? 1.0f : 0.0f);
        }
    }
}

// ------------------------------------------------------------------------------------------
// row-group path (D % 64 == 0, single destination): the packed arithmetic above plus conservative frustum
// culling per (tile of 8 rows x 64 voxels, view), decided once per block from the four corners of the tile.
// On the 480x640 / 3-view scene 47 % of the voxels are seen by no view at all and 38 % of the tiles are
// entirely dead: those cost their zero stores and no arithmetic; in the others 10 % of the tile-views are
// skipped.  Measured alternatives (TMA box stores, walking x instead of y, other gather layouts, cameras in
// the constant bank) are listed with their times in profiles/r01_k1_variant_sweep.txt.
// ------------------------------------------------------------------------------------------

// Frustum planes a point lies outside of WITH MARGIN (bit 0: behind the camera, 1: left, 2: right, 3: top,
// 4: bottom).  Every test is a linear function L of the world point with a tolerance tau * (sum of the
// magnitudes of its terms): if L > tol at the four corners of a tile (fixed X, Y and Z ranges), L > 0
// at every voxel inside, and then the voxel is invalid in the reference's arithmetic:
//   behind:  depth < 0                                     -> `img2 > 0` fails (volume.py:43)
//   left:    -img0 - m hx depth > 0  => x / hx < -m        -> |nx| = |x/hx - 1| > 1 (or depth <= 0)
//   right:   img0 - (2 + m) hx depth > 0 => x / hx > 2 + m -> |nx| > 1 (or depth <= 0)
// with m = 1e-3 (a third of a pixel) and tau = 1e-4, both orders of magnitude above the rounding error of the
// fp32 chains (<= 1e-6 relative to the same magnitude sums).  Samples that land exactly on or marginally
// outside the image border are therefore never culled; they take the exact path.  NaN / Inf cameras fail
// every comparison and cull nothing.  Only used for cameras with the rigid-pose / pinhole structure.
__device__ __forceinline__ unsigned cull_planes(const Cam& cam, float X, float Y, float Z, const Extent& e) {
    float c[3], s[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float* w = cam.w2c + 4 * r;
        c[r] = w[0] * X + w[1] * Y + w[2] * Z + w[3];
        s[r] = fabsf(w[0] * X) + fabsf(w[1] * Y) + fabsf(w[2] * Z) + fabsf(w[3]);
    }
    const float depth = c[2];
    const float img0 = cam.k[0] * c[0] + cam.k[2] * c[2], S0 = fabsf(cam.k[0]) * s[0] + fabsf(cam.k[2]) * s[2];
    const float img1 = cam.k[5] * c[1] + cam.k[6] * c[2], S1 = fabsf(cam.k[5]) * s[1] + fabsf(cam.k[6]) * s[2];
    const float tau = 1e-4f, m = 1e-3f;
    const float lx = m * e.hx, rx = (2.0f + m) * e.hx, ly = m * e.hy, ry = (2.0f + m) * e.hy;
    unsigned bits = 0;
    if (depth < -tau * s[2]) bits |= 1u;
    if (-img0 - lx * depth > tau * (S0 + lx * s[2])) bits |= 2u;
    if (img0 - rx * depth > tau * (S0 + rx * s[2])) bits |= 4u;
    if (-img1 - ly * depth > tau * (S1 + ly * s[2])) bits |= 8u;
    if (img1 - ry * depth > tau * (S1 + ry * s[2])) bits |= 16u;
    return bits;
}

// A block owns ROWS tiles of 8 rows (y) x 64 voxels (z) of one x plane.  CULL = false keeps every tile-view
// (A/B reference for the culling, variant 25 of the tuning knob).
// o + stride as ONE 64-bit add the compiler cannot turn back into (base + k * stride) multiplies
__device__ __forceinline__ float* next_plane(float* o, long long stride_elems) {
    float* n = o + stride_elems;
    asm("" : "+l"(n));
    return n;
}

// 256 bytes of zeros -> one output row segment (64 voxels of one plane), written by the bulk-copy engine: the
// store never enters the SM's L1 data pipe (STG costs it 4 wavefronts per 128 bytes; the dead tiles are 38 % of
// the 256^3 launch's 18.9 M store wavefronts) and costs one instruction per 256 bytes instead of two.
__device__ __forceinline__ void bulk_zero_row(float* dst, const float* s_zero) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 256;"
                 :: "l"(dst), "r"((uint32_t)__cvta_generic_to_shared(s_zero)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

template <bool CONSTCAM>
__device__ __forceinline__ const Cam& cam_of(const Cam* s_cam, int cam_set, int v) {
    if (CONSTCAM) return c_cam[cam_set * kCamViews + v];
    return s_cam[v];
}

// A block owns ROWS tiles of 8 rows (y) x 64 voxels (z) of one x plane.
//   CULL     conservative frustum culling per (tile, view); false = A/B reference (tuning variant 25)
//   CONSTCAM cameras from the constant bank (set `cam_set` of c_cam) instead of shared memory
//   ZBULK    dead tiles are zero-filled by the bulk-copy engine instead of STG
//   PEERS    multi-GPU: the grid covers ALL planes of the volume.  Tiles of this rank's slab [a0,a1): live ones are
//            computed and stored into every rank's tensor (own + NVLink peer mappings), dead ones are zero-filled in
//            the own tensor only.  Tiles of the other ranks' slabs: dead ones are zero-filled locally as well (every
//            rank reaches the same cull decision from the same cameras, so nobody ships zeros over NVLink), live
//            ones are left to their owner.
// Measured and not kept (profiles/r02_k1_variant_sweep.txt): walking x planes with a fixed window (L1 re-use across
// planes does not materialise: 187-225 us against 177), 5 / 6 resident blocks per SM at <= 48 / 40 registers (spills:
// 263 / 306 us), 2-3 blocks with both voxels' gathers in flight (212 us and worse).
template <bool RECIP, int ROWS, bool CULL, bool CONSTCAM, bool ZBULK, bool PEERS>
__global__ void __launch_bounds__(256, 4)
volume_agg_rowgroup_kernel(const Pair* __restrict__ feat, int nv, int H, int W, const float* __restrict__ w2c,
                           const float* __restrict__ k_stage, float k_row_scale, const float* __restrict__ grid, int D,
                           int a0, int a1, long long out_off, long long channel_stride, int min_vis_view, Extent e,
                           float* __restrict__ volume, float* __restrict__ mask_volume, int cam_set,
                           const __grid_constant__ PeerOut peers) {
    __shared__ Cam s_cam[CONSTCAM ? 1 : GENS_MAX_VIEWS];
    __shared__ float s_inv_count[GENS_MAX_VIEWS + 1];
    __shared__ unsigned s_live[ROWS];  // per tile: bit v set = view v may see a voxel of it
    __shared__ __align__(128) float s_zero[64];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (threadIdx.y == 0 && threadIdx.x <= GENS_MAX_VIEWS)
        s_inv_count[threadIdx.x] = threadIdx.x == 0 ? 1e8f : __fdiv_rn(1.0f, (float)threadIdx.x);
    if (tid < ROWS) s_live[tid] = CULL ? 0u : 0xffffffffu;
    if (ZBULK && threadIdx.y == 1) {
        s_zero[threadIdx.x] = 0.0f;
        s_zero[threadIdx.x + 32] = 0.0f;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk copies
    }
    if (CONSTCAM) __syncthreads();
    else load_cams(s_cam, w2c, k_stage, k_row_scale, nv);  // two barriers inside

    const int cz0 = blockIdx.x * 64, c0 = cz0 + threadIdx.x;
    // tile rr of this block: plane a_first, rows b_first + 8 rr ...
    const int a_first = PEERS ? (int)blockIdx.z : a0 + (int)blockIdx.z;  // plane of tensor dim 2 (world x)
    const int b_first = (int)blockIdx.y * 8 * ROWS;
    const bool own = !PEERS || (a_first >= a0 && a_first < a1);
    const float X = __ldg(grid + a_first);
    if (CULL) {
        const int total = ROWS * nv * 4;  // one item = one corner of one (tile, view); 4 consecutive lanes = one (tile, view)
        for (int base = 0; base < total; base += 256) {
            const int i = base + tid;
            const bool active = i < total;
            const int corner = i & 3, v = active ? (i >> 2) % nv : 0, rg = active ? (i >> 2) / nv : 0;
            const int b0 = b_first + 8 * rg;
            unsigned bits = 0;
            if (active && b0 < D) {
                const Cam& cam = cam_of<CONSTCAM>(s_cam, cam_set, v);
                if (cam.affine)
                    bits = cull_planes(cam, X, __ldg(grid + b0 + (corner & 1) * 7),
                                       __ldg(grid + cz0 + (corner >> 1) * 63), e);
            }
            bits &= __shfl_xor_sync(0xffffffffu, bits, 1);
            bits &= __shfl_xor_sync(0xffffffffu, bits, 2);
            if (active && corner == 0 && bits == 0) atomicOr(&s_live[rg], 1u << v);
        }
        __syncthreads();
    }
    const f32x2 Z[1] = {pk(__ldg(grid + c0), __ldg(grid + c0 + 32))};
    const int pitch = W;
    const long long map_stride = (long long)(H + 1) * pitch;

    // running output offset of this thread's voxel pair (plane 0 of the volume, and the mask): one 64-bit add per
    // tile instead of rebuilding ((x * D + y) * D + z) and nine channel offsets from scratch
    long long off = (PEERS ? 0LL : out_off) +
                    ((long long)(PEERS ? a_first : a_first - a0) * D + (b_first + threadIdx.y)) * D + c0;
    float* const own_vol = PEERS ? peers.vol[peers.self] : volume;
    float* const own_msk = PEERS ? peers.msk[peers.self] : mask_volume;
    const long long tile_step = 8LL * D;
    bool bulk_pending = false;
#pragma unroll 1
    for (int rr = 0; rr < ROWS; ++rr, off += tile_step) {
        const int b0 = b_first + 8 * rr, b = b0 + threadIdx.y;
        if (b0 >= D) break;
        const unsigned live = s_live[rr];
        if (live == 0) {  // block-uniform: nothing of this tile is visible anywhere -> zeros (min_vis_view >= 0)
            if (ZBULK) {
                // lane k < 9 of every warp: the 256-byte segment of its row in plane k (8 channels, then the mask)
                if (threadIdx.x < 9) {
                    const long long row0 = off - threadIdx.x;
                    float* dst = threadIdx.x < 8 ? own_vol + threadIdx.x * channel_stride + row0 : own_msk + row0;
                    bulk_zero_row(dst, s_zero);
                    bulk_commit();
                    bulk_pending = true;
                }
            } else {
                float* o = own_vol + off;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    __stcs(o, 0.0f);
                    __stcs(o + 32, 0.0f);
                    o = next_plane(o, channel_stride);
                }
                __stcs(own_msk + off, 0.0f);
                __stcs(own_msk + off + 32, 0.0f);
            }
            continue;
        }
        if (!own) continue;  // a live tile of another rank's slab: its owner stores it into this rank's tensor
        const float Y = __ldg(grid + b);
        Acc acc[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            acc[j].s_xy = acc[j].s_zw = acc[j].q_xy = acc[j].q_zw = bc(0.0f);
            acc[j].cnt = 0;
        }
#pragma unroll 1
        for (int v = 0; v < nv; ++v) {
            if (!((live >> v) & 1u)) continue;
            const Pair* map = feat + v * map_stride;
            const Cam& cam = cam_of<CONSTCAM>(s_cam, cam_set, v);
            if (cam.affine) accumulate_view<1, RECIP, true, 1>(cam, map, pitch, X, Y, Z, e, acc);
            else accumulate_view<1, RECIP, false, 1>(cam, map, pitch, X, Y, Z, e, acc);
        }
        float res[2][9];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int cnt = acc[j].cnt;
            const float n = cnt <= 0 ? 1e-8f : (float)cnt;
            const float r = s_inv_count[cnt];
            const f32x2 m_xy = div_count2(acc[j].s_xy, n, r), m_zw = div_count2(acc[j].s_zw, n, r);
            const f32x2 v_xy = sub2(div_count2(acc[j].q_xy, n, r), mul2_rounded(m_xy, m_xy));
            const f32x2 v_zw = sub2(div_count2(acc[j].q_zw, n, r), mul2_rounded(m_zw, m_zw));
            res[j][0] = lo(m_xy); res[j][1] = hi(m_xy); res[j][2] = lo(m_zw); res[j][3] = hi(m_zw);
            res[j][4] = lo(v_xy); res[j][5] = hi(v_xy); res[j][6] = lo(v_zw); res[j][7] = hi(v_zw);
            res[j][8] = cnt > min_vis_view ? 1.0f : 0.0f;
        }
        // plane after plane: one 64-bit pointer bump per channel, both voxels of the pair off the same register
        if (PEERS) {
#pragma unroll 1
            for (int d = 0; d < peers.n; ++d) {
                float* o = peers.vol[d] + off;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    __stcs(o, res[0][k]);
                    __stcs(o + 32, res[1][k]);
                    o = next_plane(o, channel_stride);
                }
                float* m = peers.msk[d] + off;
                __stcs(m, res[0][8]);
                __stcs(m + 32, res[1][8]);
            }
        } else {
            float* o = own_vol + off;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                __stcs(o, res[0][k]);
                __stcs(o + 32, res[1][k]);
                o = next_plane(o, channel_stride);
            }
            __stcs(own_msk + off, res[0][8]);
            __stcs(own_msk + off + 32, res[1][8]);
        }
    }
    if (ZBULK && bulk_pending) bulk_wait_read_all();  // the zero row must outlive the bulk engine's reads of it
}

// Backward w.r.t. the feature maps.  With m_v the view validity, n' the clamped count,
//   mean_k = sum_v m_v f_vk / n',  var_k = sum_v m_v f_vk^2 / n' - mean_k^2
//   d/df_vk = m_v / n' * ( g_mean_k + 2 g_var_k (f_vk - mean_k) )
// scattered to the four bilinear corners of the padded channels-last gradient map with
// 16-byte vector atomics (two passes over the views: mean first, then scatter).  The features are read
// from the forward's pixel-pair maps; the gradient map is (nv, H+1, W+1, 4).
template <bool RECIP>
__global__ void __launch_bounds__(256)
volume_agg_bwd_kernel(const Pair* __restrict__ feat, int nv, int H, int W, const float* __restrict__ w2c,
                      const float* __restrict__ k_stage, float k_row_scale, const float* __restrict__ grid, int D, int a0,
                      long long out_off, long long channel_stride, Extent e,
                      const float* __restrict__ grad_volume, float4* __restrict__ grad_feat) {
    __shared__ Cam s_cam[GENS_MAX_VIEWS];
    load_cams(s_cam, w2c, k_stage, k_row_scale, nv);
    const int c = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y * 8 + threadIdx.y, a = a0 + blockIdx.z;
    if (c >= D || b >= D) return;
    const float X = __ldg(grid + a), Y = __ldg(grid + b), Z = __ldg(grid + c);
    const long long o = out_off + ((long long)blockIdx.z * D + b) * D + c;
    float gm[4], gv[4];
    bool any = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        gm[k] = __ldg(grad_volume + k * channel_stride + o);
        gv[k] = __ldg(grad_volume + (4 + k) * channel_stride + o);
        any |= (gm[k] != 0.f) | (gv[k] != 0.f);
    }
    if (!any) return;
    const int pitch = W, g_pitch = W + 1;  // features: pixel pairs; gradient: padded channels-last float4
    const long long map_stride = (long long)(H + 1) * pitch, g_stride = (long long)(H + 1) * g_pitch;

    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt = 0;
#pragma unroll 1
    for (int v = 0; v < nv; ++v) {
        const Proj p = project_scalar<RECIP>(s_cam[v], X, Y, Z, e);
        if (p.valid) {
            const float4 f = sample_padded(feat + v * map_stride, pitch, footprint(p.ix, p.iy));
            cnt += 1;
            s.x += f.x; s.y += f.y; s.z += f.z; s.w += f.w;
        }
    }
    if (cnt == 0) return;
    const float inv_n = 1.0f / (float)cnt;
    const float mean[4] = {s.x * inv_n, s.y * inv_n, s.z * inv_n, s.w * inv_n};
#pragma unroll 1
    for (int v = 0; v < nv; ++v) {
        const Proj p = project_scalar<RECIP>(s_cam[v], X, Y, Z, e);
        if (!p.valid) continue;
        const Footprint t = footprint(p.ix, p.iy);
        const float4 f = sample_padded(feat + v * map_stride, pitch, t);
        const float fv[4] = {f.x, f.y, f.z, f.w};
        float gf[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) gf[k] = inv_n * (gm[k] + 2.0f * gv[k] * (fv[k] - mean[k]));
        float4* gp = grad_feat + v * g_stride + (t.y0 * g_pitch + t.x0);
        atomicAdd(gp, make_float4(gf[0] * t.w_nw, gf[1] * t.w_nw, gf[2] * t.w_nw, gf[3] * t.w_nw));
        atomicAdd(gp + 1, make_float4(gf[0] * t.w_ne, gf[1] * t.w_ne, gf[2] * t.w_ne, gf[3] * t.w_ne));
        atomicAdd(gp + g_pitch, make_float4(gf[0] * t.w_sw, gf[1] * t.w_sw, gf[2] * t.w_sw, gf[3] * t.w_sw));
        atomicAdd(gp + g_pitch + 1, make_float4(gf[0] * t.w_se, gf[1] * t.w_se, gf[2] * t.w_se, gf[3] * t.w_se));
    }
}

// All scales of a pyramid, (n,4,h_i,w_i) NCHW -> pixel pairs (n, h_i+1, w_i, 8), ONE launch: a thread
// writes one 32-byte texel [f(x,y,:), f(x+1,y,:)] (zeros for y == h and x+1 == w).
struct PackJob {
    const float* src;
    float4* dst;
    int h, w;
    long long first;  // index of this scale's first texel in the launch-wide numbering
};
struct PackJobs {
    PackJob job[GENS_MAX_SCALES];
    int n_jobs, n_maps;
    long long total;
    const float* poses;  // optional: (n_poses,4,4) matrices to invert beside the packing (world-to-camera)
    float* poses_inv;
    int n_poses;
    // optional: stage the cameras of `n_cam_sets` scales (K rows 0-1 times cam_scale[s]) into `cam_stage`
    // (n_cam_sets x kCamViews Cam records, the global staging copy of a constant-bank group); needs n_poses <= kCamViews
    const float* intrs;
    Cam* cam_stage;
    int n_cam_sets;
    float cam_scale[GENS_MAX_SCALES];
};

// inverse(A) for one 4x4 matrix, rounding for rounding what torch.linalg.inv_ex / torch.inverse return on
// CUDA (cuBLAS getrfBatched + getrsBatched on the identity): partial-pivoting LU whose multipliers are
// a * RN(1/pivot), fused multiply-subtract updates, unit-lower forward substitution in ascending order,
// back substitution accumulating from the last column down, true division by the diagonal.  Identified
// offline (tools/fit_inverse.py: 20 016 of 20 016 camera matrices bit-identical) and pinned by
// tests/test_volume_gpu.py::test_pose_inverse_is_bit_identical_to_torch.
__device__ void invert4x4_like_torch(const float* __restrict__ A, float* __restrict__ out) {
    float a[4][4];
    int perm[4] = {0, 1, 2, 3};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) a[i][j] = A[4 * i + j];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int p = k;
        float best = fabsf(a[k][k]);
#pragma unroll
        for (int i = k + 1; i < 4; ++i)
            if (fabsf(a[i][k]) > best) { best = fabsf(a[i][k]); p = i; }
#pragma unroll
        for (int i = k + 1; i < 4; ++i)
            if (i == p) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float t = a[k][j]; a[k][j] = a[i][j]; a[i][j] = t; }
                const int t = perm[k]; perm[k] = perm[i]; perm[i] = t;
            }
        const float r = __frcp_rn(a[k][k]);
#pragma unroll
        for (int i = k + 1; i < 4; ++i) {
            const float l = __fmul_rn(a[i][k], r);
            a[i][k] = l;
#pragma unroll
            for (int j = k + 1; j < 4; ++j) a[i][j] = __fmaf_rn(-l, a[k][j], a[i][j]);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {  // column c of the inverse: solve L U x = P e_c
        float x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float y = perm[i] == c ? 1.0f : 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < i) y = __fmaf_rn(-a[i][j], x[j], y);
            x[i] = y;
        }
#pragma unroll
        for (int i = 3; i >= 0; --i) {
            float y = x[i];
#pragma unroll
            for (int j = 3; j >= 0; --j)
                if (j > i) y = __fmaf_rn(-a[i][j], x[j], y);
            x[i] = __fdiv_rn(y, a[i][i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) out[4 * i + c] = x[i];
    }
}

__global__ void __launch_bounds__(64) invert_poses_kernel(const float* __restrict__ poses, float* __restrict__ inv,
                                                          int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) invert4x4_like_torch(poses + 16 * i, inv + 16 * i);
}

struct CamScales {
    float s[GENS_MAX_SCALES];
};
__global__ void __launch_bounds__(64)
stage_cams_kernel(const float* __restrict__ w2c, const float* __restrict__ intrs, int nv, int n_sets,
                  const __grid_constant__ CamScales scales, Cam* __restrict__ stage) {
    for (int t = threadIdx.x; t < n_sets * nv; t += blockDim.x) {
        const int set = t / nv, v = t % nv;
        stage_cam(stage + set * kCamViews + v, w2c + 16 * v, intrs + 16 * v, scales.s[set]);
    }
}

__global__ void __launch_bounds__(256)
pack_pairs_kernel(const __grid_constant__ PackJobs jobs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < jobs.n_poses) invert4x4_like_torch(jobs.poses + 16 * i, jobs.poses_inv + 16 * i);
    if (jobs.cam_stage != nullptr && blockIdx.x == 0) {  // block-uniform; n_poses <= kCamViews <= blockDim.x
        __syncthreads();  // the inverses above were written by this block
        for (int t = threadIdx.x; t < jobs.n_cam_sets * jobs.n_poses; t += blockDim.x) {
            const int set = t / jobs.n_poses, v = t % jobs.n_poses;
            stage_cam(jobs.cam_stage + set * kCamViews + v, jobs.poses_inv + 16 * v, jobs.intrs + 16 * v,
                      jobs.cam_scale[set]);
        }
    }
    if (i >= jobs.total) return;
    int s = 0;
#pragma unroll
    for (int k = 1; k < GENS_MAX_SCALES; ++k)
        if (k < jobs.n_jobs && i >= jobs.job[k].first) s = k;
    const PackJob& j = jobs.job[s];
    const long long local = i - j.first;
    const int h = j.h, w = j.w;
    const long long per = (long long)(h + 1) * w;
    const long long n = local / per;
    const int rem = (int)(local % per), y = rem / w, x = rem % w;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (y < h) {
        const long long hw = (long long)h * w;
        const float* p = j.src + n * 4 * hw + (long long)y * w + x;
        a = make_float4(__ldg(p), __ldg(p + hw), __ldg(p + 2 * hw), __ldg(p + 3 * hw));
        if (x + 1 < w) b = make_float4(__ldg(p + 1), __ldg(p + hw + 1), __ldg(p + 2 * hw + 1), __ldg(p + 3 * hw + 1));
    }
    j.dst[2 * local] = a;
    j.dst[2 * local + 1] = b;
}

// padded channels-last gradient (n, h+1, w+1, 4) -> (n,4,h,w) NCHW (padding rows/cols dropped;
// they only ever receive zero-weight contributions)
__global__ void __launch_bounds__(256)
unpack_maps_kernel(const float4* __restrict__ src, float* __restrict__ dst, int h, int w, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long hw = (long long)h * w;
    const long long n = i / hw;
    const int rem = (int)(i % hw), y = rem / w, x = rem % w;
    const float4 v = __ldg(src + n * (long long)(h + 1) * (w + 1) + (long long)y * (w + 1) + x);
    float* d = dst + n * 4 * hw + rem;
    d[0] = v.x; d[hw] = v.y; d[2 * hw] = v.z; d[3 * hw] = v.w;
}

inline Extent extent(int W, int H) {
    Extent e;
    e.hx = (float)((double)(W - 1) / 2.0);  // python: (width - 1) / 2, then cast to the tensor dtype
    e.hy = (float)((double)(H - 1) / 2.0);
    e.inv_hx = 1.0f / e.hx;  // ATen CUDA div_true: opmath_t(1.0) / scalar
    e.inv_hy = 1.0f / e.hy;
    return e;
}

int g_k1_variant = 0;  // tuning knob (gens_debug_set_variant); 0 = shipped configuration
int g_k1_sched = 3;    // build-level scheduling knob (gens_debug_set_variant(100 + s)), see gens_volume_agg_fwd_multi

inline bool bad_slab(int D, int a0, int a1, int a_base) { return a0 < 0 || a1 > D || a_base < 0 || a_base > a0; }

}  // namespace

extern "C" int gens_debug_set_variant(int variant) {
    if (variant >= 100) g_k1_sched = variant - 100;
    else g_k1_variant = variant;
    return 0;
}

extern "C" int gens_invert_poses(const float* poses, int n, float* poses_inv, void* stream) {
    if (n == 0) return 0;
    GENS_CHECK_ARG(poses && poses_inv && n > 0);
    invert_poses_kernel<<<ceil_div_i(n, 64), 64, 0, (cudaStream_t)stream>>>(poses, poses_inv, n);
    return gens_launch_status();
}

namespace {

// Per-device global staging copy of the constant-bank camera sets + the ring position of the next group.
struct CamStage {
    Cam* dev = nullptr;
    unsigned next_group = 0;
    bool tried = false;
};
std::mutex g_cam_mutex;
CamStage g_cam_stage[64];

// Reserves the next group of GENS_MAX_SCALES camera sets on the current device.  Returns the first set index (or -1)
// and the global staging address of that group.
int reserve_cam_group(Cam** stage_out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    std::lock_guard<std::mutex> lock(g_cam_mutex);
    CamStage& c = g_cam_stage[dev];
    if (!c.tried) {
        c.tried = true;
        if (cudaMalloc(&c.dev, sizeof(Cam) * kCamSets * kCamViews) != cudaSuccess) {
            cudaGetLastError();
            c.dev = nullptr;
        }
    }
    if (!c.dev) return -1;
    const int group = (int)(c.next_group++ % kCamGroups);
    *stage_out = c.dev + (size_t)group * GENS_MAX_SCALES * kCamViews;
    return group * GENS_MAX_SCALES;
}

// staged sets [first_set, first_set + n_sets) -> constant bank, ordered on `st`
int publish_cam_sets(const Cam* stage, int first_set, int n_sets, cudaStream_t st) {
    const size_t bytes = sizeof(Cam) * kCamViews * (size_t)n_sets, offset = sizeof(Cam) * kCamViews * (size_t)first_set;
    return (int)cudaMemcpyToSymbolAsync(c_cam, stage, bytes, offset, cudaMemcpyDeviceToDevice, st);
}

int pack_launch(const float* const* src_nchw, float* const* dst_pairs, const int* h, const int* w, int n_scales, int n,
                const float* poses, float* poses_inv, int n_poses, const float* intrs, Cam* cam_stage,
                const float* cam_scales, int n_cam_sets, cudaStream_t st) {
    GENS_CHECK_ARG(src_nchw && dst_pairs && h && w && n_scales > 0 && n > 0);
    GENS_CHECK_ARG(n_poses == 0 || (poses && poses_inv && n_poses > 0 && n_poses <= 256));
    if (n_scales > GENS_MAX_SCALES) return GENS_E_UNSUPPORTED;
    PackJobs jobs;
    jobs.n_jobs = n_scales;
    jobs.n_maps = n;
    jobs.poses = poses;
    jobs.poses_inv = poses_inv;
    jobs.n_poses = n_poses;
    jobs.intrs = intrs;
    jobs.cam_stage = cam_stage;
    jobs.n_cam_sets = cam_stage ? n_cam_sets : 0;
    for (int i = 0; i < GENS_MAX_SCALES; ++i) jobs.cam_scale[i] = (cam_stage && i < n_cam_sets) ? cam_scales[i] : 1.0f;
    long long total = 0;
    for (int i = 0; i < n_scales; ++i) {
        GENS_CHECK_ARG(src_nchw[i] && dst_pairs[i] && h[i] > 0 && w[i] > 0);
        jobs.job[i].src = src_nchw[i];
        jobs.job[i].dst = reinterpret_cast<float4*>(dst_pairs[i]);
        jobs.job[i].h = h[i];
        jobs.job[i].w = w[i];
        jobs.job[i].first = total;
        total += (long long)n * (h[i] + 1) * w[i];
    }
    jobs.total = total;
    pack_pairs_kernel<<<ceil_div_i(total, 256), 256, 0, st>>>(jobs);
    return gens_launch_status();
}

}  // namespace

extern "C" int gens_pack_feature_maps_multi(const float* const* src_nchw, float* const* dst_pairs, const int* h,
                                            const int* w, int n_scales, int n, const float* poses,
                                            float* poses_inv, int n_poses, void* stream) {
    return pack_launch(src_nchw, dst_pairs, h, w, n_scales, n, poses, poses_inv, n_poses, nullptr, nullptr, nullptr, 0,
                       (cudaStream_t)stream);
}

extern "C" int gens_stage_cameras(const float* w2c, const float* intrs, int nv, const float* k_row_scales, int n_scales,
                                  int* cam_slots, void* stream) {
    GENS_CHECK_ARG(w2c && intrs && k_row_scales && cam_slots && nv > 0 && n_scales > 0);
    for (int i = 0; i < n_scales; ++i) cam_slots[i] = 0;
    if (nv > kCamViews || n_scales > GENS_MAX_SCALES) return 0;  // shared-memory staging covers these
    Cam* stage = nullptr;
    const int first = reserve_cam_group(&stage);
    if (first < 0) return 0;
    CamScales sc;
    for (int i = 0; i < GENS_MAX_SCALES; ++i) sc.s[i] = i < n_scales ? k_row_scales[i] : 1.0f;
    cudaStream_t st = (cudaStream_t)stream;
    stage_cams_kernel<<<1, 64, 0, st>>>(w2c, intrs, nv, n_scales, sc, stage);
    if (int rc = gens_launch_status()) return rc;
    if (int rc = publish_cam_sets(stage, first, n_scales, st)) return rc;
    for (int i = 0; i < n_scales; ++i) cam_slots[i] = first + i + 1;
    return 0;
}

extern "C" int gens_pack_feature_maps(const float* src_nchw, float* dst_pairs, int n, int h, int w, void* stream) {
    return gens_pack_feature_maps_multi(&src_nchw, &dst_pairs, &h, &w, 1, n, nullptr, nullptr, 0, stream);
}

extern "C" int gens_unpack_feature_grads(const float* src_padded_nhwc, float* dst_nchw, int n, int h, int w,
                                         void* stream) {
    GENS_CHECK_ARG(src_padded_nhwc && dst_nchw && n > 0 && h > 0 && w > 0);
    const long long total = (long long)n * h * w;
    unpack_maps_kernel<<<ceil_div_i(total, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(src_padded_nhwc), dst_nchw, h, w, total);
    return gens_launch_status();
}

namespace {

int launch_agg_fwd(const gens_volume_scale_t& sc, int nv, const float* w2c, const float* intrs, int min_vis_view,
                   int div_mode, cudaStream_t st) {
    if (sc.n_peers < 0 || sc.n_peers > GENS_MAX_PEERS) return GENS_E_BADARG;
    for (int i = 0; i < sc.n_peers; ++i)
        if (!(sc.peer_volume[i] && sc.peer_mask[i])) return GENS_E_BADARG;
    if (!(sc.feat_padded && sc.grid && (sc.n_peers > 0 || (sc.volume && sc.mask_volume)))) return GENS_E_BADARG;
    if (!(sc.H > 0 && sc.W > 0 && sc.D > 0) || bad_slab(sc.D, sc.a0, sc.a1, sc.a_base)) return GENS_E_BADARG;
    if ((long long)(sc.H + 1) * (sc.W + 1) * nv >= (1LL << 28)) return GENS_E_UNSUPPORTED;
    if (sc.a1 <= sc.a0) return 0;
    const int D = sc.D, planes = sc.a1 - sc.a0;
    const long long out_off = (long long)(sc.a0 - sc.a_base) * D * D;
    const Extent e = extent(sc.W, sc.H);
    const Pair* feat = reinterpret_cast<const Pair*>(sc.feat_padded);
    const bool recip = div_mode == GENS_DIV_RECIP;
    const dim3 block(32, 8);
    PeerOut peers;
    peers.n = sc.n_peers;
    peers.self = sc.self_peer;
    if (sc.n_peers > 0 && (sc.self_peer < 0 || sc.self_peer >= sc.n_peers)) return GENS_E_BADARG;
    for (int i = 0; i < GENS_MAX_PEERS; ++i) {
        peers.vol[i] = i < sc.n_peers ? sc.peer_volume[i] : nullptr;
        peers.msk[i] = i < sc.n_peers ? sc.peer_mask[i] : nullptr;
    }
    const bool to_peers = sc.n_peers > 0;
#define GENS_AGG_ARGS \
    feat, nv, sc.H, sc.W, w2c, intrs, sc.k_row_scale, sc.grid, D, sc.a0, out_off, sc.channel_stride, min_vis_view, e, \
        sc.volume, sc.mask_volume, peers
#define GENS_LAUNCH_PACKED(PAIRS, MINB, GATHER, ROWS)                                                              \
    do {                                                                                                         \
        const dim3 g(D / (64 * PAIRS), ceil_div_i(D, 8 * ROWS), planes);                                         \
        if (to_peers && recip)                                                                                   \
            volume_agg_packed_kernel<PAIRS, true, MINB, GATHER, ROWS, true><<<g, block, 0, st>>>(GENS_AGG_ARGS);  \
        else if (to_peers)                                                                                       \
            volume_agg_packed_kernel<PAIRS, false, MINB, GATHER, ROWS, true><<<g, block, 0, st>>>(GENS_AGG_ARGS); \
        else if (recip)                                                                                          \
            volume_agg_packed_kernel<PAIRS, true, MINB, GATHER, ROWS><<<g, block, 0, st>>>(GENS_AGG_ARGS);        \
        else                                                                                                     \
            volume_agg_packed_kernel<PAIRS, false, MINB, GATHER, ROWS><<<g, block, 0, st>>>(GENS_AGG_ARGS);       \
    } while (0)
    const int variant = g_k1_variant;
    // Row-group kernel (frustum culling): D a multiple of 64 and a mask threshold that leaves unseen voxels at 0.
    // One destination: D >= 256 (smaller volumes are a single wave of blocks, where the per-block culling prologue
    // costs more than it saves: 128^3 35.6 us packed vs 37.8 us culled).  Multi-GPU (peer stores): every D % 64 == 0,
    // because there the culling also decides which tiles cross NVLink at all.
    // Tuning knob: 10 forces the packed kernel, 20 / 25 the row-group kernel with / without culling at any
    // D % 64 == 0, 11 the previous round's shipped configuration (shared-memory cameras, STG zero fill), 12 / 13
    // only one of the two changes.
    const bool rg = variant == 20 || variant == 25 || ((variant == 0 || (variant >= 11 && variant <= 13)) && D >= 256) ||
                    (to_peers && variant != 10);
    if (rg && D % 64 == 0 && min_vis_view >= 0 && (!to_peers || (sc.a_base == 0 && variant != 25))) {
        const int cam_set = sc.cam_slot - 1;  // public ids are 1-based, 0 = none
        const bool constcam = cam_set >= 0 && cam_set < kCamSets && nv <= kCamViews && variant != 11 && variant != 13;
        const bool zbulk = variant != 11 && variant != 12;
        const dim3 g(D / 64, ceil_div_i(D, 64), to_peers ? D : planes);
#define GENS_RG(RECIP, CULL, CC, ZB, PE)                                                                           \
    volume_agg_rowgroup_kernel<RECIP, 8, CULL, CC, ZB, PE><<<g, block, 0, st>>>(                                     \
        feat, nv, sc.H, sc.W, w2c, intrs, sc.k_row_scale, sc.grid, D, sc.a0, sc.a1, out_off, sc.channel_stride,      \
        min_vis_view, e, sc.volume, sc.mask_volume, cam_set, peers)
#define GENS_RG_R(CULL, CC, ZB, PE)                \
    do {                                           \
        if (recip) GENS_RG(true, CULL, CC, ZB, PE); \
        else GENS_RG(false, CULL, CC, ZB, PE);      \
    } while (0)
        if (to_peers) {
            if (constcam) GENS_RG_R(true, true, true, true);
            else GENS_RG_R(true, false, true, true);
        } else if (variant == 25) {
            GENS_RG_R(false, false, false, false);
        } else if (constcam && zbulk) {
            GENS_RG_R(true, true, true, false);
        } else if (constcam) {
            GENS_RG_R(true, true, false, false);
        } else if (zbulk) {
            GENS_RG_R(true, false, true, false);
        } else {
            GENS_RG_R(true, false, false, false);
        }
#undef GENS_RG_R
#undef GENS_RG
        return gens_launch_status();
    }
    // rows per block: enough to amortise the camera staging, few enough to keep >= ~4 waves of blocks
    const int rows = variant == 1 ? 1 : variant == 2 ? 2 : variant == 4 ? 4 : variant == 8 ? 8 : variant == 16 ? 16
                   : (D >= 256 ? 8 : D >= 128 ? 2 : 1);
    if (D % 64 == 0 && variant != 9) {
        switch (rows) {
            case 16: GENS_LAUNCH_PACKED(1, 4, 1, 16); break;
            case 8: GENS_LAUNCH_PACKED(1, 4, 1, 8); break;
            case 4: GENS_LAUNCH_PACKED(1, 4, 1, 4); break;
            case 2: GENS_LAUNCH_PACKED(1, 4, 1, 2); break;
            default: GENS_LAUNCH_PACKED(1, 4, 1, 1); break;
        }
    } else {
        const dim3 g(ceil_div_i(D, 32), ceil_div_i(D, 8), planes);
        if (recip) volume_agg_scalar_kernel<true><<<g, block, 0, st>>>(GENS_AGG_ARGS);
        else volume_agg_scalar_kernel<false><<<g, block, 0, st>>>(GENS_AGG_ARGS);
    }
#undef GENS_LAUNCH_PACKED
#undef GENS_AGG_ARGS
    return gens_launch_status();
}

}  // namespace

namespace {
// Auxiliary streams per device: the small scales of a build run beside the largest one instead of behind it
// (their launch latencies and tails disappear into the big kernel).  Fork/join with events, so the work is
// still ordered entirely by the caller's stream (and capturable in a CUDA graph).  Measured through bench.py
// (5-scale build, 480x640, 3 views): no fork 259-289 us; big kernel first + one auxiliary stream 248 us; big
// first + one stream per small scale 234 us; small scales first, one stream each, big kernel last 227 us
// (shipped: the block scheduler drains the kernels in launch order, so the short ones must be ahead of the
// 4096-block launch to run inside it rather than in its tail).
constexpr int kAuxStreams = 4;
struct Fork {
    cudaStream_t aux[kAuxStreams] = {};
    cudaEvent_t forked = nullptr, joined[kAuxStreams] = {};
    bool ok = false, tried = false;
};
std::mutex g_fork_mutex;
Fork g_forks[64];

Fork* fork_for_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    Fork& f = g_forks[dev];
    if (!f.tried) {
        f.tried = true;
        f.ok = cudaEventCreateWithFlags(&f.forked, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < kAuxStreams && f.ok; ++i)
            f.ok = cudaStreamCreateWithFlags(&f.aux[i], cudaStreamNonBlocking) == cudaSuccess &&
                   cudaEventCreateWithFlags(&f.joined[i], cudaEventDisableTiming) == cudaSuccess;
        if (!f.ok) cudaGetLastError();
    }
    return f.ok ? &f : nullptr;
}

}  // namespace

extern "C" int gens_volume_agg_fwd_multi(const gens_volume_scale_t* scales, int n_scales, int nv, const float* w2c,
                                         const float* intrs, int min_vis_view, int div_mode, void* stream) {
    GENS_CHECK_ARG(scales && n_scales > 0 && w2c && intrs && nv > 0);
    if (nv > GENS_MAX_VIEWS) return GENS_E_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    int big = 0;
    for (int i = 1; i < n_scales; ++i)
        if (scales[i].D > scales[big].D) big = i;
    std::lock_guard<std::mutex> lock(g_fork_mutex);
    const int sched = g_k1_sched;
    // sched 3 (shipped): the small scales first, round-robin over the auxiliary streams, the big one last on `st`;
    // 0: big first, the rest in order on one auxiliary stream; 1: the rest first on one stream; 2: like 3 but big
    // first; 7: no fork
    Fork* f = n_scales > 1 && g_k1_variant != 7 && sched != 7 ? fork_for_current_device() : nullptr;
    const int n_aux = sched >= 2 ? kAuxStreams : 1;
    if (f) {
        bool ok = cudaEventRecord(f->forked, st) == cudaSuccess;
        for (int i = 0; i < n_aux && ok; ++i) ok = cudaStreamWaitEvent(f->aux[i], f->forked, 0) == cudaSuccess;
        if (!ok) {
            cudaGetLastError();
            f = nullptr;
        }
    }
    int rc = 0;
    if (f) {
        const bool big_first = sched == 0 || sched == 2;
        if (big_first) rc = launch_agg_fwd(scales[big], nv, w2c, intrs, min_vis_view, div_mode, st);
        int k = 0;
        for (int i = 0; i < n_scales && rc == 0; ++i)
            if (i != big) rc = launch_agg_fwd(scales[i], nv, w2c, intrs, min_vis_view, div_mode, f->aux[k++ % n_aux]);
        if (!big_first && rc == 0) rc = launch_agg_fwd(scales[big], nv, w2c, intrs, min_vis_view, div_mode, st);
        // always join, even after an error, so the auxiliary streams never outlive the call's ordering
        for (int i = 0; i < n_aux; ++i)
            if (cudaEventRecord(f->joined[i], f->aux[i]) != cudaSuccess || cudaStreamWaitEvent(st, f->joined[i], 0) != cudaSuccess)
                return rc != 0 ? rc : gens_launch_status();
        return rc;
    }
    for (int i = 0; i < n_scales; ++i) {
        rc = launch_agg_fwd(scales[i], nv, w2c, intrs, min_vis_view, div_mode, st);
        if (rc != 0) return rc;
    }
    return 0;
}

// One host call per build: pack every scale's maps + invert the poses (one launch), then all aggregation
// launches.  Between the two stages the GPU waits only for this function, not for the caller's interpreter.
extern "C" int gens_volume_build(const float* const* src_nchw, float* const* dst_pairs, const int* h, const int* w,
                                 const gens_volume_scale_t* scales, int n_scales, int nv, const float* c2ws,
                                 float* w2c_out, const float* intrs, int min_vis_view, int div_mode, void* stream) {
    GENS_CHECK_ARG(scales && n_scales > 0 && c2ws && w2c_out && intrs && nv > 0);
    if (nv > GENS_MAX_VIEWS || n_scales > GENS_MAX_SCALES) return GENS_E_UNSUPPORTED;
    for (int i = 0; i < n_scales; ++i) GENS_CHECK_ARG(dst_pairs && scales[i].feat_padded == dst_pairs[i]);
    cudaStream_t st = (cudaStream_t)stream;
    // camera sets for the constant bank: staged by block 0 of the pack launch (no extra launch), published by one
    // stream-ordered device-to-device copy
    gens_volume_scale_t local[GENS_MAX_SCALES];
    float cam_scales[GENS_MAX_SCALES];
    for (int i = 0; i < n_scales; ++i) {
        local[i] = scales[i];
        local[i].cam_slot = 0;
        cam_scales[i] = scales[i].k_row_scale;
    }
    Cam* stage = nullptr;
    const int first = nv <= kCamViews && g_k1_variant != 11 && g_k1_variant != 13 ? reserve_cam_group(&stage) : -1;
    if (int rc = pack_launch(src_nchw, dst_pairs, h, w, n_scales, nv, c2ws, w2c_out, nv, intrs, first >= 0 ? stage : nullptr,
                             cam_scales, n_scales, st))
        return rc;
    if (first >= 0) {
        if (int rc = publish_cam_sets(stage, first, n_scales, st)) return rc;
        for (int i = 0; i < n_scales; ++i) local[i].cam_slot = first + i + 1;
    }
    return gens_volume_agg_fwd_multi(local, n_scales, nv, w2c_out, intrs, min_vis_view, div_mode, stream);
}

extern "C" int gens_volume_agg_fwd(const float* feat_padded, int nv, int H, int W, const float* w2c,
                                   const float* intrs, float k_row_scale, const float* grid, int D, int a0, int a1,
                                   int a_base, long long channel_stride, int min_vis_view, int div_mode,
                                   float* volume, float* mask_volume, void* stream) {
    gens_volume_scale_t sc;
    sc.feat_padded = feat_padded; sc.H = H; sc.W = W; sc.D = D; sc.a0 = a0; sc.a1 = a1; sc.a_base = a_base;
    sc.channel_stride = channel_stride; sc.k_row_scale = k_row_scale; sc.grid = grid; sc.volume = volume;
    sc.mask_volume = mask_volume;
    sc.n_peers = 0;
    sc.self_peer = 0;
    sc.cam_slot = 0;
    return gens_volume_agg_fwd_multi(&sc, 1, nv, w2c, intrs, min_vis_view, div_mode, stream);
}

extern "C" int gens_volume_project_debug(int nv, int H, int W, const float* w2c, const float* intrs,
                                         float k_row_scale, const float* grid, int D, int div_mode, int32_t* ix0,
                                         int32_t* iy0, uint8_t* valid, void* stream) {
    GENS_CHECK_ARG(w2c && intrs && grid && ix0 && iy0 && valid && nv > 0 && D > 0 && H > 0 && W > 0);
    if (nv > GENS_MAX_VIEWS) return GENS_E_UNSUPPORTED;
    const Extent e = extent(W, H);
    const dim3 block(32, 8), g(ceil_div_i(D, 32), ceil_div_i(D, 8), D);
    cudaStream_t st = (cudaStream_t)stream;
    if (div_mode == GENS_DIV_RECIP)
        volume_project_debug_kernel<true><<<g, block, 0, st>>>(nv, w2c, intrs, k_row_scale, grid, D, e, ix0, iy0, valid);
    else
        volume_project_debug_kernel<false><<<g, block, 0, st>>>(nv, w2c, intrs, k_row_scale, grid, D, e, ix0, iy0, valid);
    return gens_launch_status();
}

extern "C" int gens_volume_agg_bwd(const float* feat_padded, int nv, int H, int W, const float* w2c,
                                   const float* intrs, float k_row_scale, const float* grid, int D, int a0, int a1,
                                   int a_base, long long channel_stride, int div_mode, const float* grad_volume,
                                   float* grad_feat_padded, void* stream) {
    GENS_CHECK_ARG(feat_padded && w2c && intrs && grid && grad_volume && grad_feat_padded);
    GENS_CHECK_ARG(nv > 0 && H > 0 && W > 0 && D > 0 && !bad_slab(D, a0, a1, a_base));
    if (nv > GENS_MAX_VIEWS || (long long)(H + 1) * (W + 1) * nv >= (1LL << 31)) return GENS_E_UNSUPPORTED;
    if (a1 <= a0) return 0;
    const long long out_off = (long long)(a0 - a_base) * D * D;
    const Extent e = extent(W, H);
    cudaStream_t st = (cudaStream_t)stream;
    const Pair* feat = reinterpret_cast<const Pair*>(feat_padded);
    float4* gfeat = reinterpret_cast<float4*>(grad_feat_padded);
    const dim3 block(32, 8), g(ceil_div_i(D, 32), ceil_div_i(D, 8), a1 - a0);
    if (div_mode == GENS_DIV_RECIP)
        volume_agg_bwd_kernel<true><<<g, block, 0, st>>>(feat, nv, H, W, w2c, intrs, k_row_scale, grid, D, a0, out_off,
                                                          channel_stride, e, grad_volume, gfeat);
    else
        volume_agg_bwd_kernel<false><<<g, block, 0, st>>>(feat, nv, H, W, w2c, intrs, k_row_scale, grid, D, a0, out_off,
                                                           channel_stride, e, grad_volume, gfeat);
    return gens_launch_status();
}

// ---- multi-GPU: scatter an all-gathered, rank-major slab buffer into the final NCDHW tensors --------
namespace {
// recv: per rank r a block of `rank_stride` floats; inside it, at `scale_off`, 9 channel planes of
// (planes, D, D) floats (8 volume channels then the mask).  One thread moves one float4.
__global__ void __launch_bounds__(256)
unpack_slabs_kernel(const float4* __restrict__ recv, long long rank_stride4, long long scale_off4, int D, int planes,
                    float4* __restrict__ volume, float4* __restrict__ mask, long long total4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const long long plane4 = (long long)D * D / 4;             // float4s per tensor plane
    const long long chan4 = plane4 * D;                         // float4s per full channel
    const int ch = (int)(i / chan4);
    const long long rem = i % chan4;
    const int a = (int)(rem / plane4);
    const int r = a / planes;
    const long long src = (long long)r * rank_stride4 + scale_off4 + ((long long)ch * planes + (a - r * planes)) * plane4 +
                          rem % plane4;
    const float4 v = __ldcs(recv + src);
    if (ch < 8) __stcs(volume + i, v);
    else __stcs(mask + (i - 8 * chan4), v);
}
}  // namespace

extern "C" int gens_unpack_slabs(const float* recv, int world, long long rank_stride, long long scale_off, int D,
                                 float* volume, float* mask_volume, void* stream) {
    GENS_CHECK_ARG(recv && volume && mask_volume && world > 0 && D > 0);
    if (D % world != 0 || ((long long)D * D) % 4 != 0 || rank_stride % 4 != 0 || scale_off % 4 != 0)
        return GENS_E_UNSUPPORTED;
    const long long total4 = 9LL * D * D * D / 4;
    unpack_slabs_kernel<<<ceil_div_i(total4, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(recv), rank_stride / 4, scale_off / 4, D, D / world,
        reinterpret_cast<float4*>(volume), reinterpret_cast<float4*>(mask_volume), total4);
    return gens_launch_status();
}
