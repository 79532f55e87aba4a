// K12: marching cubes on the device-resident SDF lattice (sm_100a).
//
// Replaces the last step of ImplicitSurface.extract_geometry (reference models/modules/implicit_surface.py:423:
// `mcubes.marching_cubes(u, threshold)` on a 512^3 host array, after 512 device->host block copies): the lattice
// produced by ImplicitSurface.sdf_grid stays in HBM and the mesh is extracted there; only the mesh goes to the host.
//
// Three passes, all one thread per item, no atomics, deterministic output order:
//   classify   every lattice point: which of ITS three edges (+x, +y, +z) the surface crosses (a vertex is owned by
//              the lower end point of its edge, so vertices are never duplicated), and the number of triangles of the
//              cell whose minimum corner it is (case index = corners with value < iso, table from mc_tables.py)
//   (host)     torch.nonzero compacts the active points / cells (a surface touches O(R^2) of the R^3 cells) and two
//              exclusive scans assign output ranges
//   vertices   per active point: its 1-3 vertices, position = lattice index + t along the edge,
//              t = (iso - f0) / (f1 - f0) in double precision like the reference's C++ meshing code
//   triangles  per active cell: the case's edge triples -> vertex ids, each found by a binary search of the edge's
//              owner point in the sorted list of active points
// HBM-bound: classify reads the lattice once (8 corner reads per point, served by L1/L2) and writes 2 bytes per
// point; the other passes touch only the active set.
#include "common.cuh"

namespace {

__device__ __forceinline__ bool inside_of(float v, float iso) { return v < iso; }

__global__ void __launch_bounds__(256)
mc_classify_kernel(const float* __restrict__ u, int rx, int ry, int rz, float iso, const uint8_t* __restrict__ tri_count,
                   uint8_t* __restrict__ vmask, uint8_t* __restrict__ ntri) {
    const long long n = (long long)rx * ry * rz;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int k = (int)(p % rz), j = (int)((p / rz) % ry), i = (int)(p / ((long long)rz * ry));
    const long long sx = (long long)ry * rz, sy = rz;
    const bool hx = i + 1 < rx, hy = j + 1 < ry, hz = k + 1 < rz;
    const bool c0 = inside_of(__ldg(u + p), iso);
    const bool c1 = hx && inside_of(__ldg(u + p + sx), iso);
    const bool c3 = hy && inside_of(__ldg(u + p + sy), iso);
    const bool c4 = hz && inside_of(__ldg(u + p + 1), iso);
    unsigned m = 0;
    if (hx && c1 != c0) m |= 1u;
    if (hy && c3 != c0) m |= 2u;
    if (hz && c4 != c0) m |= 4u;
    vmask[p] = (uint8_t)m;
    unsigned nt = 0;
    if (hx && hy && hz) {
        const bool c2 = inside_of(__ldg(u + p + sx + sy), iso);
        const bool c5 = inside_of(__ldg(u + p + sx + 1), iso);
        const bool c6 = inside_of(__ldg(u + p + sx + sy + 1), iso);
        const bool c7 = inside_of(__ldg(u + p + sy + 1), iso);
        const unsigned cs = (unsigned)c0 | ((unsigned)c1 << 1) | ((unsigned)c2 << 2) | ((unsigned)c3 << 3) |
                            ((unsigned)c4 << 4) | ((unsigned)c5 << 5) | ((unsigned)c6 << 6) | ((unsigned)c7 << 7);
        nt = __ldg(tri_count + cs);
    }
    ntri[p] = (uint8_t)nt;
}

__global__ void __launch_bounds__(256)
mc_vertices_kernel(const float* __restrict__ u, int rx, int ry, int rz, float iso, const long long* __restrict__ pts,
                   const long long* __restrict__ vbase, const uint8_t* __restrict__ vmask, long long n_pts,
                   double off_x, double off_y, double off_z, double* __restrict__ verts) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pts) return;
    const long long p = pts[t];
    const int k = (int)(p % rz), j = (int)((p / rz) % ry), i = (int)(p / ((long long)rz * ry));
    const long long step[3] = {(long long)ry * rz, (long long)rz, 1};
    const unsigned m = vmask[p];
    const double f0 = (double)__ldg(u + p);
    long long o = vbase[t];
    for (int a = 0; a < 3; ++a) {
        if (!((m >> a) & 1u)) continue;
        const double f1 = (double)__ldg(u + p + step[a]);
        const double w = ((double)iso - f0) / (f1 - f0);
        double pos[3] = {(double)i + off_x, (double)j + off_y, (double)k + off_z};
        pos[a] += w;
        verts[3 * o] = pos[0];
        verts[3 * o + 1] = pos[1];
        verts[3 * o + 2] = pos[2];
        ++o;
    }
}

// index of `key` in the ascending array a[0..n) (it is always present)
__device__ __forceinline__ long long find_sorted(const long long* __restrict__ a, long long n, long long key) {
    long long lo = 0, hi = n - 1;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

struct EdgeOwner {
    int8_t di[12], dj[12], dk[12], axis[12];
};

__global__ void __launch_bounds__(256)
mc_triangles_kernel(const float* __restrict__ u, int rx, int ry, int rz, float iso, const long long* __restrict__ cells,
                    const long long* __restrict__ tbase, long long n_cells, const long long* __restrict__ pts,
                    const long long* __restrict__ vbase, long long n_pts, const uint8_t* __restrict__ vmask,
                    const uint8_t* __restrict__ tri_count, const int8_t* __restrict__ tri_edges, int max_tris,
                    const __grid_constant__ EdgeOwner own, long long vert_offset, long long* __restrict__ tris) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cells) return;
    const long long p = cells[t];
    const long long sx = (long long)ry * rz, sy = rz;
    const long long corner[8] = {p, p + sx, p + sx + sy, p + sy, p + 1, p + sx + 1, p + sx + sy + 1, p + sy + 1};
    unsigned cs = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) cs |= (unsigned)inside_of(__ldg(u + corner[c]), iso) << c;
    const int nt = __ldg(tri_count + cs);
    const int8_t* e = tri_edges + (long long)cs * max_tris * 3;
    long long o = tbase[t];
    for (int q = 0; q < nt; ++q, ++o) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int edge = e[3 * q + r];
            const long long owner = p + own.di[edge] * sx + own.dj[edge] * sy + own.dk[edge];
            const long long at = find_sorted(pts, n_pts, owner);
            const unsigned m = vmask[owner];
            const int slot = __popc(m & ((1u << own.axis[edge]) - 1u));
            tris[3 * o + r] = vert_offset + vbase[at] + slot;
        }
    }
}

}  // namespace

extern "C" int gens_mc_classify(const float* u, int rx, int ry, int rz, float iso, const uint8_t* tri_count,
                                uint8_t* vmask, uint8_t* ntri, void* stream) {
    GENS_CHECK_ARG(u && tri_count && vmask && ntri && rx > 0 && ry > 0 && rz > 0);
    const long long n = (long long)rx * ry * rz;
    mc_classify_kernel<<<ceil_div_i(n, 256), 256, 0, (cudaStream_t)stream>>>(u, rx, ry, rz, iso, tri_count, vmask, ntri);
    return gens_launch_status();
}

extern "C" int gens_mc_vertices(const float* u, int rx, int ry, int rz, float iso, const long long* pts,
                                const long long* vbase, const uint8_t* vmask, long long n_pts, double off_x,
                                double off_y, double off_z, double* verts, void* stream) {
    if (n_pts == 0) return 0;
    GENS_CHECK_ARG(u && pts && vbase && vmask && verts && n_pts > 0);
    mc_vertices_kernel<<<ceil_div_i(n_pts, 256), 256, 0, (cudaStream_t)stream>>>(u, rx, ry, rz, iso, pts, vbase, vmask, n_pts,
                                                                                off_x, off_y, off_z, verts);
    return gens_launch_status();
}

extern "C" int gens_mc_triangles(const float* u, int rx, int ry, int rz, float iso, const long long* cells,
                                 const long long* tbase, long long n_cells, const long long* pts, const long long* vbase,
                                 long long n_pts, const uint8_t* vmask, const uint8_t* tri_count, const int8_t* tri_edges,
                                 int max_tris, const int8_t* edge_owner, long long vert_offset, long long* tris,
                                 void* stream) {
    if (n_cells == 0) return 0;
    GENS_CHECK_ARG(u && cells && tbase && pts && vbase && vmask && tri_count && tri_edges && edge_owner && tris &&
                   n_pts > 0 && max_tris > 0);
    EdgeOwner own;  // host array (12 x 4): di, dj, dk, axis per cube edge
    for (int e = 0; e < 12; ++e) {
        own.di[e] = edge_owner[4 * e];
        own.dj[e] = edge_owner[4 * e + 1];
        own.dk[e] = edge_owner[4 * e + 2];
        own.axis[e] = edge_owner[4 * e + 3];
    }
    mc_triangles_kernel<<<ceil_div_i(n_cells, 256), 256, 0, (cudaStream_t)stream>>>(
        u, rx, ry, rz, iso, cells, tbase, n_cells, pts, vbase, n_pts, vmask, tri_count, tri_edges, max_tris, own,
        vert_offset, tris);
    return gens_launch_status();
}
