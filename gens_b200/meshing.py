"""Iso-surface extraction on the device -- drop-in for `mcubes.marching_cubes(u, isovalue)`, the last step of the
reference's ImplicitSurface.extract_geometry (models/modules/implicit_surface.py:423; PyMCubes 0.1.4 is pinned in the
reference's requirements.txt and absent from this image).

The lattice produced by `ImplicitSurface.sdf_grid` stays in HBM; K12 (csrc/marching_cubes.cu) classifies the cells,
two scans assign output ranges, and the vertices / triangles are written on the device.  Only the mesh is copied to
the host.  Conventions of the published algorithm as PyMCubes implements it: vertices in lattice-index coordinates
(x = first array axis), linear interpolation t = (iso - f0) / (f1 - f0) along the crossed edge in double precision,
one vertex per crossed edge (shared by the cells around it).  The case table is derived in gens_b200/mc_tables.py.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from .mc_tables import EDGE_OWNER, build_tables

_TABLES = {}


def _tables(device):
    key = str(device)
    if key not in _TABLES:
        count, tris = build_tables()
        _TABLES[key] = (torch.from_numpy(count).to(device), torch.from_numpy(tris.reshape(-1).copy()).to(device),
                        int(tris.shape[1]))
    return _TABLES[key]


_OWNER = (np.array(EDGE_OWNER, dtype=np.int8).reshape(-1)).tobytes()


def marching_cubes_device(u: torch.Tensor, isovalue: float, index_offset=(0.0, 0.0, 0.0),
                          vertex_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """u (rx,ry,rz) float32 CUDA lattice -> (vertices (n,3) float64, triangles (m,3) int64), both on the device.
    `index_offset` is added to the vertex coordinates (a slab of a larger lattice), `vertex_offset` to the triangle
    indices (concatenating the meshes of several slabs)."""
    _lib.require_cuda(u)
    if u.dim() != 3:
        raise RuntimeError(f"marching_cubes expects a 3-D lattice, got {tuple(u.shape)}")
    u = _lib.f32c(u)
    rx, ry, rz = (int(x) for x in u.shape)
    dev = u.device
    L, st = _lib.lib(), _lib.stream_ptr(dev)
    tri_count, tri_edges, max_tris = _tables(dev)
    vmask = torch.empty(rx * ry * rz, device=dev, dtype=torch.uint8)
    ntri = torch.empty_like(vmask)
    _lib.check(L.gens_mc_classify(_lib.ptr(u), rx, ry, rz, float(isovalue), _lib.ptr(tri_count), _lib.ptr(vmask),
                                  _lib.ptr(ntri), st), "gens_mc_classify")
    pts = torch.nonzero(vmask).reshape(-1)          # ascending linear ids of the points that own a vertex
    cells = torch.nonzero(ntri).reshape(-1)
    n_own = ((vmask[pts] & 1) + ((vmask[pts] >> 1) & 1) + ((vmask[pts] >> 2) & 1)).to(torch.int64)
    vbase = torch.cumsum(n_own, 0) - n_own
    n_t = ntri[cells].to(torch.int64)
    tbase = torch.cumsum(n_t, 0) - n_t
    n_verts = int(n_own.sum().item()) if pts.numel() else 0
    n_tris = int(n_t.sum().item()) if cells.numel() else 0
    verts = torch.empty((n_verts, 3), device=dev, dtype=torch.float64)
    tris = torch.empty((n_tris, 3), device=dev, dtype=torch.int64)
    if n_verts:
        _lib.check(L.gens_mc_vertices(_lib.ptr(u), rx, ry, rz, float(isovalue), _lib.ptr(pts), _lib.ptr(vbase),
                                      _lib.ptr(vmask), pts.numel(), float(index_offset[0]), float(index_offset[1]),
                                      float(index_offset[2]), _lib.ptr(verts), st), "gens_mc_vertices")
    if n_tris:
        _lib.check(L.gens_mc_triangles(_lib.ptr(u), rx, ry, rz, float(isovalue), _lib.ptr(cells), _lib.ptr(tbase),
                                       cells.numel(), _lib.ptr(pts), _lib.ptr(vbase), pts.numel(), _lib.ptr(vmask),
                                       _lib.ptr(tri_count), _lib.ptr(tri_edges), max_tris, _OWNER, int(vertex_offset),
                                       _lib.ptr(tris), st), "gens_mc_triangles")
    return verts, tris


def marching_cubes(u, isovalue: float):
    """Same call as `mcubes.marching_cubes(u, isovalue)`: (vertices (n,3) float64, triangles (m,3)) as numpy arrays.
    `u` may be a CUDA tensor (stays on the device) or a host array (uploaded once)."""
    if not isinstance(u, torch.Tensor):
        if not torch.cuda.is_available():
            raise RuntimeError("gens_b200.meshing.marching_cubes needs a CUDA device (no CPU fallback)")
        u = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float32)).cuda()
    v, t = marching_cubes_device(u, isovalue)
    return v.cpu().numpy(), t.cpu().numpy()


def sharded_marching_cubes(u_slab: torch.Tensor, isovalue: float, x0: int, rank: int, world: int, group=None,
                           dst: Optional[int] = 0, mesher=None):
    """Mesh of a lattice that is sharded by x-slabs (parallel.sharded_sdf_grid without the gather): every rank
    meshes the cells of its own slab -- `u_slab` holds its planes PLUS the first plane of the next rank's slab (the
    far face of its last cells; the last rank has none) -- and only the meshes travel: vertices / triangles are
    gathered on rank `dst` (or everywhere with dst=None) with the triangle indices re-based.  Vertices on a slab
    boundary plane are emitted by both neighbours (the mesh is geometrically watertight, not index-welded there)."""
    import torch.distributed as dist
    # `mesher(u, iso, index_offset=...) -> (vertices (n,3) f64, triangles (m,3) i64)`: the device kernels by default (a
    # CPU mesher is injected by the gloo test of this gather logic)
    v, t = (mesher or marching_cubes_device)(u_slab, isovalue, index_offset=(float(x0), 0.0, 0.0))
    if world == 1:
        return v, t
    counts = torch.tensor([v.shape[0], t.shape[0]], device=v.device, dtype=torch.int64)
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    nv = [int(c[0]) for c in all_counts]
    nt = [int(c[1]) for c in all_counts]
    t = t + sum(nv[:rank])
    top_v, top_t = max(max(nv), 1), max(max(nt), 1)
    pv = torch.zeros((top_v, 3), device=v.device, dtype=torch.float64)
    pv[: v.shape[0]] = v
    pt = torch.zeros((top_t, 3), device=v.device, dtype=torch.int64)
    pt[: t.shape[0]] = t
    if dst is None:
        gv = [torch.empty_like(pv) for _ in range(world)]
        gt = [torch.empty_like(pt) for _ in range(world)]
        dist.all_gather(gv, pv, group=group)
        dist.all_gather(gt, pt, group=group)
    else:
        gv = [torch.empty_like(pv) for _ in range(world)] if rank == dst else None
        gt = [torch.empty_like(pt) for _ in range(world)] if rank == dst else None
        dist.gather(pv, gv, dst=dst, group=group)
        dist.gather(pt, gt, dst=dst, group=group)
        if rank != dst:
            return None, None
    return (torch.cat([g[:n] for g, n in zip(gv, nv)], 0), torch.cat([g[:n] for g, n in zip(gt, nt)], 0))
