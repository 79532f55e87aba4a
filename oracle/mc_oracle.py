"""CPU oracle of the iso-surface extraction (K12) -- TEST INFRASTRUCTURE, never imported by gens_b200/.

PARITY UNPINNED against the reference's mesher: the reference calls `mcubes.marching_cubes` (PyMCubes 0.1.4, pinned in
its requirements.txt), a third-party package that is absent from /root/reference and from this image, and the
reference holds no mesh fixture.  What is restated here is the published algorithm (Lorensen & Cline, with P. Bourke's
numbering as PyMCubes uses it) at the reference's call site, models/modules/implicit_surface.py:423:
  * `edge_vertices`: the vertex SET, independent of any case table -- one vertex per lattice edge whose end points
    lie on different sides of the isovalue, at index + (iso - f0) / (f1 - f0) along the edge (double precision);
  * `marching_cubes_numpy`: a plain sequential cell loop over the case table of gens_b200/mc_tables.py, for
    triangle-for-triangle comparison with the CUDA kernels;
  * `mesh_report`: the properties a correct extraction must have whatever the table -- closed 2-manifold (every edge
    shared by exactly two triangles, traversed in opposite directions), Euler characteristic, outward orientation.
"""
from __future__ import annotations

import numpy as np


def edge_vertices(u: np.ndarray, iso: float) -> np.ndarray:
    """(n,3) float64 vertex positions in lattice-index coordinates, sorted lexicographically."""
    u = np.asarray(u, dtype=np.float32)
    inside = u < np.float32(iso)
    out = []
    for axis in range(3):
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[axis], hi[axis] = slice(0, -1), slice(1, None)
        cross = inside[tuple(lo)] != inside[tuple(hi)]
        idx = np.argwhere(cross)
        f0 = u[tuple(lo)][cross].astype(np.float64)
        f1 = u[tuple(hi)][cross].astype(np.float64)
        pos = idx.astype(np.float64)
        pos[:, axis] += (np.float64(np.float32(iso)) - f0) / (f1 - f0)
        out.append(pos)
    v = np.concatenate(out, 0) if out else np.zeros((0, 3))
    return v[np.lexsort((v[:, 2], v[:, 1], v[:, 0]))]


def marching_cubes_numpy(u: np.ndarray, iso: float):
    """Sequential marching cubes with the case table of gens_b200.mc_tables: (vertices (n,3) f64, triangles (m,3))."""
    from gens_b200.mc_tables import CORNERS, EDGES, build_tables
    count, table = build_tables()
    u = np.asarray(u, dtype=np.float32)
    rx, ry, rz = u.shape
    iso32 = np.float32(iso)
    verts, index, tris = [], {}, []

    def vertex(p, q):
        key = (p, q) if p <= q else (q, p)
        if key not in index:
            a, b = key
            f0, f1 = np.float64(u[a]), np.float64(u[b])
            t = (np.float64(iso32) - f0) / (f1 - f0)
            index[key] = len(verts)
            verts.append(np.array(a, np.float64) + t * (np.array(b, np.float64) - np.array(a, np.float64)))
        return index[key]

    for i in range(rx - 1):
        for j in range(ry - 1):
            for k in range(rz - 1):
                corner = [(i + int(c[0]), j + int(c[1]), k + int(c[2])) for c in CORNERS]
                case = sum((1 << c) for c in range(8) if u[corner[c]] < iso32)
                for q in range(int(count[case])):
                    tris.append([vertex(corner[EDGES[e][0]], corner[EDGES[e][1]]) for e in table[case, q]])
    return (np.array(verts, np.float64).reshape(-1, 3), np.array(tris, np.int64).reshape(-1, 3))


def canonical_triangles(verts: np.ndarray, tris: np.ndarray) -> np.ndarray:
    """Triangles as coordinate triples, each rotated so that its smallest vertex comes first (orientation kept), the
    list sorted: two meshes are the same surface triangle for triangle iff these arrays are equal."""
    if len(tris) == 0:
        return np.zeros((0, 9))
    p = verts[tris]                                   # (m,3,3)
    keys = p[:, :, 0] * 1e12 + p[:, :, 1] * 1e6 + p[:, :, 2]
    first = np.argmin(keys, axis=1)
    rolled = np.stack([np.roll(p[i], -first[i], axis=0) for i in range(len(p))]).reshape(len(p), 9)
    order = np.lexsort(rolled.T[::-1])
    return rolled[order]


def mesh_report(verts: np.ndarray, tris: np.ndarray) -> dict:
    """Topology of a triangle mesh: is it a closed, consistently oriented 2-manifold; V - E + F."""
    directed = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]], 0)
    und = np.sort(directed, axis=1)
    uniq, inv, counts = np.unique(und, axis=0, return_inverse=True, return_counts=True)
    closed = bool(np.all(counts == 2))
    # consistent orientation: the two triangles of an edge traverse it in opposite directions
    sign = np.where(directed[:, 0] < directed[:, 1], 1, -1)
    balance = np.zeros(len(uniq), np.int64)
    np.add.at(balance, inv.reshape(-1), sign)
    used = np.unique(tris)
    return {"closed": closed, "oriented": bool(np.all(balance == 0)), "euler": int(len(used) - len(uniq) + len(tris)),
            "degenerate": int(np.sum((tris[:, 0] == tris[:, 1]) | (tris[:, 1] == tris[:, 2]) | (tris[:, 0] == tris[:, 2]))),
            "unused_vertices": int(len(verts) - len(used))}
