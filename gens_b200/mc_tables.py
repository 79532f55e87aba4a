"""Marching-cubes case tables, derived (not typed in) from the cube's geometry.

The reference meshes its SDF lattice with `mcubes.marching_cubes` (PyMCubes 0.1.4, requirements.txt; the package is a
third-party dependency that is absent from the reference tree and from this image), i.e. Lorensen & Cline's marching
cubes with the corner / edge numbering popularised by P. Bourke:

    corners  0:(0,0,0) 1:(1,0,0) 2:(1,1,0) 3:(0,1,0) 4:(0,0,1) 5:(1,0,1) 6:(1,1,1) 7:(0,1,1)      (x, y, z)
    edges    0:0-1 1:1-2 2:2-3 3:3-0 4:4-5 5:5-6 6:6-7 7:7-4 8:0-4 9:1-5 10:2-6 11:3-7

Bit c of the case index is set when corner c is INSIDE (value < isovalue).  For every one of the 256 cases the
iso-surface patches are built the way the published algorithm defines them: on each cube face the crossed edges are
joined by segments (a face with four crossed edges is the ambiguous case: each inside corner is cut off on its own,
the same rule on both cubes that share the face, so neighbouring cells always agree and the mesh is watertight), the
segments close into loops over the cube, every loop is oriented so that its normal points from inside to outside and
split into triangles without ever drawing a diagonal inside a cube face (see _split_loop).  Vertices live on lattice edges, so the vertex SET of the mesh is independent of the table; the
table only decides how a cell's loop is split into triangles.
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np

CORNERS = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], np.int64)
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
# per edge: the lattice point that owns it as (corner offset, axis) -- a point owns its +x, +y, +z edges
EDGE_OWNER = []
for _a, _b in EDGES:
    _lo = CORNERS[_a] if tuple(CORNERS[_a]) <= tuple(CORNERS[_b]) else CORNERS[_b]
    _axis = int(np.nonzero(CORNERS[_a] != CORNERS[_b])[0][0])
    EDGE_OWNER.append((int(_lo[0]), int(_lo[1]), int(_lo[2]), _axis))
FACES = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7)]  # cyclic corner order
_EDGE_ID = {frozenset(e): i for i, e in enumerate(EDGES)}
MAX_TRIS = 5  # like the classic table: no case needs more than five triangles


def _case_loops(case: int):
    inside = [(case >> c) & 1 for c in range(8)]
    links = {}  # crossed edge -> the crossed edges it is joined to (exactly two, one per adjacent face)

    def join(e1, e2):
        links.setdefault(e1, []).append(e2)
        links.setdefault(e2, []).append(e1)

    for face in FACES:
        fe = [_EDGE_ID[frozenset((face[i], face[(i + 1) % 4]))] for i in range(4)]       # edge i joins corner i, i+1
        crossed = [i for i in range(4) if inside[face[i]] != inside[face[(i + 1) % 4]]]
        if len(crossed) == 2:
            join(fe[crossed[0]], fe[crossed[1]])
        elif len(crossed) == 4:
            for i in range(4):  # cut off every inside corner: join the two face edges that meet at it
                if inside[face[i]]:
                    join(fe[(i - 1) % 4], fe[i])
    loops, seen = [], set()
    for start in sorted(links):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            nxt = [e for e in links[cur] if e != prev]
            # a two-edge "loop" cannot occur on a cube; pick the first unvisited neighbour, close when back at start
            step = next((e for e in nxt if e not in seen), None)
            if step is None:
                break
            loop.append(step)
            seen.add(step)
            prev, cur = cur, step
        loops.append(loop)
    return loops, inside


def _orient(loop, inside):
    mid = [(CORNERS[EDGES[e][0]] + CORNERS[EDGES[e][1]]) / 2.0 for e in loop]
    normal = np.zeros(3)
    for i in range(len(mid)):  # Newell
        p, q = mid[i], mid[(i + 1) % len(mid)]
        normal += np.cross(p, q)
    outward = np.zeros(3)
    for e in loop:
        a, b = EDGES[e]
        outward += (CORNERS[b] - CORNERS[a]) * (1.0 if inside[a] else -1.0)  # from the inside corner to the outside one
    return loop if float(normal @ outward) > 0 else loop[::-1]


def _same_face(e1: int, e2: int) -> bool:
    """Do two cube edges lie on a common face?"""
    c1, c2 = set(EDGES[e1]), set(EDGES[e2])
    return any(c1 <= set(f) and c2 <= set(f) for f in FACES)


def _triangulations(poly):
    """All triangulations of a convex polygon given as a vertex list, each a list of triples (orientation kept)."""
    if len(poly) < 3:
        return [[]]
    if len(poly) == 3:
        return [[tuple(poly)]]
    out = []
    a, b = poly[0], poly[-1]
    for m in range(1, len(poly) - 1):  # the triangle on the edge (last, first) has apex poly[m]
        for left in _triangulations(poly[: m + 1]):
            for right in _triangulations(poly[m:]):
                out.append(left + [(a, poly[m], b)] + right)
    return out


def _split_loop(loop):
    """Triangles of one loop.  A diagonal whose two vertices lie on a common cube face would lie IN that face, where
    the neighbouring cell may draw the very same diagonal (an edge shared by four triangles) or a segment that crosses
    it: among all triangulations of the loop the first one (fans first) without such a diagonal is taken."""
    n = len(loop)
    if n == 3:
        return [tuple(loop)]
    boundary = {frozenset((loop[i], loop[(i + 1) % n])) for i in range(n)}
    best, best_bad = None, None
    fans = [[(loop[s], loop[(s + i) % n], loop[(s + i + 1) % n]) for i in range(1, n - 1)] for s in range(n)]
    for cand in fans + _triangulations(list(loop)):
        diags = set()
        for t in cand:
            for x, y in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
                if frozenset((x, y)) not in boundary:
                    diags.add(frozenset((x, y)))
        bad = sum(1 for d in diags if _same_face(*tuple(d)))
        if best is None or bad < best_bad:
            best, best_bad = cand, bad
        if bad == 0:
            break
    assert best_bad == 0, (loop, best_bad)
    return best


@lru_cache(maxsize=1)
def build_tables():
    """(tri_count (256,) uint8, tri_edges (256, MAX_TRIS, 3) int8 padded with -1)."""
    count = np.zeros(256, np.uint8)
    tris = -np.ones((256, MAX_TRIS, 3), np.int8)
    for case in range(256):
        loops, inside = _case_loops(case)
        n = 0
        for loop in loops:
            assert len(loop) >= 3, (case, loop)
            loop = _orient(loop, inside)
            for tri in _split_loop(loop):
                tris[case, n] = tri
                n += 1
        count[case] = n
        crossed = sum(1 for a, b in EDGES if inside[a] != inside[b])
        assert sum(len(l) for l in loops) == crossed, case
    assert int(count.max()) <= MAX_TRIS
    return count, tris
