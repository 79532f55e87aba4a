"""Case table of the marching-cubes kernels (gens_b200/mc_tables.py) checked on the CPU: every case uses exactly its
crossed edges, and a sequential extraction with the table (oracle/mc_oracle.py) yields closed, consistently outward
oriented 2-manifolds with the right Euler characteristic and the table-independent vertex set."""
import numpy as np

from gens_b200.mc_tables import EDGES, build_tables
from oracle import mc_oracle


def _lattice(n, fn):
    g = np.linspace(-1, 1, n)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    return fn(x, y, z).astype(np.float32)


def test_every_case_triangulates_exactly_its_crossed_edges():
    count, table = build_tables()
    assert count[0] == 0 and count[255] == 0 and int(count.max()) == 5
    for case in range(256):
        inside = [(case >> c) & 1 for c in range(8)]
        crossed = {e for e, (a, b) in enumerate(EDGES) if inside[a] != inside[b]}
        used = {int(e) for e in table[case, : count[case]].reshape(-1)}
        assert used == crossed, case
        assert np.all(table[case, count[case]:] == -1)
        # a loop of k vertices gives k - 2 triangles: sum over loops = crossed - 2 * loops
        assert (len(crossed) - int(count[case])) % 2 == 0


def test_sphere_and_torus_are_closed_oriented_manifolds():
    for name, fn, euler in (
            ("sphere", lambda x, y, z: np.sqrt(x * x + y * y + z * z) - 0.62, 2),
            ("torus", lambda x, y, z: np.sqrt((np.sqrt(x * x + y * y) - 0.55) ** 2 + z * z) - 0.23, 0),
            ("two spheres", lambda x, y, z: np.minimum(np.sqrt((x - .45) ** 2 + y * y + z * z),
                                                       np.sqrt((x + .45) ** 2 + y * y + z * z)) - 0.3, 4)):
        u = _lattice(24, fn)
        v, t = mc_oracle.marching_cubes_numpy(u, 0.0)
        rep = mc_oracle.mesh_report(v, t)
        assert rep["closed"] and rep["oriented"] and rep["degenerate"] == 0 and rep["unused_vertices"] == 0, (name, rep)
        assert rep["euler"] == euler, (name, rep)
        assert np.array_equal(v[np.lexsort((v[:, 2], v[:, 1], v[:, 0]))], mc_oracle.edge_vertices(u, 0.0)), name
        # outward: inside = value < iso, triangle normals point towards growing values
        p = v[t]
        n = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
        c = p.mean(1) / 23.0 * 2 - 1
        eps = 1e-3
        grad = np.stack([fn(c[:, 0] + eps, c[:, 1], c[:, 2]) - fn(c[:, 0] - eps, c[:, 1], c[:, 2]),
                         fn(c[:, 0], c[:, 1] + eps, c[:, 2]) - fn(c[:, 0], c[:, 1] - eps, c[:, 2]),
                         fn(c[:, 0], c[:, 1], c[:, 2] + eps) - fn(c[:, 0], c[:, 1], c[:, 2] - eps)], 1)
        assert np.mean((n * grad).sum(1) > 0) > 0.999, name


def test_ambiguous_faces_stay_watertight_on_noise():
    """White noise hits every one of the 256 cases, ambiguous faces included: no cracks between cells (the lattice
    boundary is the only place where edges may be open)."""
    rng = np.random.default_rng(3)
    u = rng.standard_normal((12, 12, 12)).astype(np.float32)
    u = np.pad(u, 1, constant_values=5.0)  # everything closes inside the padded lattice
    v, t = mc_oracle.marching_cubes_numpy(u, 0.0)
    rep = mc_oracle.mesh_report(v, t)
    assert rep["closed"] and rep["oriented"] and rep["degenerate"] == 0, rep
    assert np.array_equal(v[np.lexsort((v[:, 2], v[:, 1], v[:, 0]))], mc_oracle.edge_vertices(u, 0.0))
