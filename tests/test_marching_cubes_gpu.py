"""K12 (csrc/marching_cubes.cu) through the C ABI against the CPU oracle (oracle/mc_oracle.py; parity with PyMCubes
itself is unpinned, see there): triangle-for-triangle equality with the sequential extraction on small lattices (all 256
cases via white noise), the table-independent vertex set and the manifold properties at 192^3, x-slab extraction =
full extraction, and extract_geometry end to end."""
import numpy as np
import pytest
import torch

from gens_b200.meshing import marching_cubes, marching_cubes_device
from oracle import mc_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _lattice(n, fn):
    g = np.linspace(-1, 1, n)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    return fn(x, y, z).astype(np.float32)


def _sorted(v):
    return v[np.lexsort((v[:, 2], v[:, 1], v[:, 0]))]


@pytest.mark.parametrize("name", ["sphere", "torus", "noise", "ragged"])
def test_kernel_matches_sequential_extraction(cuda_lib, name):
    rng = np.random.default_rng(5)
    if name == "sphere":
        u, iso = _lattice(24, lambda x, y, z: np.sqrt(x * x + y * y + z * z) - 0.62), 0.0
    elif name == "torus":
        u, iso = _lattice(24, lambda x, y, z: np.sqrt((np.sqrt(x * x + y * y) - 0.55) ** 2 + z * z) - 0.23), 0.013
    elif name == "noise":
        u, iso = rng.standard_normal((14, 14, 14)).astype(np.float32), 0.1   # open at the lattice boundary
    else:
        u, iso = rng.standard_normal((9, 17, 5)).astype(np.float32), -0.2    # rx != ry != rz
    v_ref, t_ref = mc_oracle.marching_cubes_numpy(u, iso)
    v, t = marching_cubes(torch.from_numpy(u).to(DEV), iso)
    assert v.dtype == np.float64 and v.shape == v_ref.shape and t.shape == t_ref.shape
    assert np.array_equal(_sorted(v), _sorted(v_ref))
    assert np.array_equal(_sorted(v), mc_oracle.edge_vertices(u, iso))
    assert np.array_equal(mc_oracle.canonical_triangles(v, t), mc_oracle.canonical_triangles(v_ref, t_ref))
    assert t.min() >= 0 and t.max() == len(v) - 1 if len(t) else True


def test_large_lattice_properties_and_slabs(cuda_lib):
    n = 192
    fn = lambda x, y, z: np.sqrt(x * x + y * y + z * z) - 0.7 + 0.05 * np.sin(9 * x) * np.cos(7 * y) * np.sin(5 * z)
    u = _lattice(n, fn)
    ud = torch.from_numpy(u).to(DEV)
    v, t = marching_cubes(ud, 0.0)
    assert np.array_equal(_sorted(v), mc_oracle.edge_vertices(u, 0.0))
    rep = mc_oracle.mesh_report(v, t)
    assert rep["closed"] and rep["oriented"] and rep["degenerate"] == 0 and rep["unused_vertices"] == 0, rep
    assert rep["euler"] == 2, rep
    # every vertex sits on the linearly interpolated surface of its edge: |f(v)| small compared to the cell size
    p = v / (n - 1) * 2 - 1
    assert np.abs(fn(p[:, 0], p[:, 1], p[:, 2])).max() < 2e-3
    # x-slabs with one plane of overlap (the multi-GPU extraction): same triangles as the full lattice
    full = mc_oracle.canonical_triangles(v, t)
    parts = []
    for x0, x1 in ((0, 50), (50, 121), (121, n)):
        hi = min(x1 + 1, n)
        vs, ts = marching_cubes_device(ud[x0:hi], 0.0, index_offset=(float(x0), 0.0, 0.0))
        parts.append(mc_oracle.canonical_triangles(vs.cpu().numpy(), ts.cpu().numpy()))
    merged = np.concatenate(parts, 0)
    assert np.array_equal(merged[np.lexsort(merged.T[::-1])], full)


def test_extract_geometry_on_the_device(cuda_lib):
    """ImplicitSurface.extract_geometry (reference implicit_surface.py:407-427) end to end with the device mesher:
    the mesh is the iso-surface of the lattice sdf_grid returns, in world coordinates."""
    from gens_b200.config import gens_model_conf
    from gens_b200.implicit_surface import ImplicitSurface
    from gens_b200.synthetic import make_reg_volumes
    torch.manual_seed(0)
    surf = ImplicitSurface(gens_model_conf(perturb=0.0)["implicit_surface"]).to(DEV)
    vols = [x.to(DEV) for x in make_reg_volumes([32, 16, 8, 4, 2], seed=4)]
    bmin, bmax = torch.tensor([-1.0, -1, -1], device=DEV), torch.tensor([1.0, 1, 1], device=DEV)
    res = 72
    with torch.no_grad():
        verts, tris = surf.extract_geometry(vols, bmin, bmax, res, 0.0)
        u = surf.sdf_grid(vols, bmin, bmax, res).cpu().numpy()
    assert len(verts) > 1000 and len(tris) > 1000
    idx = (verts + 1.0) / 2.0 * (res - 1.0)
    assert np.allclose(_sorted(idx), mc_oracle.edge_vertices(u, 0.0), atol=1e-9)
    rep = mc_oracle.mesh_report(verts, tris)
    assert rep["oriented"] and rep["degenerate"] == 0 and rep["unused_vertices"] == 0, rep
