"""Patch the B200 hot path into an imported reference tree (prstrive/GenS), leaving models/gens.py and
runner.py untouched.

    import sys; sys.path.insert(0, "/path/to/GenS")
    import gens_b200; gens_b200.install()          # before GenS(...) is constructed
    from models.gens import GenS

What is replaced (and only this): `models.modules.volume.Volume`, `models.modules.implicit_surface.
ImplicitSurface` / `sample_pdf`, and the by-value imports of `lookup_volume`, `lookup_feature`,
`surface_patch_warp` in `models.modules.{projector,sdf_network,implicit_surface}` (the reference imports
them with `from .projector import ...`, so the importing modules' globals are patched too), plus
`models.gens.Volume` / `models.gens.ImplicitSurface`, `compute_LNCC` in `models.losses.{ncc,loss}`, and (regulariser=True)
`RegNetwork` in `models.modules.reg_network` / `models.gens` -- same parameter names; with autograd enabled it runs the
reference's own op sequence, under no_grad on CUDA the K13 kernels.  The reference's own `cuda_gridsample` JIT build is
never triggered: a stub module is registered first, so no nvcc run happens at import time.
"""
from __future__ import annotations

import sys
import types


def install(stub_grid_sample_ext: bool = True, regulariser: bool = True):
    from . import implicit_surface, networks, projector, reg_network, volume

    if stub_grid_sample_ext and "models.modules.grid_sample_cuda.cuda_gridsample" not in sys.modules:
        # the reference JIT-builds its extension at import (cuda_gridsample.py:5); our autograd triple
        # replaces it, so register a thin module exposing the same two entry points instead.
        name = "models.modules.grid_sample_cuda.cuda_gridsample"
        shim = types.ModuleType(name)

        def grid_sample_3d(inp, grid, padding_mode="zeros", align_corners=True):
            if padding_mode != "zeros" or not align_corners:
                raise RuntimeError("gens_b200: only zeros padding / align_corners=True is on the hot path")
            pts = grid.reshape(-1, 3).flip(-1)
            out = projector.lookup_volume(pts, inp, "grad")  # (n, C)
            return out.t().reshape(1, inp.shape[1], *grid.shape[1:4])

        def grid_sample_2d(*_a, **_k):
            raise RuntimeError("gens_b200: cuda_gridsample.grid_sample_2d is dead code in the reference "
                               "(never called by the live model) and is not provided")

        shim.grid_sample_3d, shim.grid_sample_2d = grid_sample_3d, grid_sample_2d
        pkg = types.ModuleType("models.modules.grid_sample_cuda")
        pkg.__path__ = []
        pkg.cuda_gridsample = shim
        sys.modules["models.modules.grid_sample_cuda"] = pkg
        sys.modules[name] = shim

    import importlib
    ref_projector = importlib.import_module("models.modules.projector")
    ref_sdf = importlib.import_module("models.modules.sdf_network")
    ref_surface = importlib.import_module("models.modules.implicit_surface")
    ref_volume = importlib.import_module("models.modules.volume")

    for mod in (ref_projector, ref_sdf, ref_surface):
        for fn in ("lookup_volume", "lookup_feature", "surface_patch_warp"):
            if hasattr(mod, fn):
                setattr(mod, fn, getattr(projector, fn))
    ref_volume.Volume = volume.Volume
    ref_surface.ImplicitSurface = implicit_surface.ImplicitSurface
    ref_surface.sample_pdf = implicit_surface.sample_pdf
    ref_surface.SDFNetwork = networks.SDFNetwork
    ref_surface.BlendingNetwork = networks.BlendingNetwork
    ref_surface.SingleVarianceNetwork = networks.SingleVarianceNetwork
    # the loss-side consumer of the patches (imported by value in models/losses/loss.py:5)
    try:
        from . import losses
        ref_ncc = importlib.import_module("models.losses.ncc")
        ref_loss = importlib.import_module("models.losses.loss")
        ref_ncc.compute_LNCC = losses.compute_LNCC
        ref_loss.compute_LNCC = losses.compute_LNCC
    except ImportError:
        pass
    if regulariser:
        importlib.import_module("models.modules.reg_network").RegNetwork = reg_network.RegNetwork
    if "models.gens" in sys.modules:
        gens = sys.modules["models.gens"]
        gens.Volume, gens.ImplicitSurface = volume.Volume, implicit_surface.ImplicitSurface
        if regulariser:
            gens.RegNetwork = reg_network.RegNetwork
    return {"volume": ref_volume, "implicit_surface": ref_surface, "projector": ref_projector}
