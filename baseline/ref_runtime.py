"""Loader of the UNMODIFIED reference staged under baseline/_ref/ by baseline/setup_ref.py.

Test / bench infrastructure only -- `gens_b200/` never imports this.  It makes
`import models.gens` work from the staged copy on a box that has neither /root/reference nor a
network, without editing a single reference file:

  * `gridsample_grad2` (the reference's only native component) is NOT re-built: the reference calls
    `torch.utils.cpp_extension.load(name='gridsample_grad2', ...)` at import time
    (models/modules/grid_sample_cuda/cuda_gridsample.py:5); while the reference is being imported that
    one call is answered with the module pre-built by setup_ref.py (baseline/_ref/_ext/).
  * `mcubes` (PyMCubes, absent from the image) -> a stub whose `marching_cubes(u, thr)` records the lattice
    it was handed (`mcubes.last_u`) and returns an empty mesh: the SDF lattice of extract_geometry can
    be compared, the CPU meshing itself is outside the hot path.
  * `torchvision.models.mnasnet1_0(pretrained=True)` needs a download -> random initialisation
    (`weights=None`) while the reference's FeatureNetwork is constructed, as BASELINE.json's configs say
    ("random-init GenS weights").
  * `pyhocon` is absent -> `gens_b200.config.Conf` implements the four ConfigTree getters the model uses.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
TREE = os.path.join(HERE, "_ref", "GenS")
EXT = os.path.join(HERE, "_ref", "_ext", "gridsample_grad2")
EXT_SO = os.path.join(EXT, "gridsample_grad2.so")


def available(need_ext: bool = True) -> bool:
    return os.path.isdir(os.path.join(TREE, "models")) and (os.path.exists(EXT_SO) or not need_ext)


def why_unavailable() -> str:
    if not os.path.isdir(os.path.join(TREE, "models")):
        return "baseline/_ref/GenS is not staged (python baseline/setup_ref.py in the build container)"
    if not os.path.exists(EXT_SO):
        return "baseline/_ref/_ext/gridsample_grad2/gridsample_grad2.so is not built (python baseline/setup_ref.py)"
    return ""


def _mcubes_stub():
    import numpy as np
    m = types.ModuleType("mcubes")
    m.last_u = None

    def marching_cubes(u, threshold):
        m.last_u = u
        return np.zeros((0, 3), np.float64), np.zeros((0, 3), np.int64)

    m.marching_cubes = marching_cubes
    m.__gens_stub__ = True
    return m


@contextlib.contextmanager
def _prebuilt_extension():
    """Answer the reference's import-time cpp_extension.load('gridsample_grad2', ...) with the pre-built module."""
    from torch.utils import cpp_extension
    real = cpp_extension.load

    def load(name, sources, *a, **k):
        if name == "gridsample_grad2" and os.path.exists(EXT_SO):
            return cpp_extension._import_module_from_library(name, EXT, True)
        return real(name, sources, *a, **k)

    cpp_extension.load = load
    try:
        yield
    finally:
        cpp_extension.load = real


@contextlib.contextmanager
def _random_init_mnasnet():
    import torchvision.models as tvm
    real = tvm.mnasnet1_0

    def mnasnet1_0(*a, pretrained=False, **k):
        k.pop("weights", None)
        return real(weights=None, **k)

    tvm.mnasnet1_0 = mnasnet1_0
    try:
        yield
    finally:
        tvm.mnasnet1_0 = real


def purge():
    """Forget the reference's modules (a later load() re-imports them un-patched)."""
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    if TREE in sys.path:
        sys.path.remove(TREE)
    m = sys.modules.get("mcubes")
    if m is not None and getattr(m, "__gens_stub__", False):
        del sys.modules["mcubes"]


def load():
    """Import the staged reference (fresh, un-patched).  Returns a namespace with the modules the hot path
    touches; `ns.GenS(conf)` / `ns.Loss(conf)` construct the reference's own model and loss."""
    if not available():
        raise RuntimeError(why_unavailable())
    purge()
    sys.path.insert(0, TREE)
    if "mcubes" not in sys.modules:
        sys.modules["mcubes"] = _mcubes_stub()
    import importlib
    with _prebuilt_extension():
        gens = importlib.import_module("models.gens")
    ns = types.SimpleNamespace()
    ns.gens = gens
    ns.volume = importlib.import_module("models.modules.volume")
    ns.projector = importlib.import_module("models.modules.projector")
    ns.implicit_surface = importlib.import_module("models.modules.implicit_surface")
    ns.sdf_network = importlib.import_module("models.modules.sdf_network")
    ns.cuda_gridsample = importlib.import_module("models.modules.grid_sample_cuda.cuda_gridsample")
    ns.loss = importlib.import_module("models.losses.loss")
    ns.ncc = importlib.import_module("models.losses.ncc")
    ns.mcubes = sys.modules["mcubes"]

    def make_gens(conf):
        with _random_init_mnasnet():
            return sys.modules["models.gens"].GenS(conf)

    ns.GenS = make_gens
    ns.Loss = ns.loss.Loss
    return ns


# train.loss block of confs/gens.conf:47-58
LOSS_CONF = dict(color_weight=1.0, sparse_scale_factor=100.0, sparse_weight=0.02, igr_weight=0.1, mfc_weight=1.0,
                 smooth_weight=0.0001, tv_weight=0.0001, depth_weight=0.0, pseudo_sdf_weight=1.0,
                 pseudo_depth_weight=0.05)
