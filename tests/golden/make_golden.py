"""Generate golden input/output vectors by running the UNMODIFIED reference on CPU.

Runs only in the build container (needs /root/reference); the .npz fixtures it writes are
committed so that tests on the GPU box never touch the reference tree.

    python tests/golden/make_golden.py [volume] [render] ...

Fixtures
  volume_agg.npz   Volume.agg_mean_var (reference models/modules/volume.py:13-63) on a
                   96x128, 3-view scene, volume_dims [32,16,8,4,2]; also per-view validity
                   masks (single-view runs with min_vis_view=0) and the normalised grid the
                   reference handed to F.grid_sample (captured, to derive corner indices).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from gens_b200.synthetic import make_scene  # noqa: E402


class Conf(dict):
    """Minimal stand-in for the pyhocon ConfigTree the reference reads (pyhocon is absent)."""

    def _get(self, key):
        node = self
        for part in key.split("."):
            node = node[part]
        return node

    def get_list(self, key):
        return list(self._get(key))

    def get_int(self, key):
        return int(self._get(key))

    def get_float(self, key):
        return float(self._get(key))

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        return Conf(v) if isinstance(v, dict) and not isinstance(v, Conf) else v


def load_ref_module(rel_path: str, name: str):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel_path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden_volume():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    volume_mod = load_ref_module("models/modules/volume.py", "ref_volume")
    dims = [32, 16, 8, 4, 2]
    vol = volume_mod.Volume(Conf(volume_dims=dims))
    scene = make_scene(96, 128, 3, seed=0)

    captured = []
    real_grid_sample = F.grid_sample

    def spy(inp, grid, *a, **k):
        captured.append(grid.detach().clone())
        return real_grid_sample(inp, grid, *a, **k)

    volume_mod.F.grid_sample = spy
    try:
        volumes, masks = vol.agg_mean_var(scene.features, scene.intrs, scene.c2ws)
    finally:
        volume_mod.F.grid_sample = real_grid_sample
    out = {
        "dims": np.array(dims), "hw": np.array(scene.hw),
        "intrs": scene.intrs.numpy(), "c2ws": scene.c2ws.numpy(),
    }
    for i, d in enumerate(dims):
        out[f"feat{i}"] = scene.features[i].numpy()
        out[f"volume{i}"] = volumes[i][0].numpy()
        out[f"mask{i}"] = masks[i][0, 0].numpy()
        out[f"grid{i}"] = captured[i][:, 0].numpy()  # (nv, D^3, 2) normalised x,y
        # exact visible-view count from the strict threshold at 0,1,2
        cnt = np.zeros((d, d, d), np.int32)
        for t in range(3):
            _, mk = vol.agg_mean_var(scene.features, scene.intrs, scene.c2ws, min_vis_view=t)
            cnt += mk[i][0, 0].numpy().astype(np.int32)
        out[f"count{i}"] = cnt
        pv = []
        for v in range(3):
            _, mk = vol.agg_mean_var([f[v:v + 1] for f in scene.features], scene.intrs[v:v + 1],
                                     scene.c2ws[v:v + 1], min_vis_view=0)
            pv.append(mk[i][0, 0].numpy().astype(np.uint8))
        out[f"viewmask{i}"] = np.stack(pv)
        assert (out[f"viewmask{i}"].sum(0) == cnt).all(), "per-view masks inconsistent with counts"
    np.savez_compressed(os.path.join(HERE, "volume_agg.npz"), **out)
    print("volume_agg.npz written; mask fill per scale:",
          [float(out[f"mask{i}"].mean()) for i in range(len(dims))])


if __name__ == "__main__":
    which = sys.argv[1:] or ["volume"]
    if "volume" in which:
        golden_volume()
