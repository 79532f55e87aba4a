"""Minimal pyhocon-ConfigTree look-alike plus the hot-path constants of the reference's confs/gens.conf
(:64-98), so the modules can be built without pyhocon (absent from this image).  A real ConfigTree
works just as well: only get_int / get_float / get_list / get_bool / [] are used."""
from __future__ import annotations


class Conf(dict):
    def _get(self, key):
        node = self
        for part in key.split("."):
            node = node[part]
        return node

    def get_list(self, key, default=None):
        try:
            return list(self._get(key))
        except KeyError:
            if default is None:
                raise
            return default

    def get_int(self, key, default=None):
        try:
            return int(self._get(key))
        except KeyError:
            if default is None:
                raise
            return default

    def get_float(self, key, default=None):
        try:
            return float(self._get(key))
        except KeyError:
            if default is None:
                raise
            return default

    def get_bool(self, key, default=None):
        try:
            return bool(self._get(key))
        except KeyError:
            if default is None:
                raise
            return default

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        return Conf(v) if isinstance(v, dict) and not isinstance(v, Conf) else v


def gens_model_conf(perturb: float = 1.0) -> Conf:
    """model { ... } block of confs/gens.conf."""
    return Conf({
        "feature_network": {"d_out": [4, 4, 4, 4, 4]},
        "volume": {"volume_dims": [256, 128, 64, 32, 16]},
        "reg_network": {"d_voluem": [8, 8, 8, 8, 8], "d_out": [4, 4, 4, 4, 4], "d_base": 8},
        "implicit_surface": {
            "sdf_network": dict(d_out=129, d_in=3, d_hidden=128, n_layers=6, skip_in=[3], multires=4, bias=0.5,
                                scale=1.0, geometric_init=True, weight_norm=True, feat_channels=20),
            "color_network": dict(d_feature=20),
            "variance_network": dict(init_val=0.3),
            "render": dict(n_samples=64, n_importance=64, up_sample_steps=4, perturb=perturb),
        },
    })
