"""gens_b200 -- B200-native hot path of GenS (volume construction + ray marching).

Drop-in replacements for the reference's models/modules functions, backed by hand-written
sm_100a kernels behind a C ABI (include/gens_b200.h, libgens_b200.so).  No CPU fallback.
"""
from .implicit_surface import ImplicitSurface, sample_pdf  # noqa: F401
from .install import install  # noqa: F401
from .losses import compute_LNCC  # noqa: F401
from .networks import BlendingNetwork, SDFNetwork, SingleVarianceNetwork  # noqa: F401
from .projector import lookup_feature, lookup_volume, surface_patch_warp  # noqa: F401
from .reg_network import RegNetwork  # noqa: F401
from .volume import Volume  # noqa: F401

__all__ = ["Volume", "ImplicitSurface", "sample_pdf", "SDFNetwork", "BlendingNetwork", "SingleVarianceNetwork",
           "lookup_volume", "lookup_feature", "surface_patch_warp", "compute_LNCC", "RegNetwork", "install"]
