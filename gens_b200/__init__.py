"""gens_b200 -- B200-native hot path of GenS (volume construction + ray marching).

Drop-in replacements for the reference's models/modules functions, backed by hand-written
sm_100a kernels behind a C ABI (include/gens_b200.h, libgens_b200.so).  No CPU fallback.
"""
from .volume import Volume  # noqa: F401

__all__ = ["Volume"]
