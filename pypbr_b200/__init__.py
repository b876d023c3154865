"""
pypbr_b200 — B200-native (sm_100a) implementation of PyPBR's per-texel shading hot path behind
PyPBR's Python API: CookTorranceBRDF (forward + backward), the BasecolorMetallicMaterial /
DiffuseSpecularMaterial workflow conversions, blend_materials and the material transforms.  See DESIGN.md.

The arithmetic runs in hand-written CUDA kernels reached through a C ABI (include/pbrcuda.h); there is
no CPU path in this package.
"""

from . import blending, io, materials, models, transforms, utils

__version__ = "0.1.0"

__all__ = ["blending", "io", "materials", "models", "transforms", "utils"]
