"""pypbr_b200.utils — mirrors pypbr/utils/__init__.py."""
from .enums import NormalConvention
from .functions import (compute_height_from_normal, compute_normal_from_height, invert_normal, linear_to_srgb, rotate_normals,
                        srgb_to_linear)

__all__ = ["NormalConvention", "linear_to_srgb", "srgb_to_linear", "rotate_normals", "invert_normal",
           "compute_normal_from_height", "compute_height_from_normal"]
