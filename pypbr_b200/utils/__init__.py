"""pypbr_b200.utils — mirrors the hot-path part of pypbr/utils/__init__.py."""

from .enums import NormalConvention
from .functions import linear_to_srgb, srgb_to_linear

__all__ = ["NormalConvention", "linear_to_srgb", "srgb_to_linear"]
