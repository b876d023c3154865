"""pypbr_b200.blending — mirrors pypbr/blending/__init__.py."""

from .blending import BlendFactory, BlendMethod, GradientBlend, HeightBlend, MaskBlend, PropertyBlend
from .functional import blend_materials, blend_on_height, blend_on_properties, blend_with_gradient, blend_with_mask

__all__ = [
    "blend_materials", "blend_with_mask", "blend_on_height", "blend_on_properties", "blend_with_gradient",
    "BlendMethod", "MaskBlend", "HeightBlend", "PropertyBlend", "GradientBlend", "BlendFactory",
]
