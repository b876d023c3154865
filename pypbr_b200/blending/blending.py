"""
pypbr_b200.blending.blending — argument-holding callables over the functional API.
Mirrors pypbr/blending/blending.py:50-214 (MaskBlend / HeightBlend / PropertyBlend / GradientBlend /
BlendFactory); no arithmetic of its own.
"""

from __future__ import annotations

from abc import ABC, abstractmethod

import torch

from ..materials import MaterialBase
from . import functional as BF


class BlendMethod(ABC):
    @abstractmethod
    def __call__(self, material1: MaterialBase, material2: MaterialBase) -> MaterialBase:
        pass


class MaskBlend(BlendMethod):
    """Blend with a fixed mask of shape [1, H, W] or [H, W]."""

    def __init__(self, mask: torch.Tensor):
        if mask.dim() == 2:
            self.mask = mask.unsqueeze(0)
        elif mask.dim() == 3 and mask.size(0) == 1:
            self.mask = mask
        else:
            raise ValueError("Mask must have shape [1, H, W] or [H, W].")

    def __call__(self, material1, material2):
        return BF.blend_with_mask(material1, material2, self.mask)


class HeightBlend(BlendMethod):
    def __init__(self, blend_width: float = 0.1, shift: float = 0.0):
        self.blend_width = blend_width
        self.shift = shift

    def __call__(self, material1, material2):
        return BF.blend_on_height(material1, material2, self.blend_width, self.shift)


class PropertyBlend(BlendMethod):
    def __init__(self, property_name: str = "metallic", blend_width: float = 0.1):
        self.property_name = property_name
        self.blend_width = blend_width

    def __call__(self, material1, material2):
        return BF.blend_on_properties(material1, material2, self.property_name, self.blend_width)


class GradientBlend(BlendMethod):
    def __init__(self, direction: str = "horizontal"):
        if direction not in ["horizontal", "vertical"]:
            raise ValueError("Direction must be 'horizontal' or 'vertical'.")
        self.direction = direction

    def __call__(self, material1, material2):
        return BF.blend_with_gradient(material1, material2, self.direction)


class BlendFactory:
    @staticmethod
    def get_blend_method(method_name: str, **kwargs) -> BlendMethod:
        name = method_name.lower()
        if name == "mask":
            return MaskBlend(**kwargs)
        if name == "height":
            return HeightBlend(**kwargs)
        if name == "properties":
            return PropertyBlend(**kwargs)
        if name == "gradient":
            return GradientBlend(**kwargs)
        raise ValueError(f"Unknown blending method: {method_name}")
