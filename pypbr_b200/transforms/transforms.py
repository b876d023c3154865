"""
pypbr_b200.transforms.transforms — the callable transform classes of pypbr/transforms/transforms.py:33-386: argument
holders in front of `pypbr_b200.transforms.functional`, plus `Compose`.  Constructor and call signatures are the
reference's (including `RandomHorizontalFlip()(material, p)`, which takes its probability at call time, :297-312).
"""

from __future__ import annotations

from typing import Callable, List, Tuple

from ..materials import MaterialBase
from . import functional as _F

__all__ = ["Compose", "Resize", "RandomResize", "Crop", "CenterCrop", "RandomCrop", "Tile", "Rotate", "RandomRotate",
           "FlipHorizontal", "FlipVertical", "RandomHorizontalFlip", "RandomVerticalFlip", "Roll", "InvertNormal",
           "AdjustNormalStrength", "ToLinear", "ToSrgb"]

_PADDING_MODES = ("constant", "circular")


class Compose:
    """Apply the given transforms one after the other (transforms.py:33-59)."""

    def __init__(self, transforms: List[Callable[[MaterialBase], MaterialBase]]):
        self.transforms = transforms

    def __call__(self, material: MaterialBase) -> MaterialBase:
        for step in self.transforms:
            material = step(material)
        return material


class _Bound:
    """A functional transform with its keyword arguments stored as attributes of the same names."""

    _fn: Callable = None
    _fields: Tuple[str, ...] = ()

    def _kwargs(self):
        return {name: getattr(self, name) for name in self._fields}

    def __call__(self, material: MaterialBase) -> MaterialBase:
        return type(self)._fn(material, **self._kwargs())

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(f'{k}={v!r}' for k, v in self._kwargs().items())})"


class Resize(_Bound):
    _fn, _fields = staticmethod(_F.resize), ("size", "antialias")

    def __init__(self, size: Tuple[int, int], antialias: bool = True):
        self.size, self.antialias = size, antialias


class RandomResize(_Bound):
    _fn, _fields = staticmethod(_F.random_resize), ("min_size", "max_size", "antialias")

    def __init__(self, min_size: int, max_size: int, antialias: bool = True):
        self.min_size, self.max_size, self.antialias = min_size, max_size, antialias


class Crop(_Bound):
    _fn, _fields = staticmethod(_F.crop), ("top", "left", "height", "width")

    def __init__(self, top: int, left: int, height: int, width: int):
        self.top, self.left, self.height, self.width = top, left, height, width


class _SizedCrop(_Bound):
    def __init__(self, height: int, width: int):
        self.height, self.width = height, width

    def _kwargs(self):
        return {"crop_size": (self.height, self.width)}


class CenterCrop(_SizedCrop):
    _fn = staticmethod(_F.center_crop)


class RandomCrop(_SizedCrop):
    _fn = staticmethod(_F.random_crop)


class Tile(_Bound):
    _fn, _fields = staticmethod(_F.tile), ("num_tiles",)

    def __init__(self, num_tiles: int):
        self.num_tiles = num_tiles


class Rotate(_Bound):
    _fn, _fields = staticmethod(_F.rotate), ("angle", "expand", "padding_mode")

    def __init__(self, angle: float, expand: bool = False, padding_mode: str = "constant"):
        assert padding_mode in _PADDING_MODES, "Invalid padding mode."
        self.angle, self.expand, self.padding_mode = angle, expand, padding_mode


class RandomRotate(_Bound):
    _fn, _fields = staticmethod(_F.random_rotate), ("min_angle", "max_angle", "expand", "padding_mode")

    def __init__(self, min_angle: float = 0.0, max_angle: float = 360.0, expand: bool = False, padding_mode: str = "constant"):
        assert padding_mode in _PADDING_MODES, "Invalid padding mode."
        self.min_angle, self.max_angle, self.expand, self.padding_mode = min_angle, max_angle, expand, padding_mode


class FlipHorizontal(_Bound):
    _fn = staticmethod(_F.flip_horizontal)


class FlipVertical(_Bound):
    _fn = staticmethod(_F.flip_vertical)


class RandomHorizontalFlip:
    def __call__(self, material: MaterialBase, p: float = 0.5) -> MaterialBase:
        return _F.random_horizontal_flip(material, p)


class RandomVerticalFlip:
    def __call__(self, material: MaterialBase, p: float = 0.5) -> MaterialBase:
        return _F.random_vertical_flip(material, p)


class Roll(_Bound):
    _fn, _fields = staticmethod(_F.roll), ("shift",)

    def __init__(self, shift: Tuple[int, int]):
        self.shift = shift


class InvertNormal(_Bound):
    _fn = staticmethod(_F.invert_normal_map)


class AdjustNormalStrength(_Bound):
    _fn, _fields = staticmethod(_F.adjust_normal_strength), ("strength_factor",)

    def __init__(self, strength_factor: float):
        self.strength_factor = strength_factor


class ToLinear(_Bound):
    _fn = staticmethod(_F.to_linear)


class ToSrgb(_Bound):
    _fn = staticmethod(_F.to_srgb)
