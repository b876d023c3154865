"""
pypbr_b200.transforms.functional — the functional transforms of pypbr/transforms/functional.py:52-380 on top of the
material methods whose CUDA halves are the index-transform / normal / colour kernels.

Same names, signatures, return convention (a NEW material; the argument is never modified) and the same draws from
`random.random()` in the same order as the reference, so a seeded pipeline picks the same crops / sizes / angles.

One deliberate difference in HOW the new material comes about.  The reference deep-copies the material
(`material.clone()`: every map read and written once) and then lets the method replace every map again.  For the
transforms that rebuild every map anyway (resize, crop, tile, flip, roll) this module starts from a copy of the
material that SHARES the source tensors, runs the method - on CUDA maps: one pbr_index_transform launch for all maps -
and afterwards clones only what still aliases the source (crop returns views; torchvision's resize returns its input
when the size already matches).  Same result, no aliasing ever leaks to the caller, half the HBM traffic of an
augmentation step.  Transforms that touch a single map or work in place (rotate's normal rotation, invert_normal,
adjust_normal_strength, the colour-space conversions) keep the reference's deep copy.
"""

from __future__ import annotations

import copy
from random import random
from typing import Tuple

from ..materials import MaterialBase

__all__ = ["resize", "random_resize", "crop", "center_crop", "random_crop", "tile", "rotate", "random_rotate",
           "flip_horizontal", "flip_vertical", "random_horizontal_flip", "random_vertical_flip", "roll",
           "invert_normal_map", "adjust_normal_strength", "to_linear", "to_srgb"]


def _sharing_copy(material: MaterialBase) -> MaterialBase:
    """A new material object of the same class whose map registry is its own dict but whose tensors are the source's."""
    new = material.__class__.__new__(material.__class__)
    for key, val in material.__dict__.items():
        object.__setattr__(new, key, dict(val) if key == "_maps" else copy.copy(val))
    return new


def _rebuilt(material: MaterialBase, method: str, *args, **kwargs) -> MaterialBase:
    """`method` replaces every map of the material: run it on a tensor-sharing copy, then un-alias what is left."""
    new = _sharing_copy(material)
    getattr(new, method)(*args, **kwargs)
    for name, t in new._maps.items():
        src = material._maps.get(name, None)
        if t is not None and src is not None and t.untyped_storage().data_ptr() == src.untyped_storage().data_ptr():
            new._maps[name] = t.clone()
    return new


def _cloned(material: MaterialBase, method: str, *args, **kwargs) -> MaterialBase:
    """The reference's own scheme: deep copy, then the (possibly in-place) method on the copy."""
    new = material.clone()
    getattr(new, method)(*args, **kwargs)
    return new


# ------------------------------------------------------------------------------------------ resizing (functional.py:52-93)
def resize(material: MaterialBase, size: Tuple[int, int], antialias: bool = True) -> MaterialBase:
    """All maps resized to `size` = (height, width) (TF.resize, the reference's own library call)."""
    return _rebuilt(material, "resize", size=size, antialias=antialias)


def random_resize(material: MaterialBase, min_size: int, max_size: int, antialias: bool = True) -> MaterialBase:
    """Height, then width, drawn uniformly from [min_size, max_size) (functional.py:88-90)."""
    span = max_size - min_size
    height = int(min_size + span * random())
    width = int(min_size + span * random())
    return _rebuilt(material, "resize", size=(height, width), antialias=antialias)


# ------------------------------------------------------------------------------------------ cropping (functional.py:96-157)
def crop(material: MaterialBase, top: int, left: int, height: int, width: int) -> MaterialBase:
    """The region [top, top + height) x [left, left + width) of every map, as tensors of their own."""
    return _rebuilt(material, "crop", top=top, left=left, height=height, width=width)


def center_crop(material: MaterialBase, crop_size: Tuple[int, int]) -> MaterialBase:
    full_h, full_w = material.size
    h, w = crop_size
    return crop(material, (full_h - h) // 2, (full_w - w) // 2, h, w)


def random_crop(material: MaterialBase, crop_size: Tuple[int, int]) -> MaterialBase:
    """Top, then left, drawn uniformly (functional.py:153-154)."""
    full_h, full_w = material.size
    h, w = crop_size
    top = int((full_h - h) * random())
    left = int((full_w - w) * random())
    return crop(material, top, left, h, w)


# ------------------------------------------------------------------------------------------ tiling (functional.py:160-175)
def tile(material: MaterialBase, num_tiles: int) -> MaterialBase:
    return _rebuilt(material, "tile", num_tiles=num_tiles)


# ------------------------------------------------------------------------------------------ rotation (functional.py:179-227)
def rotate(material: MaterialBase, angle: float, expand: bool = False, padding_mode: str = "constant") -> MaterialBase:
    return _cloned(material, "rotate", angle=angle, expand=expand, padding_mode=padding_mode)


def random_rotate(material: MaterialBase, min_angle: float = 0.0, max_angle: float = 360.0, expand: bool = False,
                  padding_mode: str = "constant") -> MaterialBase:
    return rotate(material, min_angle + (max_angle - min_angle) * random(), expand, padding_mode)


# ------------------------------------------------------------------------------------------ flipping (functional.py:231-294)
def flip_horizontal(material: MaterialBase) -> MaterialBase:
    """Mirrored along W, the normal's x negated (one gather launch for all maps on CUDA)."""
    return _rebuilt(material, "flip_horizontal")


def flip_vertical(material: MaterialBase) -> MaterialBase:
    """Mirrored along H, the normal's y negated."""
    return _rebuilt(material, "flip_vertical")


def random_horizontal_flip(material: MaterialBase, p: float = 0.5) -> MaterialBase:
    return flip_horizontal(material) if random() < p else material.clone()


def random_vertical_flip(material: MaterialBase, p: float = 0.5) -> MaterialBase:
    return flip_vertical(material) if random() < p else material.clone()


# ------------------------------------------------------------------------------------------ translation (functional.py:298-313)
def roll(material: MaterialBase, shift: Tuple[int, int]) -> MaterialBase:
    """torch.roll by `shift` = (rows, columns) on every map."""
    return _rebuilt(material, "roll", shift=shift)


# ------------------------------------------------------------------------------------------ normal map (functional.py:317-349)
def invert_normal_map(material: MaterialBase) -> MaterialBase:
    return _cloned(material, "invert_normal")


def adjust_normal_strength(material: MaterialBase, strength_factor: float) -> MaterialBase:
    return _cloned(material, "adjust_normal_strength", strength_factor=strength_factor)


# ------------------------------------------------------------------------------------------ colour space (functional.py:353-380)
def to_linear(material: MaterialBase) -> MaterialBase:
    return _cloned(material, "to_linear")


def to_srgb(material: MaterialBase) -> MaterialBase:
    return _cloned(material, "to_srgb")
