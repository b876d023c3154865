"""
pypbr_b200.utils.functions — colour-space conversion behind the reference's names.

`srgb_to_linear` / `linear_to_srgb` replace pypbr/utils/functions.py:31-66 with one streaming kernel
(pbr_color_convert).  CUDA float32 tensors only: there is no CPU implementation in this package.
"""

from __future__ import annotations

import torch

from .. import _cabi


class _ColorFn(torch.autograd.Function):
    """Autograd wrapper so `material.to_linear()` style code stays differentiable."""

    @staticmethod
    def forward(ctx, tex: torch.Tensor, to_linear: bool):
        out = _color_convert(tex, to_linear)
        ctx.save_for_backward(tex)
        ctx.to_linear = to_linear
        return out

    @staticmethod
    def backward(ctx, grad):
        (tex,) = ctx.saved_tensors
        # d/dx through clamp -> piecewise curve -> clamp (SURVEY.md §8a autograd conventions); rarely
        # needed (the shading kernels fuse their own colour adjoints), so it is written with torch ops.
        t = tex.clamp(0, 1)
        inside = ((tex >= 0) & (tex <= 1)).to(tex.dtype)
        if ctx.to_linear:
            d = torch.where(t <= 0.04045, torch.full_like(t, 1 / 12.92), (2.4 / 1.055) * ((t + 0.055) / 1.055) ** 1.4)
        else:
            d = torch.where(t <= 0.0031308, torch.full_like(t, 12.92), (1.055 / 2.4) * t.clamp_min(1e-12) ** (1 / 2.4 - 1))
        return grad * d * inside, None


def _color_convert(tex: torch.Tensor, to_linear: bool) -> torch.Tensor:
    _cabi.require_cuda(tex, "texture")
    lib = _cabi.load()
    shape = tex.shape
    if tex.dim() < 2:
        raise ValueError("texture must have at least 2 dimensions (..., H, W)")
    H, W = shape[-2], shape[-1]
    src = tex.reshape(1, -1, H, W) if tex.dim() != 4 else tex
    src = _cabi.rowmajor(src)
    out = torch.empty(src.shape, dtype=torch.float32, device=tex.device)
    d = _cabi.PbrColorDesc(src.shape[0], src.shape[1], H, W, 1 if to_linear else 0, _cabi.plane(src), _cabi.plane(out))
    with torch.cuda.device(tex.device):
        _cabi.check(lib.pbr_color_convert(_cabi.byref(d), _cabi.stream_ptr(tex.device)), "pbr_color_convert")
    return out.reshape(shape)


def srgb_to_linear(texture: torch.Tensor) -> torch.Tensor:
    """Convert an sRGB texture (..., C, H, W) to linear space (pypbr/utils/functions.py:31-47)."""
    if texture.requires_grad:
        return _ColorFn.apply(texture, True)
    return _color_convert(texture, True)


def linear_to_srgb(texture: torch.Tensor) -> torch.Tensor:
    """Convert a linear texture (..., C, H, W) to sRGB space (pypbr/utils/functions.py:50-66)."""
    if texture.requires_grad:
        return _ColorFn.apply(texture, False)
    return _color_convert(texture, False)
