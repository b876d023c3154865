"""
pypbr_b200.utils.functions — colour-space conversion and the per-texel normal utilities behind the reference's names.

`srgb_to_linear` / `linear_to_srgb` replace pypbr/utils/functions.py:31-66 with one streaming kernel
(pbr_color_convert); `rotate_normals`, `compute_normal_from_height` replace pypbr/utils/functions.py:69-108 and
:123-177 with pbr_normal_op; `invert_normal` is :111-120; `compute_height_from_normal` (:180-323) runs its per-texel half (gradient field + divergence)
in pbr_normal_op and the Poisson solve as two cuFFT calls.  CUDA float32 tensors only: there is no CPU
implementation in this package.
"""

from __future__ import annotations

import math

import torch

from .. import _cabi
from .enums import NormalConvention


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


def _normal_op(src: torch.Tensor, out: torch.Tensor, op: int, cos_a=0.0, sin_a=0.0, scale=0.0, flip_y=0, aux=None) -> torch.Tensor:
    lib = _cabi.load()
    B = src.shape[0] if src.dim() == 4 else 1
    d = _cabi.PbrNormalOpDesc(B, src.shape[-2], src.shape[-1], op, cos_a, sin_a, scale, flip_y, _cabi.plane(src), _cabi.plane(out),
                              _cabi.plane(aux))
    with torch.cuda.device(src.device):
        _cabi.check(lib.pbr_normal_op(_cabi.byref(d), _cabi.stream_ptr(src.device)), "pbr_normal_op")
    return out


def rotate_normals(normal_map: torch.Tensor, angle: float) -> torch.Tensor:
    """
    Rotate the (x, y) part of every normal by `angle` degrees and renormalise (pypbr/utils/functions.py:69-108).
    Like the reference, the map is modified IN PLACE and returned.  (3,H,W) or (B,3,H,W).
    A map that requires grad is NOT modified: the rotated map is returned as a new tensor with an adjoint kernel behind it
    (pbr_normal_op ROTATE_BWD) - the reference's in-place assignment is tracked by autograd on a non-leaf tensor and refused
    on a leaf; this form is differentiable for both.
    """
    _cabi.require_cuda(normal_map, "normal_map")
    if normal_map.dim() not in (3, 4) or normal_map.shape[-3] != 3:
        raise ValueError(f"normal_map must have shape (3, H, W) or (B, 3, H, W), got {tuple(normal_map.shape)}")
    theta = math.radians(angle)
    if torch.is_grad_enabled() and normal_map.requires_grad:
        return _RotateNormalsFn.apply(normal_map, math.cos(theta), math.sin(theta))
    target = normal_map.detach()
    work = _cabi.rowmajor(target)
    _normal_op(work, work, _cabi.NORMAL_OP_ROTATE, cos_a=math.cos(theta), sin_a=math.sin(theta))
    if work.data_ptr() != target.data_ptr():   # W-strided view: the kernel worked on a packed copy
        target.copy_(work)
    else:
        _cabi.touch(normal_map)
    return normal_map


class _RotateNormalsFn(torch.autograd.Function):
    """normalize(R(angle) (x, y), z) out of place, with its adjoint (pbr_normal_op ROTATE / ROTATE_BWD).  cos = strength,
    sin = 0 is MaterialBase.adjust_normal_strength (base.py:689-706)."""

    @staticmethod
    def forward(ctx, normal_map, cos_a: float, sin_a: float):
        src = _cabi.rowmajor(normal_map.detach())
        ctx.save_for_backward(src)
        ctx.args = (cos_a, sin_a)
        out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
        return _normal_op(src, out, _cabi.NORMAL_OP_ROTATE, cos_a=cos_a, sin_a=sin_a)

    @staticmethod
    def backward(ctx, g):
        (src,) = ctx.saved_tensors
        cos_a, sin_a = ctx.args
        d_in = torch.empty(src.shape, dtype=torch.float32, device=src.device)
        _normal_op(src, d_in, _cabi.NORMAL_OP_ROTATE_BWD, cos_a=cos_a, sin_a=sin_a, aux=_cabi.rowmajor(g.contiguous()))
        return d_in, None, None


def invert_normal(normals: torch.Tensor) -> torch.Tensor:
    """Flip the Y component in place (pypbr/utils/functions.py:111-120)."""
    if normals is not None:
        normals.select(-3, 1).neg_()
    return normals


def compute_normal_from_height(height_map: torch.Tensor, scale: float = 1.0,
                               convention: NormalConvention = NormalConvention.OPENGL) -> torch.Tensor:
    """
    Normal map from a height map (pypbr/utils/functions.py:123-177): one-sided differences of the zero-padded
    height, (-gx*scale, -gy*scale, 1) for OpenGL, (-gx*scale, +gy*scale, 1) for DirectX, normalised.
    (H,W) or (1,H,W) -> (3,H,W); (B,1,H,W) -> (B,3,H,W).
    """
    if height_map is None:
        raise ValueError("Height map is required to compute normals.")
    if convention not in (NormalConvention.OPENGL, NormalConvention.DIRECTX):
        raise ValueError("Unsupported normal convention.")
    _cabi.require_cuda(height_map, "height_map")
    if height_map.dim() == 2:
        height_map = height_map.unsqueeze(0)
    if height_map.dim() not in (3, 4) or height_map.shape[-3] != 1:
        raise ValueError(f"height_map must have shape (H, W), (1, H, W) or (B, 1, H, W), got {tuple(height_map.shape)}")
    flip = 1 if convention == NormalConvention.DIRECTX else 0
    if torch.is_grad_enabled() and height_map.requires_grad:
        return _NormalFromHeightFn.apply(height_map, float(scale), flip)
    return _normal_from_height(_cabi.rowmajor(height_map.detach()), float(scale), flip)


def _normal_from_height(src: torch.Tensor, scale: float, flip: int) -> torch.Tensor:
    shape = list(src.shape)
    shape[-3] = 3
    out = torch.empty(shape, dtype=torch.float32, device=src.device)
    return _normal_op(src, out, _cabi.NORMAL_OP_FROM_HEIGHT, scale=scale, flip_y=flip)


class _NormalFromHeightFn(torch.autograd.Function):
    """compute_normal_from_height with its adjoint (pbr_normal_op FROM_HEIGHT_BWD): the reference's pad / difference /
    normalise sequence (utils/functions.py:144-175) is differentiable, so a height map can be fitted through its normals."""

    @staticmethod
    def forward(ctx, height_map, scale: float, flip: int):
        src = _cabi.rowmajor(height_map.detach())
        ctx.save_for_backward(src)
        ctx.args = (scale, flip)
        return _normal_from_height(src, scale, flip)

    @staticmethod
    def backward(ctx, g):
        (src,) = ctx.saved_tensors
        scale, flip = ctx.args
        d_h = torch.empty(src.shape, dtype=torch.float32, device=src.device)
        _normal_op(src, d_h, _cabi.NORMAL_OP_FROM_HEIGHT_BWD, scale=scale, flip_y=flip, aux=_cabi.rowmajor(g))
        return d_h, None, None


def compute_height_from_normal(normal_map: torch.Tensor, scale: float = 1.0,
                               convention: NormalConvention = NormalConvention.OPENGL) -> torch.Tensor:
    """
    Height map from a normal map by Poisson reconstruction (pypbr/utils/functions.py:180-323), (3,H,W) -> (1,H,W) in
    [0,1].  The gradient field (-Nx, -+Ny)/(Nz + 1e-8)*scale and its forward-difference divergence are one kernel
    (pbr_normal_op DIVERGENCE); the solve divides the 2-D FFT of the divergence by the eigenvalues of the periodic
    5-point Laplacian, 2cos(2 pi x/W) + 2cos(2 pi y/H) - 4, with the zero frequency removed (cuFFT through torch.fft);
    the result is shifted and scaled to [0, 1] like the reference does.
    """
    if normal_map is None:
        raise ValueError("Normal map is required to compute height.")
    if normal_map.shape[0] != 3:
        raise ValueError("Normal map must have three channels.")
    if convention not in (NormalConvention.OPENGL, NormalConvention.DIRECTX):
        raise ValueError("Unsupported normal convention.")
    _cabi.require_cuda(normal_map, "normal_map")
    flip = 1 if convention == NormalConvention.DIRECTX else 0
    if torch.is_grad_enabled() and normal_map.requires_grad:
        # differentiable like the reference's op sequence: the divergence kernel has an adjoint (DIVERGENCE_BWD), the solve and
        # the [0, 1] normalisation below are torch ops autograd tracks
        div = _DivergenceFn.apply(normal_map, float(scale), flip)
        src = div
    else:
        src = _cabi.rowmajor(normal_map.detach())
        div = _divergence(src, float(scale), flip)
    H, W = src.shape[-2:]
    wy = 2 * math.pi * torch.arange(H, dtype=torch.float32, device=src.device).view(-1, 1) / H
    wx = 2 * math.pi * torch.arange(W, dtype=torch.float32, device=src.device).view(1, -1) / W
    eig = (2 * torch.cos(wx) - 2) + (2 * torch.cos(wy) - 2)
    eig[0, 0] = 1.0
    spec = torch.fft.fft2(div[0]) / eig
    if spec.requires_grad:       # the same zero at the mean frequency, without writing into a tensor autograd holds
        keep = torch.ones_like(eig)
        keep[0, 0] = 0.0
        spec = spec * keep
    else:
        spec[0, 0] = 0
    height = torch.fft.ifft2(spec).real
    height = height - height.mean()
    lo, hi = height.min(), height.max()
    return ((height - lo) / (hi - lo + 1e-8)).unsqueeze(0)


def _divergence(src: torch.Tensor, scale: float, flip: int) -> torch.Tensor:
    H, W = src.shape[-2:]
    div = torch.empty((1, H, W), dtype=torch.float32, device=src.device)
    return _normal_op(src, div, _cabi.NORMAL_OP_DIVERGENCE, scale=scale, flip_y=flip)


class _DivergenceFn(torch.autograd.Function):
    """The gradient field of a normal map and its divergence (utils/functions.py:241-283) with the adjoint kernel."""

    @staticmethod
    def forward(ctx, normal_map, scale: float, flip: int):
        src = _cabi.rowmajor(normal_map.detach())
        ctx.save_for_backward(src)
        ctx.args = (scale, flip)
        return _divergence(src, scale, flip)

    @staticmethod
    def backward(ctx, g):
        (src,) = ctx.saved_tensors
        scale, flip = ctx.args
        d_n = torch.empty(src.shape, dtype=torch.float32, device=src.device)
        _normal_op(src, d_n, _cabi.NORMAL_OP_DIVERGENCE_BWD, scale=scale, flip_y=flip, aux=_cabi.rowmajor(g.contiguous()))
        return d_n, None, None
