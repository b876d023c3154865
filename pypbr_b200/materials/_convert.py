"""Launch helpers for the workflow-conversion kernels (pbr_convert_m2s / pbr_convert_s2m) and their adjoints."""

from __future__ import annotations

import torch

from .. import _cabi


def _match_size(src: torch.Tensor, like: torch.Tensor, antialias) -> torch.Tensor:
    """metallic.py:93-96 / diffuse.py:118-121: the secondary map is resized to the albedo's size."""
    if src.shape[-2:] == like.shape[-2:]:
        return src
    from torchvision.transforms import functional as TF

    if antialias is None:
        return TF.resize(src, like.shape[-2:])
    return TF.resize(src, like.shape[-2:], antialias=antialias)


def _desc(a: torch.Tensor, s: torch.Tensor, albedo_is_srgb: bool, m2s: bool, out0=None, out1=None) -> "_cabi.PbrConvDesc":
    B = a.shape[0] if a.dim() == 4 else 1
    d = _cabi.PbrConvDesc(B, a.shape[-2], a.shape[-1], int(bool(albedo_is_srgb)), _cabi.plane(a), _cabi.plane(s),
                          _cabi.plane(out0), _cabi.plane(out1))
    d.metallic_channels = s.shape[-3] if m2s else 0
    return d


def _launch(a: torch.Tensor, s: torch.Tensor, albedo_is_srgb: bool, m2s: bool):
    lib = _cabi.load()
    out0 = torch.empty(a.shape, dtype=torch.float32, device=a.device)
    out1 = torch.empty(a.shape, dtype=torch.float32, device=a.device)
    d = _desc(a, s, albedo_is_srgb, m2s, out0, out1)
    fn = lib.pbr_convert_m2s if m2s else lib.pbr_convert_s2m
    with torch.cuda.device(a.device):
        _cabi.check(fn(_cabi.byref(d), _cabi.stream_ptr(a.device)), "pbr_convert")
    return out0, out1


class _ConvertFn(torch.autograd.Function):
    """
    The reference's conversions are plain differentiable torch ops (metallic.py:103-109, diffuse.py:129-147); this keeps
    a fit that goes through one of them differentiable: the backward is ONE streaming kernel that recomputes the
    forward per texel (pbr_convert_m2s_backward / pbr_convert_s2m_backward).
    """

    @staticmethod
    def forward(ctx, albedo, second, albedo_is_srgb: bool, m2s: bool):
        a, s = _cabi.rowmajor(albedo.detach()), _cabi.rowmajor(second.detach())
        out0, out1 = _launch(a, s, albedo_is_srgb, m2s)
        ctx.save_for_backward(a, s)
        ctx.flags = (albedo_is_srgb, m2s)
        return out0, out1

    @staticmethod
    def backward(ctx, g0, g1):
        a, s = ctx.saved_tensors
        albedo_is_srgb, m2s = ctx.flags
        lib = _cabi.load()
        g0 = _cabi.rowmajor(g0) if g0 is not None else None
        g1 = _cabi.rowmajor(g1) if g1 is not None else None
        need_a, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_a = torch.empty(a.shape, dtype=torch.float32, device=a.device) if need_a else None
        d_s = torch.empty(s.shape, dtype=torch.float32, device=a.device) if need_s else None
        d = _desc(a, s, albedo_is_srgb, m2s)
        g = _cabi.PbrConvGrads(_cabi.plane(g0), _cabi.plane(g1), _cabi.plane(d_a), _cabi.plane(d_s))
        fn = lib.pbr_convert_m2s_backward if m2s else lib.pbr_convert_s2m_backward
        with torch.cuda.device(a.device):
            _cabi.check(fn(_cabi.byref(d), _cabi.byref(g), _cabi.stream_ptr(a.device)), "pbr_convert_backward")
        return d_a, d_s, None, None


def convert(albedo: torch.Tensor, second: torch.Tensor, albedo_is_srgb: bool, m2s: bool):
    """
    m2s: (albedo, metallic 1ch | 3ch) -> (diffuse, specular).   pypbr/materials/metallic.py:90-109
    s2m: (diffuse, raw specular 3ch) -> (basecolor, metallic 3ch).   pypbr/materials/diffuse.py:112-147
    Differentiable w.r.t. both inputs (torch.autograd.Function around the kernels).
    """
    _cabi.require_cuda(albedo, "albedo")
    _cabi.require_cuda(second, "metallic" if m2s else "specular")
    if second.dim() == albedo.dim() - 1:
        second = second.unsqueeze(-3)  # (H, W) -> (1, H, W), as metallic.py:99-100 / diffuse.py:124-125
    second = _match_size(second, albedo, None if m2s else True)
    if albedo.shape[-3] != 3:
        raise ValueError(f"albedo must have 3 channels, got {albedo.shape[-3]}")
    allowed = (1, 3)   # m2s: the per-channel metallic that s2m produces broadcasts channel by channel (metallic.py:103-106)
    if second.shape[-3] not in allowed:
        raise ValueError(f"{'metallic' if m2s else 'specular'} must have 1 or 3 channels, got {second.shape[-3]}")
    if not m2s and second.shape[-3] != 3:
        second = second.expand(*second.shape[:-3], 3, *second.shape[-2:])
    if second.dim() != albedo.dim():
        raise ValueError("albedo and the second map must both be batched or both unbatched")
    if torch.is_grad_enabled() and (albedo.requires_grad or second.requires_grad):
        return _ConvertFn.apply(albedo, second, bool(albedo_is_srgb), bool(m2s))
    return _launch(_cabi.rowmajor(albedo.detach()), _cabi.rowmajor(second.detach()), albedo_is_srgb, m2s)
