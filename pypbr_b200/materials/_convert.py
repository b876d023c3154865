"""Launch helpers for the workflow-conversion kernels (pbr_convert_m2s / pbr_convert_s2m)."""

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


def convert(albedo: torch.Tensor, second: torch.Tensor, albedo_is_srgb: bool, m2s: bool):
    """
    m2s: (albedo, metallic 1ch) -> (diffuse, specular).   pypbr/materials/metallic.py:90-109
    s2m: (diffuse, raw specular 3ch) -> (basecolor, metallic 3ch).   pypbr/materials/diffuse.py:112-147
    """
    _cabi.require_cuda(albedo, "albedo")
    _cabi.require_cuda(second, "metallic" if m2s else "specular")
    lib = _cabi.load()
    if second.dim() == albedo.dim() - 1:
        second = second.unsqueeze(-3)  # (H, W) -> (1, H, W), as metallic.py:99-100 / diffuse.py:124-125
    second = _match_size(second, albedo, None if m2s else True)
    want = 1 if m2s else 3
    if albedo.shape[-3] != 3:
        raise ValueError(f"albedo must have 3 channels, got {albedo.shape[-3]}")
    if second.shape[-3] not in (want, 1):
        raise ValueError(f"{'metallic' if m2s else 'specular'} must have {want} channel(s), got {second.shape[-3]}")
    if second.shape[-3] != want:
        second = second.expand(*second.shape[:-3], want, *second.shape[-2:])
    a = _cabi.rowmajor(albedo.detach())
    s = _cabi.rowmajor(second.detach())
    B = a.shape[0] if a.dim() == 4 else 1
    if s.dim() != a.dim():
        raise ValueError("albedo and the second map must both be batched or both unbatched")
    out0 = torch.empty(a.shape, dtype=torch.float32, device=a.device)
    out1 = torch.empty(a.shape, dtype=torch.float32, device=a.device)
    d = _cabi.PbrConvDesc(B, a.shape[-2], a.shape[-1], int(bool(albedo_is_srgb)), _cabi.plane(a), _cabi.plane(s),
                          _cabi.plane(out0), _cabi.plane(out1))
    fn = lib.pbr_convert_m2s if m2s else lib.pbr_convert_s2m
    with torch.cuda.device(a.device):
        _cabi.check(fn(_cabi.byref(d), _cabi.stream_ptr(a.device)), "pbr_convert")
    return out0, out1
