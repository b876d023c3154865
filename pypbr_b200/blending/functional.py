"""
pypbr_b200.blending.functional — blend_materials and friends behind the reference's names.

Mirrors pypbr/blending/functional.py (same signatures, defaults, exceptions and the
``(blended_material, mask)`` tuple return).  Mask construction (sigmoid of a height / property
difference, or a linspace gradient), the per-map lerp and the normalise-lerp-normalise of normal maps
all run in ONE streaming kernel launch (pbr_blend) over every map the two materials share; the
kernel also returns min(blended normal), the probe MaterialBase._process_normal_map needs.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import _cabi
from ..materials import MaterialBase
from ..materials.base import _normal_ingest_cuda


def blend_materials(material1: MaterialBase, material2: MaterialBase, method: str = "mask", **kwargs) -> MaterialBase:
    """
    Blend two materials.  method: 'mask' (kwarg mask), 'height' (blend_width=0.1),
    'properties' (property_name='metallic', blend_width=0.1), 'gradient' (direction='horizontal').
    Dispatch as pypbr/blending/functional.py:27-61 (`shift` is not forwarded there either).
    """
    if method == "mask":
        mask = kwargs.get("mask", None)
        if mask is None:
            raise ValueError("Mask must be provided for 'mask' blending method.")
        return blend_with_mask(material1, material2, mask)
    if method == "height":
        return blend_on_height(material1, material2, kwargs.get("blend_width", 0.1))
    if method == "properties":
        return blend_on_properties(material1, material2, kwargs.get("property_name", "metallic"), kwargs.get("blend_width", 0.1))
    if method == "gradient":
        return blend_with_gradient(material1, material2, kwargs.get("direction", "horizontal"))
    raise ValueError(f"Unknown blending method: {method}")


# ------------------------------------------------------------------------------------------------
def _common_device(material1: MaterialBase, material2: MaterialBase) -> torch.device:
    for m in (material1, material2):
        for t in m._maps.values():
            if t is not None:
                if not t.is_cuda:
                    raise RuntimeError(
                        f"pypbr_b200: blending runs on CUDA only (found a map on {t.device}); there is no CPU fallback. "
                        "Move both materials with material.to('cuda')."
                    )
                return t.device
    raise ValueError("Materials must have at least one map to blend.")


def _run_blend(material1: MaterialBase, material2: MaterialBase, mode: int, mask=None, prop1=None, prop2=None,
               blend_width: float = 0.0, shift: float = 0.0, apply_shift: bool = False, size=None):
    """Shared tail of every blend_* function: pypbr/blending/functional.py:76-116."""
    lib = _cabi.load()
    device = _common_device(material1, material2)
    blended = material1.__class__()

    names = sorted(set(material1._maps.keys()).union(material2._maps.keys()))
    pairs, passthrough = [], {}
    for name in names:
        a, b = material1._maps.get(name, None), material2._maps.get(name, None)
        if a is None and b is None:
            passthrough[name] = None
        elif a is None:
            passthrough[name] = b
        elif b is None:
            passthrough[name] = a
        else:
            pairs.append((name, a, b))

    if size is None:
        ref_t = pairs[0][1] if pairs else (mask if mask is not None else prop1)
        size = tuple(ref_t.shape[-2:])
    H, W = size
    batched = any(t.dim() == 4 for _, a, b in pairs for t in (a, b)) or (mask is not None and mask.dim() == 4)
    B = 1
    if batched:
        B = max([t.shape[0] for _, a, b in pairs for t in (a, b) if t.dim() == 4] + ([mask.shape[0]] if (mask is not None and mask.dim() == 4) else []))

    def conform(t: torch.Tensor, channels: Optional[int], what: str) -> torch.Tensor:
        _cabi.require_cuda(t, what)
        t = t.detach()
        if batched and t.dim() == 3:
            t = t.unsqueeze(0)
        if tuple(t.shape[-2:]) != (H, W):
            raise ValueError(f"{what} has spatial size {tuple(t.shape[-2:])}, expected {(H, W)}")
        if channels is not None and t.shape[-3] != channels:
            if t.shape[-3] != 1:
                raise ValueError(f"{what} has {t.shape[-3]} channels, cannot broadcast to {channels}")
            t = t.expand(*t.shape[:-3], channels, H, W)
        if batched and t.shape[0] != B:
            if t.shape[0] != 1:
                raise ValueError(f"{what} has batch {t.shape[0]}, expected {B}")
            t = t.expand(B, *t.shape[1:])
        return _cabi.rowmajor(t)

    keep = []
    mask_out = None
    d = _cabi.PbrBlendDesc()
    d.B, d.H, d.W = B, H, W
    d.mask_mode = mode
    d.blend_width, d.shift, d.apply_shift = float(blend_width), float(shift), int(apply_shift)
    if mode == _cabi.MASK_GIVEN:
        mk = conform(mask.to(device), 1, "mask")
        keep.append(mk)
        d.mask = _cabi.plane(mk)
    else:
        if mode == _cabi.MASK_SIGMOID:
            p1, p2 = conform(prop1, 1, "property map 1"), conform(prop2, 1, "property map 2")
            keep.extend((p1, p2))
            d.prop1, d.prop2 = _cabi.plane(p1), _cabi.plane(p2)
        mask_out = torch.empty((B, 1, H, W) if batched else (1, H, W), dtype=torch.float32, device=device)
        d.mask_out = _cabi.plane(mask_out)

    normal_min = torch.full((1,), float("inf"), dtype=torch.float32, device=device)
    outs = {}
    jobs = []
    for name, a, b in pairs:
        ch = max(a.shape[-3], b.shape[-3])
        ta, tb = conform(a, ch, f"{name} (material1)"), conform(b, ch, f"{name} (material2)")
        out = torch.empty(ta.shape, dtype=torch.float32, device=device)
        is_normal = name == "normal"
        if is_normal and ch != 3:
            raise ValueError("Normal maps must have 3 channels to be blended.")
        if ch > 4:
            raise ValueError(f"map '{name}' has {ch} channels; at most 4 are supported per map")
        jobs.append((ta, tb, out, ch, is_normal))
        outs[name] = out
        keep.extend((ta, tb))

    with torch.cuda.device(device):
        first = True
        chunk = _cabi.PBR_MAX_BLEND_MAPS
        for start in range(0, max(len(jobs), 1), chunk):
            part = jobs[start : start + chunk]
            d.n_maps = len(part)
            for i, (ta, tb, out, ch, is_normal) in enumerate(part):
                d.maps[i] = _cabi.PbrBlendMap(_cabi.plane(ta), _cabi.plane(tb), _cabi.plane(out), ch, int(is_normal))
            d.normal_min = normal_min.data_ptr() if any(j[4] for j in part) else None
            if not first:
                d.mask_out = _cabi.PbrPlane(None, 0, 0, 0)  # already written by the first launch
            _cabi.check(lib.pbr_blend(_cabi.byref(d), _cabi.stream_ptr(device)), "pbr_blend")
            first = False

    # hand the maps to the new material the way `setattr(blended_material, name, map)` would
    blended.device = device
    for name in names:
        if name in outs:
            t = outs[name]
            if name == "normal":
                # base.py:210-217 on the blended normal: min() < 0 -> keep, else remap *2-1 and renormalise
                if not (float(normal_min.item()) < 0):
                    t = _normal_ingest_cuda(t, 3)
            blended._maps[name] = t
        else:
            setattr(blended, name, passthrough[name])
    blended.albedo_is_srgb = material1.albedo_is_srgb
    blended.device = material1.device
    return blended, (mask_out if mask_out is not None else mask)


def blend_with_mask(material1: MaterialBase, material2: MaterialBase, mask: torch.Tensor) -> MaterialBase:
    """
    Blend with a given mask of shape [1, H, W] or [H, W] (pypbr/blending/functional.py:64-116).
    Returns (blended_material, mask) like the reference.
    """
    if mask.dim() == 2:
        mask = mask.unsqueeze(0)
    elif mask.dim() == 4 and mask.size(1) == 1:
        pass  # batched extension: (B, 1, H, W)
    elif mask.dim() != 3 or mask.size(0) != 1:
        raise ValueError("Mask must have shape [1, H, W] or [H, W].")
    return _run_blend(material1, material2, _cabi.MASK_GIVEN, mask=mask)


def _sigmoid_blend(material1, material2, p1, p2, blend_width, shift, apply_shift):
    if p1.shape != p2.shape:
        from torchvision.transforms import functional as TF

        p2 = TF.resize(p2, p1.shape[-2:], antialias=True)
    return _run_blend(material1, material2, _cabi.MASK_SIGMOID, prop1=p1, prop2=p2, blend_width=blend_width,
                      shift=shift, apply_shift=apply_shift, size=tuple(p1.shape[-2:]))


def blend_on_height(material1: MaterialBase, material2: MaterialBase, blend_width: float = 0.1, shift: float = 0.0) -> MaterialBase:
    """mask = sigmoid((height1 + shift - height2) / (blend_width + 1e-6)) (functional.py:148-196)."""
    h1, h2 = material1._maps.get("height", None), material2._maps.get("height", None)
    if h1 is None or h2 is None:
        raise ValueError("Both materials must have height maps for height-based blending.")
    return _sigmoid_blend(material1, material2, h1, h2, blend_width, shift, True)


def blend_on_properties(material1: MaterialBase, material2: MaterialBase, property_name: str = "metallic",
                        blend_width: float = 0.1) -> MaterialBase:
    """mask = sigmoid((prop1 - prop2) / (blend_width + 1e-6)) (functional.py:199-239)."""
    p1, p2 = material1._maps.get(property_name, None), material2._maps.get(property_name, None)
    if p1 is None or p2 is None:
        raise ValueError(f"Both materials must have '{property_name}' maps for property-based blending.")
    return _sigmoid_blend(material1, material2, p1, p2, blend_width, 0.0, False)


def blend_with_gradient(material1: MaterialBase, material2: MaterialBase, direction: str = "horizontal") -> MaterialBase:
    """mask = linspace(0, 1) along W ('horizontal') or H ('vertical') (functional.py:242-286)."""
    size: Optional[Tuple[int, int]] = material1.size
    if size is None:
        raise ValueError("Materials must have at least one map to determine size.")
    if direction == "horizontal":
        mode = _cabi.MASK_GRADIENT_H
    elif direction == "vertical":
        mode = _cabi.MASK_GRADIENT_V
    else:
        raise ValueError("Direction must be 'horizontal' or 'vertical'.")
    return _run_blend(material1, material2, mode, size=size)
