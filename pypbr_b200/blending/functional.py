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
from ..materials.base import _normal_ingest_cuda, normal_ingest


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


class _BlendMeta:
    """Everything about one blend that is not a tensor."""

    __slots__ = ("B", "H", "W", "batched", "mode", "blend_width", "shift", "apply_shift", "device", "jobs")


def _fill_blend_desc(meta: _BlendMeta) -> "_cabi.PbrBlendDesc":
    d = _cabi.PbrBlendDesc()
    d.B, d.H, d.W = meta.B, meta.H, meta.W
    d.mask_mode = meta.mode
    d.blend_width, d.shift, d.apply_shift = float(meta.blend_width), float(meta.shift), int(meta.apply_shift)
    return d


def _blend_launch(meta: _BlendMeta, lead, maps):
    """pbr_blend over every (a, b) pair of `maps` (chunks of PBR_MAX_BLEND_MAPS).  lead: (mask,) | (prop1, prop2) | ().
    Returns (mask written by the kernel or None, outputs, device scalar min(blended normal))."""
    lib = _cabi.load()
    device = meta.device
    d = _fill_blend_desc(meta)
    mask_out = None
    if meta.mode == _cabi.MASK_GIVEN:
        d.mask = _cabi.plane(lead[0])
    else:
        if meta.mode == _cabi.MASK_SIGMOID:
            d.prop1, d.prop2 = _cabi.plane(lead[0]), _cabi.plane(lead[1])
        mask_out = torch.empty((meta.B, 1, meta.H, meta.W) if meta.batched else (1, meta.H, meta.W), dtype=torch.float32, device=device)
        d.mask_out = _cabi.plane(mask_out)
    normal_min = torch.full((1,), float("inf"), dtype=torch.float32, device=device)
    outs = [torch.empty(maps[2 * i].shape, dtype=torch.float32, device=device) for i in range(len(meta.jobs))]
    with torch.cuda.device(device):
        first = True
        chunk = _cabi.PBR_MAX_BLEND_MAPS
        for start in range(0, max(len(meta.jobs), 1), chunk):
            part = range(start, min(start + chunk, len(meta.jobs)))
            d.n_maps = len(part)
            for i, j in enumerate(part):
                _name, ch, is_normal = meta.jobs[j]
                d.maps[i] = _cabi.PbrBlendMap(_cabi.plane(maps[2 * j]), _cabi.plane(maps[2 * j + 1]), _cabi.plane(outs[j]), ch, int(is_normal))
            d.normal_min = normal_min.data_ptr() if any(meta.jobs[j][2] for j in part) else None
            if not first:
                d.mask_out = _cabi.PbrPlane(None, 0, 0, 0)  # already written by the first launch
            _cabi.check(lib.pbr_blend(_cabi.byref(d), _cabi.stream_ptr(device)), "pbr_blend")
            first = False
    return mask_out, outs, normal_min


class _BlendFn(torch.autograd.Function):
    """
    pbr_blend with its adjoint (pbr_blend_backward).  The reference's blend is a chain of differentiable torch ops
    (functional.py:104-108, :134-143, :187-194), so a fit that goes through a blend needs d_a = mask g, d_b = (1-mask) g,
    the three normalisations of a normal map, and the mask's own gradient (to the given mask, or through the sigmoid to
    the two height / property maps).
    Inputs: meta, then lead tensors ((mask,) | (prop1, prop2) | ()), then a0, b0, a1, b1, ...
    Outputs: (the mask the kernel wrote | a dummy for GIVEN, min(blended normal), out0, out1, ...).
    """

    @staticmethod
    def forward(ctx, meta: _BlendMeta, n_lead: int, *tensors):
        lead = [t.detach() for t in tensors[:n_lead]]
        maps = [t.detach() for t in tensors[n_lead:]]
        mask_out, outs, normal_min = _blend_launch(meta, lead, maps)
        ctx.meta, ctx.n_lead = meta, n_lead
        used = lead[0] if meta.mode == _cabi.MASK_GIVEN else mask_out
        ctx.save_for_backward(used, *lead, *maps)
        if mask_out is None:
            mask_out = torch.empty(0, device=meta.device)
        ctx.mark_non_differentiable(normal_min)
        if meta.mode == _cabi.MASK_GIVEN:
            ctx.mark_non_differentiable(mask_out)
        return (mask_out, normal_min, *outs)

    @staticmethod
    def backward(ctx, g_mask_out, _g_min, *g_outs):
        meta, n_lead = ctx.meta, ctx.n_lead
        used, *rest = ctx.saved_tensors
        lead, maps = rest[:n_lead], rest[n_lead:]
        lib = _cabi.load()
        device = meta.device
        need = ctx.needs_input_grad[2:]
        need_lead, need_maps = need[:n_lead], need[n_lead:]
        d_lead = [torch.empty(t.shape, dtype=torch.float32, device=device) if n else None for t, n in zip(lead, need_lead)]
        d_maps = []
        for j in range(len(meta.jobs)):
            g = g_outs[j]
            for side in (0, 1):
                if not need_maps[2 * j + side]:
                    d_maps.append(None)
                elif g is None:   # zero gradient into this map's output: the kernel skips it
                    d_maps.append(torch.zeros(maps[2 * j + side].shape, dtype=torch.float32, device=device))
                else:
                    d_maps.append(torch.empty(maps[2 * j + side].shape, dtype=torch.float32, device=device))
        given = meta.mode == _cabi.MASK_GIVEN
        g_mask = None if (given or g_mask_out is None) else _cabi.rowmajor(g_mask_out)
        chunk = _cabi.PBR_MAX_BLEND_MAPS
        starts = list(range(0, max(len(meta.jobs), 1), chunk))
        carry = g_mask
        keep = []
        with torch.cuda.device(device):
            for ci, start in enumerate(starts):
                last = ci == len(starts) - 1
                part = range(start, min(start + chunk, len(meta.jobs)))
                d = _fill_blend_desc(meta)
                gr = _cabi.PbrBlendGrads()
                gr.mask = _cabi.plane(used)
                gr.g_mask_out = _cabi.plane(carry)
                if not last:
                    # more maps than one launch takes: intermediate launches only accumulate the raw mask gradient
                    d.mask_mode = _cabi.MASK_GIVEN
                    carry = torch.empty(used.shape, dtype=torch.float32, device=device)
                    keep.append(carry)
                    gr.d_mask = _cabi.plane(carry)
                elif given:
                    gr.d_mask = _cabi.plane(d_lead[0]) if d_lead else _cabi.plane(None)
                elif meta.mode == _cabi.MASK_SIGMOID:
                    gr.d_prop1, gr.d_prop2 = _cabi.plane(d_lead[0]), _cabi.plane(d_lead[1])
                d.n_maps = len(part)
                for i, j in enumerate(part):
                    _name, ch, is_normal = meta.jobs[j]
                    d.maps[i] = _cabi.PbrBlendMap(_cabi.plane(maps[2 * j]), _cabi.plane(maps[2 * j + 1]), _cabi.plane(None), ch, int(is_normal))
                    g = g_outs[j]
                    if g is not None:
                        g = _cabi.rowmajor(g)
                        keep.append(g)
                    gr.maps[i] = _cabi.PbrBlendGradMap(_cabi.plane(g), _cabi.plane(d_maps[2 * j] if g is not None else None),
                                                       _cabi.plane(d_maps[2 * j + 1] if g is not None else None))
                _cabi.check(lib.pbr_blend_backward(_cabi.byref(d), _cabi.byref(gr), _cabi.stream_ptr(device)), "pbr_blend_backward")
        return (None, None, *d_lead, *d_maps)


def _run_blend(material1: MaterialBase, material2: MaterialBase, mode: int, mask=None, prop1=None, prop2=None,
               blend_width: float = 0.0, shift: float = 0.0, apply_shift: bool = False, size=None):
    """Shared tail of every blend_* function: pypbr/blending/functional.py:76-116."""
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
        # view ops only (unsqueeze / expand / contiguous): the autograd graph of a map that requires grad stays attached
        _cabi.require_cuda(t, what)
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

    meta = _BlendMeta()
    meta.B, meta.H, meta.W, meta.batched, meta.mode, meta.device = B, H, W, batched, mode, device
    meta.blend_width, meta.shift, meta.apply_shift = blend_width, shift, apply_shift
    lead = []
    if mode == _cabi.MASK_GIVEN:
        lead = [conform(mask.to(device), 1, "mask")]
    elif mode == _cabi.MASK_SIGMOID:
        lead = [conform(prop1, 1, "property map 1"), conform(prop2, 1, "property map 2")]
    maps, jobs = [], []
    for name, a, b in pairs:
        ch = max(a.shape[-3], b.shape[-3])
        is_normal = name == "normal"
        if is_normal and ch != 3:
            raise ValueError("Normal maps must have 3 channels to be blended.")
        if ch > 4:
            raise ValueError(f"map '{name}' has {ch} channels; at most 4 are supported per map")
        maps.extend((conform(a, ch, f"{name} (material1)"), conform(b, ch, f"{name} (material2)")))
        jobs.append((name, ch, is_normal))
    meta.jobs = jobs

    with_grad = torch.is_grad_enabled() and any(t.requires_grad for t in lead + maps)
    if with_grad:
        mask_out, normal_min, *outs = _BlendFn.apply(meta, len(lead), *lead, *maps)
        if mode == _cabi.MASK_GIVEN:
            mask_out = None
    else:
        mask_out, outs, normal_min = _blend_launch(meta, [t.detach() for t in lead], [t.detach() for t in maps])
    outs = {job[0]: o for job, o in zip(jobs, outs)}

    # hand the maps to the new material the way `setattr(blended_material, name, map)` would
    blended.device = device
    for name in names:
        if name in outs:
            t = outs[name]
            if name == "normal":
                # base.py:210-217 on the blended normal: min() < 0 -> keep, else remap *2-1 and renormalise
                if with_grad:
                    # (the remap needs its own autograd node: decided on the host from the 4-byte probe result)
                    if not (float(normal_min.item()) < 0):
                        t = normal_ingest(t, 3)
                else:
                    # decided on the device: the remap kernel reads the probe result the blend kernel left there and
                    # returns at once when it is negative - no read-back, the stream is never drained
                    _normal_ingest_cuda(t, 3, out=t, cond_min=normal_min)
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
