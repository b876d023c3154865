"""
oracle/pbr_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement (PyTorch eager, CPU tensors, fp32 or fp64) of the reference's per-texel
shading hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` may import this module; nothing under
``pypbr_b200/`` does.  It is the checker, never the thing shipped or measured as "ours".

Why torch and not numpy/C: every arithmetic operation of the reference on this path is an
ATen eager op (SURVEY.md §8c "Third-party arithmetic").  Parity is stated in fp32 with rel 1e-5,
which is only meaningful against the same rounding sequence, so the restatement issues the same
ATen ops in the same order as the reference; the same code in fp64 is the arbiter.

Pinned: ``tests/golden/make_golden.py`` imports the real reference from /root/reference in the
build container, checks that this oracle is BIT-IDENTICAL to it on every golden case (forward
and autograd gradients) and writes the fixtures in ``tests/golden/*.npz`` that the CPU test suite
replays.  The reference's own tests hold no numeric vector for this path, so the
reference itself, executed, is the pin.

Each function cites the reference file:line it follows (paths relative to /root/reference).

Batch / multi-light semantics (not in the reference): a batch is B independent
reference calls; L lights are L independent reference calls with ``return_srgb=False`` combined as
  per_light : out[b, l] = encode(call(b, l))
  accumulate: out[b]    = encode(clamp(sum_l call(b, l), 0, 1))
where ``encode`` is linear_to_srgb when return_srgb else identity.  Both reduce to the reference
for B = L = 1 (clamp and the sRGB clamp are idempotent on [0, 1]).
"""

from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------
# colour space  (pypbr/utils/functions.py:31-66)
# --------------------------------------------------------------------------------------


def srgb_to_linear(tex: Tensor) -> Tensor:
    """pypbr/utils/functions.py:31-47 — clamp, piecewise EOTF by boolean-mask assignment, clamp."""
    t = tex.clamp(0, 1)
    low = t <= 0.04045
    out = torch.zeros_like(t)
    out[low] = t[low] / 12.92
    out[~low] = ((t[~low] + 0.055) / 1.055) ** 2.4
    return out.clamp(0, 1)


def linear_to_srgb(tex: Tensor) -> Tensor:
    """pypbr/utils/functions.py:50-66."""
    t = tex.clamp(0, 1)
    low = t <= 0.0031308
    out = torch.zeros_like(t)
    out[low] = t[low] * 12.92
    out[~low] = 1.055 * torch.pow(t[~low], 1 / 2.4) - 0.055
    return out.clamp(0, 1)


# --------------------------------------------------------------------------------------
# Cook-Torrance, one material, one light  (pypbr/models/cooktorrance.py:68-260)
# --------------------------------------------------------------------------------------


def _ggx_ndf(n: Tensor, h: Tensor, rough: Tensor) -> Tensor:
    """pypbr/models/cooktorrance.py:198-218 — alpha = roughness (not squared)."""
    a2 = rough * rough
    ndh = torch.clamp((n * h).sum(dim=0, keepdim=True), 0.0, 1.0)
    dn = ndh * ndh * (a2 - 1.0) + 1.0
    return a2 / (torch.pi * (dn**2) + 1e-7)


def _g1(ndx: Tensor, rough: Tensor) -> Tensor:
    """pypbr/models/cooktorrance.py:220-235 — Schlick-GGX, k = (r+1)^2/8."""
    r = rough + 1.0
    k = (r**2) / 8.0
    return ndx / (ndx * (1.0 - k) + k + 1e-7)


def shade_linear(
    maps: Dict[str, Tensor],
    view_dir: Tensor,
    light: Tensor,
    intensity: Tensor,
    light_size: Optional[float] = None,
    light_type: str = "point",
    albedo_is_srgb: bool = True,
    specular_is_srgb: bool = True,
) -> Tensor:
    """
    One reference call with return_srgb=False: clamp((diffuse+specular)*radiance, 0, 1), shape (3,H,W).
    Follows pypbr/models/cooktorrance.py:92-177.  ``maps`` holds (C,H,W) tensors under the reference's
    names: albedo, roughness, optional normal, and metallic (metallic workflow wins, :103) or specular.
    """
    dt = maps["albedo"].dtype
    v = F.normalize(view_dir, dim=0)  # :95
    inten = intensity.view(3, 1, 1)  # :96
    rough = maps["roughness"]  # :99
    nmap = maps.get("normal")  # :100

    albedo = maps["albedo"]
    base = srgb_to_linear(albedo) if albedo_is_srgb else albedo  # base.py:262-277
    metallic = maps.get("metallic")
    if metallic is not None:  # :103-107
        f0 = torch.lerp(torch.full_like(base, 0.04), base, metallic)
    elif maps.get("specular") is not None:  # :108-113, diffuse.py:76-91
        spec_map = maps["specular"]
        f0 = srgb_to_linear(spec_map) if specular_is_srgb else spec_map
    else:
        raise ValueError("Material must have either 'metallic' or 'specular' property.")  # :115-118

    _, H, W = base.shape  # :120
    vmap = v.view(3, 1, 1).expand(3, H, W)
    att = 1.0
    if light_type == "directional":  # :125-127
        ldir = F.normalize(light, dim=0)
        lmap = ldir.view(3, 1, 1).expand(3, H, W)
    elif light_type == "point":  # :128-140
        lpos = light.view(3, 1, 1)
        s = light_size or 1.0
        # linspace stays fp32 in the reference even when the maps are fp64 (it promotes afterwards)
        # (device=: the reference builds the grid on `device`, cooktorrance.py:132-133; the CPU default is unchanged)
        x = torch.linspace(-s / 2, s / 2, W, device=base.device)
        y = torch.linspace(-s / 2, s / 2, H, device=base.device)
        yv, xv = torch.meshgrid(y, x, indexing="ij")
        pos = torch.stack([xv, -yv, torch.zeros_like(xv)], dim=0)
        lmap = lpos - pos
        dist = torch.norm(lmap, dim=0, keepdim=True)
        lmap = lmap / (dist + 1e-7)
        att = 1.0 / (dist**2 + 1e-7)
    else:
        raise ValueError("Invalid light_type")

    if nmap is None:  # :143-151
        nmap = torch.tensor([0.0, 0.0, 1.0], dtype=dt, device=base.device).view(3, 1, 1).expand(3, H, W)
    n = F.normalize(nmap, dim=0)  # :153
    h = F.normalize(vmap + lmap, dim=0)  # :154
    cos_t = torch.clamp((h * vmap).sum(dim=0, keepdim=True), 0.0, 1.0)  # :155-157
    fs = f0 + (1.0 - f0) * torch.pow(1.0 - cos_t.expand_as(f0), 5.0)  # :184-196
    ndf = _ggx_ndf(n, h, rough)  # :160
    # geometry_smith (:237-260) recomputes its own N.V / N.L; the values equal :163-164 but autograd
    # accumulates d(normal) through both copies, so the duplicate is kept for bit-identical gradients.
    g_ndv = torch.clamp((n * vmap).sum(dim=0, keepdim=True), 0.0, 1.0)  # :254
    g_ndl = torch.clamp((n * lmap).sum(dim=0, keepdim=True), 0.0, 1.0)  # :255
    g = _g1(g_ndv, rough) * _g1(g_ndl, rough)  # :256-260
    ndv = torch.clamp((n * vmap).sum(dim=0, keepdim=True), 0.0, 1.0)  # :163
    ndl = torch.clamp((n * lmap).sum(dim=0, keepdim=True), 0.0, 1.0)  # :164
    spec = (fs * ndf * g) / (4.0 * ndv * ndl + 1e-7)  # :165-166
    kd = (1.0 - fs) * (1.0 - metallic) if metallic is not None else 1.0 - fs  # :169-172
    diff = kd * base / math.pi  # :174
    rad = inten * (ndl * att)  # :175
    return torch.clamp((diff + spec) * rad, 0.0, 1.0)  # :176-177


def render(
    maps: Dict[str, Tensor],
    view_dir: Tensor,
    lights: Tensor,
    intensities: Tensor,
    light_size: Optional[float] = None,
    light_type: str = "point",
    albedo_is_srgb: bool = True,
    specular_is_srgb: bool = True,
    return_srgb: bool = True,
    accumulate: bool = True,
) -> Tensor:
    """
    Batched / multi-light shading built only from single reference calls (module docstring).
    maps: (C,H,W) or (B,C,H,W) per name.  lights/intensities: (3,) or (L,3).
    Returns (3,H,W) for unbatched single-light input (exactly the reference call), (B,3,H,W) for a
    batch in accumulate mode, (B,L,3,H,W) in per-light mode ((L,3,H,W) when unbatched).
    """
    batched = maps["albedo"].dim() == 4
    multi = lights.dim() == 2
    B = maps["albedo"].shape[0] if batched else 1
    lts = lights if multi else lights.view(1, 3)
    ins = intensities if intensities.dim() == 2 else intensities.view(1, 3).expand(lts.shape[0], 3)
    enc = linear_to_srgb if return_srgb else (lambda c: c)
    outs = []
    for b in range(B):
        mb = {k: (t[b] if batched else t) for k, t in maps.items() if t is not None}
        per = [
            shade_linear(mb, view_dir, lts[l], ins[l], light_size, light_type, albedo_is_srgb, specular_is_srgb)
            for l in range(lts.shape[0])
        ]
        if accumulate or not multi:
            if len(per) == 1:
                outs.append(enc(per[0]))
            else:
                tot = per[0]
                for p in per[1:]:
                    tot = tot + p
                outs.append(enc(torch.clamp(tot, 0.0, 1.0)))
        else:
            outs.append(torch.stack([enc(p) for p in per], dim=0))
    return torch.stack(outs, dim=0) if batched else outs[0]


# --------------------------------------------------------------------------------------
# workflow conversions
# --------------------------------------------------------------------------------------


def metallic_to_specular(albedo: Tensor, metallic: Tensor, albedo_is_srgb: bool = True) -> Tuple[Tensor, Tensor]:
    """pypbr/materials/metallic.py:90-109 — returns (diffuse, specular), both linear."""
    a = srgb_to_linear(albedo) if albedo_is_srgb else albedo
    dielectric = torch.full_like(a, 0.04)
    diffuse = a * (1.0 - metallic)
    specular = dielectric * (1.0 - metallic) + a * metallic
    return diffuse, specular


def specular_to_metallic(albedo: Tensor, specular: Tensor, albedo_is_srgb: bool = True) -> Tuple[Tensor, Tensor]:
    """pypbr/materials/diffuse.py:112-147 — returns (basecolor, metallic[3ch]); uses the RAW specular map."""
    d = srgb_to_linear(albedo) if albedo_is_srgb else albedo
    eps = 1e-6
    num = specular - 0.04
    den = d - 0.04 + eps
    m = torch.clamp(num / (den + eps), 0.0, 1.0)
    m = torch.where(den < eps, torch.zeros_like(m), m)
    b = d / (1.0 - m + eps)
    b = torch.where(m >= 0.95, specular, b)
    b = torch.clamp(b, 0.0, 1.0)
    return b, m


def process_normal_map(nm: Tensor) -> Tensor:
    """pypbr/materials/base.py:191-242 — ingestion quirk applied on every `material.normal = x`."""
    if nm.shape[0] == 2:
        xy = nm * 2 - 1
        x, y = xy[0:1], xy[1:2]
        z = torch.sqrt(torch.clamp(1.0 - (x**2 + y**2), min=1e-6))
        return F.normalize(torch.cat([x, y, z], dim=0), dim=0)
    if nm.shape[0] == 3:
        if nm.min() < 0:
            return nm
        return F.normalize(nm * 2.0 - 1.0, dim=0)
    raise ValueError("Normal map must have 2 or 3 channels.")


# --------------------------------------------------------------------------------------
# blending
# --------------------------------------------------------------------------------------


def blend_normals(n1: Tensor, n2: Tensor, mask: Tensor) -> Tensor:
    """pypbr/blending/functional.py:119-145."""
    a = F.normalize(n1, dim=0)
    b = F.normalize(n2, dim=0)
    return F.normalize(mask * a + (1 - mask) * b, dim=0)


def blend_maps(maps1: Dict[str, Tensor], maps2: Dict[str, Tensor], mask: Tensor) -> Dict[str, Optional[Tensor]]:
    """
    pypbr/blending/functional.py:64-116 (map arithmetic only; the material object and the normal
    re-ingestion through setattr are host logic, see process_normal_map).
    """
    if mask.dim() == 2:
        mask = mask.unsqueeze(0)
    elif mask.dim() != 3 or mask.size(0) != 1:
        raise ValueError("Mask must have shape [1, H, W] or [H, W].")
    out: Dict[str, Optional[Tensor]] = {}
    for name in sorted(set(maps1) | set(maps2)):
        a, b = maps1.get(name), maps2.get(name)
        if a is None and b is None:
            out[name] = None
        elif a is None:
            out[name] = b
        elif b is None:
            out[name] = a
        elif name == "normal":
            out[name] = blend_normals(a, b, mask)
        else:
            out[name] = mask * a + (1 - mask) * b
    return out


def sigmoid_mask(p1: Tensor, p2: Tensor, blend_width: float = 0.1, shift: float = 0.0) -> Tensor:
    """pypbr/blending/functional.py:187-194 (height, with shift) and :233-237 (property, shift = 0)."""
    return torch.sigmoid(((p1 + shift) - p2) / (blend_width + 1e-6))


def property_mask(p1: Tensor, p2: Tensor, blend_width: float = 0.1) -> Tensor:
    """pypbr/blending/functional.py:233-237 — no `+ shift` op at all on this path."""
    return torch.sigmoid((p1 - p2) / (blend_width + 1e-6))


def gradient_mask(H: int, W: int, direction: str = "horizontal") -> Tensor:
    """pypbr/blending/functional.py:267-282."""
    if direction == "horizontal":
        return torch.linspace(0, 1, steps=W).unsqueeze(0).unsqueeze(0).repeat(1, H, 1)
    if direction == "vertical":
        return torch.linspace(0, 1, steps=H).unsqueeze(0).unsqueeze(2).repeat(1, 1, W)
    raise ValueError("Direction must be 'horizontal' or 'vertical'.")


# --------------------------------------------------------------------------------------
# rendering loss (docs/source/tutorials/06_advanced.rst:73-107): MSE of two renders
# --------------------------------------------------------------------------------------


def rendering_loss(pred_maps, target_render: Tensor, view_dir, lights, intensities, **kw) -> Tensor:
    """mean((render(pred) - target)^2) over every element, per-light mode for L > 1."""
    out = render(pred_maps, view_dir, lights, intensities, accumulate=False, **kw)
    return ((out - target_render) ** 2).mean()


# --------------------------------------------------------------------------------------
# SURVEY.md §8f "next" rows: ingestion, index transforms, optimiser step
# --------------------------------------------------------------------------------------


def ingest_image(arr, is_normal: bool = False) -> Tensor:
    """
    MaterialBase._to_tensor for a PIL-style array (pypbr/materials/base.py:122-168): uint8 (H,W[,C]) ->
    TF.to_tensor == permute + float32 / 255; uint16 (H,W) -> float32 / 65535 with a leading channel;
    then, for a normal map, process_normal_map above (base.py:191-242).
    """
    t = torch.as_tensor(arr)
    if t.dtype == torch.uint8:
        if t.dim() == 2:
            t = t.unsqueeze(-1)
        out = t.permute(2, 0, 1).contiguous().to(torch.float32).div(255)   # torchvision F.to_tensor
    else:
        out = (torch.as_tensor(arr.astype("float32")).unsqueeze(0) / 65535.0)
    return process_normal_map(out) if is_normal else out


def flip_horizontal(maps: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """pypbr/materials/base.py:605-621."""
    out = {}
    for name, t in maps.items():
        f = t.flip(-1)
        if name == "normal":
            f = f.clone()
            f[0] = -f[0]
        out[name] = f
    return out


def flip_vertical(maps: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """pypbr/materials/base.py:623-639."""
    out = {}
    for name, t in maps.items():
        f = t.flip(-2)
        if name == "normal":
            f = f.clone()
            f[1] = -f[1]
        out[name] = f
    return out


def roll(maps: Dict[str, Tensor], shift) -> Dict[str, Tensor]:
    """pypbr/materials/base.py:641-655."""
    return {k: torch.roll(t, shift, dims=(1, 2)) for k, t in maps.items()}


def tile(maps: Dict[str, Tensor], n: int) -> Dict[str, Tensor]:
    """pypbr/materials/base.py:521-537."""
    return {k: t.repeat(1, n, n) for k, t in maps.items()}


def adam_fit_steps(params: Dict[str, Tensor], grads_per_step, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, projection=None):
    """
    The optimiser of the inverse-rendering fit.  The reference stops at "Perform backpropagation and optimization
    steps here" (docs/source/tutorials/06_advanced.rst:136-137); the checker is torch.optim.Adam itself
    (single-tensor path) followed by the projection onto the valid range after every step:
    clamp(0, 1) for colour / scalar maps, F.normalize(dim=channel) for the normal map.
    grads_per_step: list of dicts name -> gradient.  Returns the parameters after the last step.
    """
    import torch.nn.functional as F

    projection = projection or {}
    ps = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    opt = torch.optim.Adam(list(ps.values()), lr=lr, betas=betas, eps=eps, foreach=False, fused=False)
    for grads in grads_per_step:
        for k, p in ps.items():
            p.grad = grads[k].clone()
        opt.step()
        with torch.no_grad():
            for k, p in ps.items():
                kind = projection.get(k, "normalize" if k == "normal" else "clamp")
                if kind == "clamp":
                    p.clamp_(0.0, 1.0)
                elif kind == "normalize":
                    p.copy_(F.normalize(p, dim=p.dim() - 3))
    return {k: v.detach() for k, v in ps.items()}
