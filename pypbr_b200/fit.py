"""
pypbr_b200.fit — the rendering-loss step of an SVBRDF inverse-rendering fit, fused and sharded.

The reference only documents this loop (docs/source/tutorials/06_advanced.rst:73-107):
``RenderingLoss.forward = MSELoss()(brdf(pred, ...), brdf(gt, ...))`` over a DataLoader of materials.
Here one kernel launch (pbr_ct_loss_fwd_bwd) renders the predicted material, compares it with the
target image, reduces the squared error (warp shuffle -> shared -> one atomic per CTA) and writes
d loss / d(albedo, normal, roughness, metallic | specular) - the rendered image never reaches HBM.

Multi-GPU: materials are independent, so the batch is sharded across ranks with no data-path
collective; the only exchange is ONE all-reduce per step over a (1 + 3L)-float buffer holding the
loss sum and the gradient of the shared light intensities (`allreduce_loss_and_shared`).
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _cabi
from .materials import MaterialBase
from .models.cooktorrance import _fill_desc, _out_plane, _prepare


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of `total` materials owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def fused_loss_step(
    material: MaterialBase,
    target: torch.Tensor,
    view_dir: torch.Tensor,
    lights: torch.Tensor,
    intensity: torch.Tensor,
    light_type: str = "point",
    light_size: Optional[float] = None,
    return_srgb: bool = True,
    multi_light: str = "per_light",
    loss_scale: Optional[float] = None,
    want_intensity_grad: bool = False,
    out: Optional[Dict[str, torch.Tensor]] = None,
):
    """
    One fused forward + MSE + backward pass over a (batched) material.

    target: the reference render, same shape as CookTorranceBRDF would return for these arguments.
    loss_scale: factor applied to the gradients (default 1/target.numel(), i.e. nn.MSELoss()).
    out: optional dict of preallocated gradient / scratch buffers to reuse between steps
         (keys d_albedo, d_normal, d_roughness, d_metspec, buf).
    Returns (buf, grads): buf is a device tensor [loss_sum, d_intensity(L*3)...] (loss_sum is the
    UNscaled sum of squared errors; multiply by loss_scale for the mean), grads a dict of tensors.
    """
    lib = _cabi.load()
    cfg, (albedo, normal, roughness, metspec), _leaf, device = _prepare(
        material, material.device, view_dir, lights, intensity, light_type, light_size, return_srgb, multi_light
    )
    _cabi.require_cuda(target, "target")
    target = target.contiguous()
    expect = ((albedo.shape[0],) if cfg.batched else ()) + ((cfg.L,) if cfg.per_light else ()) + (3, *albedo.shape[-2:])
    if tuple(target.shape) != expect:
        raise ValueError(f"target has shape {tuple(target.shape)}, expected {expect}")
    if loss_scale is None:
        loss_scale = 1.0 / target.numel()
    out = out if out is not None else {}

    def buf(key, like):
        t = out.get(key)
        if t is None or t.shape != like.shape or t.device != like.device:
            t = torch.empty(like.shape, dtype=torch.float32, device=device)
            out[key] = t
        return t

    keep: list = []
    d = _fill_desc(cfg, albedo, normal, roughness, metspec, keep)
    g = _cabi.PbrCtGrads()
    d_albedo = buf("d_albedo", albedo)
    d_normal = buf("d_normal", normal) if normal is not None else None
    d_rough = buf("d_roughness", roughness)
    d_met = buf("d_metspec", metspec)
    g.d_albedo, g.d_normal = _cabi.plane(d_albedo), _cabi.plane(d_normal)
    g.d_roughness, g.d_metspec = _cabi.plane(d_rough), _cabi.plane(d_met)
    red = out.get("buf")
    if red is None or red.numel() != 1 + 3 * cfg.L:
        red = torch.empty(1 + 3 * cfg.L, dtype=torch.float32, device=device)
        out["buf"] = red
    red.zero_()
    g.d_intensity = red[1:].data_ptr() if want_intensity_grad else None
    ls = _cabi.PbrCtLoss()
    ls.target, ls.target_sl = _out_plane(target, cfg.per_light, cfg.batched)
    ls.loss_scale = float(loss_scale)
    ls.loss_sum = red.data_ptr()
    with torch.cuda.device(device):
        _cabi.check(
            lib.pbr_ct_loss_fwd_bwd(_cabi.byref(d), _cabi.byref(ls), _cabi.byref(g), _cabi.stream_ptr(device)),
            "pbr_ct_loss_fwd_bwd",
        )
    grads = {"albedo": d_albedo, "normal": d_normal, "roughness": d_rough,
             ("metallic" if cfg.workflow == _cabi.WORKFLOW_METALLIC else "specular"): d_met}
    return red, grads


def allreduce_loss_and_shared(buf: torch.Tensor, group=None) -> torch.Tensor:
    """
    The ONLY collective of the sharded fit: SUM all-reduce of [loss_sum, d_intensity...] (1 + 3L floats)
    over NCCL / NVLink.  Per-material map gradients are never communicated.  No-op without an
    initialised process group (single GPU).
    """
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


class RenderingLoss(nn.Module):
    """
    The tutorial's RenderingLoss (06_advanced.rst:73-107) with the predicted render, the MSE and the
    backward pass fused in one kernel.  `forward(predicted_material, ground_truth_material)` returns the
    scalar loss; calling `.backward()` on it delivers the map gradients computed by that same launch.
    """

    def __init__(self, light_type="point", view_dir=torch.tensor([0.0, 0.0, 1.0]), light_dir=torch.tensor([0.1, 0.1, 1.0]),
                 light_intensity=torch.tensor([1.0, 1.0, 1.0]), light_size: Optional[float] = None):
        super().__init__()
        from .models import CookTorranceBRDF

        self.light_type = light_type
        self.brdf = CookTorranceBRDF(light_type=light_type, multi_light="per_light")
        self.view_dir = view_dir
        self.light_dir = light_dir
        self.light_intensity = light_intensity
        self.light_size = light_size

    def forward(self, predicted_material: MaterialBase, ground_truth_material) -> torch.Tensor:
        if isinstance(ground_truth_material, torch.Tensor):
            target = ground_truth_material
        else:
            with torch.no_grad():
                target = self.brdf(ground_truth_material, self.view_dir, self.light_dir, self.light_intensity, self.light_size)
        names = [k for k in ("albedo", "normal", "roughness", "metallic", "specular")
                 if predicted_material._maps.get(k) is not None]
        leaves = [predicted_material._maps[k] for k in names]
        kw = dict(light_type=self.light_type, light_size=self.light_size, multi_light="per_light")

        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *ls):
                red, grads = fused_loss_step(predicted_material, target, self.view_dir, self.light_dir,
                                             self.light_intensity, **kw)
                ctx.grads = [grads.get(k) for k in names]
                return red[0] / target.numel()

            @staticmethod
            def backward(ctx, g):
                return tuple((gr * g if gr is not None else None) for gr in ctx.grads)

        return _Fn.apply(*leaves)
