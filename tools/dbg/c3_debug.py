"""Where does the C3-shape (2048^2, 16 lights, accumulate) backward differ from the oracle?  GPU box only."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pbr_oracle as O
from test_gpu_parity import _random_case, _material, _brdf
from pypbr_b200.models import cooktorrance as ct
DEV = torch.device("cuda:0")
H = W = int(os.environ.get("DBG_HW", 2048)); L = int(os.environ.get("DBG_L", 16)); B = 2
maps, _l, _i, g = _random_case(31337, B, H, W, 1)
_m, lights, inten, _g = _random_case(1, None, 4, 4, L)
view = torch.tensor([0.0, 0.0, 1.0])
go = torch.rand(B, 3, H, W, generator=g)
b = 1
mb = {k: v[b] for k, v in maps.items()}
with torch.no_grad():
    cols = [O.shade_linear(mb, view, lights[l], inten[l], 1.0, "point") for l in range(L)]
    S = sum(cols)
S.requires_grad_(True)
ref = O.linear_to_srgb(torch.clamp(S, 0.0, 1.0)); ref.backward(go[b])
lr = {k: v.clone().requires_grad_(True) for k, v in mb.items()}
for l in range(L):
    O.shade_linear(lr, view, lights[l], inten[l], 1.0, "point").backward(S.grad)
for tag, two_pass, generic in (("one-pass", False, False), ("two-pass", True, False)):
    ct.NO_SAVED_OUT = two_pass
    mat, leaves = _material(maps, dict(light_type="point"), requires_grad=True)
    out = _brdf(dict(light_type="point"), False)(mat, view, lights, inten, 1.0)
    out.backward(go.to(DEV))
    o = out[b].detach().cpu()
    print(tag, "fwd max abs err", float((o - ref.detach()).abs().max()))
    for k in maps:
        gg = leaves[k].grad[b].cpu().double(); r = lr[k].grad.double()
        tol = 1e-4 * r.abs() + 1e-4 * r.abs().mean()
        ratio = (gg - r).abs() / tol
        bad = ratio > 1
        idx = int(ratio.argmax()); c, y, x = np.unravel_index(idx, ratio.shape)
        print(f"  {k}: max ratio {float(ratio.max()):.2f}, bad {int(bad.sum())} of {bad.numel()}, worst at c={c} y={y} x={x}: got {float(gg[c,y,x]):.6e} ref {float(r[c,y,x]):.6e}; "
              f"S there {S.detach()[:, y, x].tolist()} out {o[:, y, x].tolist()} rough {float(mb['roughness'][0,y,x]):.4f}")
        if int(bad.sum()):
            ys, xs = np.nonzero(bad.any(0).numpy())
            print("    bad rows", np.unique(ys)[:20], "bad cols", np.unique(xs)[:20], "S range at bad", float(S.detach()[:, ys, xs].min()), float(S.detach()[:, ys, xs].max()))
ct.NO_SAVED_OUT = False
