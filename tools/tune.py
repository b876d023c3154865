"""
Time the Cook-Torrance forward / backward kernels of every variant library in pypbr_b200/lib/variants/
directly through the C ABI (GPU box only).  Prints one line per variant and writes gpurun_out/tune.json.
"""
import ctypes
import glob
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pypbr_b200 import _cabi  # noqa: E402
from pypbr_b200.models.cooktorrance import _out_plane  # noqa: E402

sys.path.insert(0, ROOT)
from bench import synth_maps  # noqa: E402

B, H, W, L = int(os.environ.get("TUNE_B", 32)), 1024, 1024, int(os.environ.get("TUNE_L", 1))
dev = torch.device("cuda:0")
PAD = int(os.environ.get("TUNE_PAD", 0))   # extra rows per plane: breaks the power-of-two plane stride


def padded(t):
    if PAD == 0:
        return t
    big = torch.empty(t.shape[0], t.shape[1], H + PAD, W, device=dev)
    big[:, :, :H].copy_(t)
    return big[:, :, :H]


maps = {k: padded(v) for k, v in synth_maps(B, H, W, dev, 3).items()}
PER_LIGHT_SHAPE = (B, L, 3, H, W) if int(os.environ.get("TUNE_PER_LIGHT", 0)) else (B, 3, H, W)
out = torch.empty(PER_LIGHT_SHAPE, device=dev) if len(PER_LIGHT_SHAPE) == 5 else padded(torch.empty(B, 3, H, W, device=dev))
go = torch.rand(PER_LIGHT_SHAPE, device=dev) if len(PER_LIGHT_SHAPE) == 5 else padded(torch.rand(B, 3, H, W, device=dev))
grads = {k: padded(torch.empty_like(v)) for k, v in maps.items()}
import math
if L == 1:
    lights = [0.1, 0.1, 1.0]; inten = [1.0, 1.0, 1.0]
else:
    lights = sum(([0.4 * math.cos(2 * math.pi * l / L), 0.4 * math.sin(2 * math.pi * l / L), 1.0] for l in range(L)), [])
    inten = [1.0 / L] * (3 * L)
hv, hl, hi = _cabi.host_floats([0.0, 0.0, 1.0]), _cabi.host_floats(lights), _cabi.host_floats(inten)

d = _cabi.PbrCtDesc()
d.B, d.H, d.W, d.L = B, H, W, L
PER_LIGHT = int(os.environ.get("TUNE_PER_LIGHT", 0))   # 1: (B,L,3,H,W) output / target, the fused-loss entry point is timed as "bwd"
d.workflow, d.light_type, d.albedo_is_srgb, d.specular_is_srgb, d.return_srgb, d.per_light = 0, 1, 1, 1, 1, PER_LIGHT
d.light_size = 1.0
d.metallic_channels = 1
d.albedo, d.normal, d.roughness, d.metspec = (_cabi.plane(maps[k]) for k in ("albedo", "normal", "roughness", "metallic"))
d.view, d.lights, d.intensity = (ctypes.cast(x, ctypes.c_void_p) for x in (hv, hl, hi))
d.out, d.out_sl = _out_plane(out, bool(PER_LIGHT), True)
g = _cabi.PbrCtGrads()
g.grad_out, g.grad_out_sl = _out_plane(go, bool(PER_LIGHT), True)
if L > 1 and not PER_LIGHT and int(os.environ.get("TUNE_FWD_OUT", 1)):
    g.fwd_out = _out_plane(out, False, True)[0]   # one-pass accumulate backward from the forward launch's output (what autograd hands over)
loss_buf = torch.zeros(1, device=dev)
ls = _cabi.PbrCtLoss()
ls.target, ls.target_sl = _out_plane(go, bool(PER_LIGHT), True)
ls.loss_scale = 1.0 / go.numel()
ls.loss_sum = loss_buf.data_ptr()
g.d_albedo, g.d_normal, g.d_roughness, g.d_metspec = (_cabi.plane(grads[k]) for k in ("albedo", "normal", "roughness", "metallic"))

stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
results = {}
ref_out = ref_g = None
paths = sorted(glob.glob(os.path.join(ROOT, "pypbr_b200", "lib", "variants", "libpbrcuda_*.so")))
if os.environ.get("TUNE_ONLY"):
    paths = [p for p in paths if os.path.basename(p)[len("libpbrcuda_"):-3] in os.environ["TUNE_ONLY"].split(",")]
paths.sort(key=lambda p: (not p.endswith("_scalar.so"), p))  # scalar-lane build first: it is the accuracy reference
texels = B * H * W
for path in paths:
    name = os.path.basename(path)[len("libpbrcuda_"):-3]
    lib = ctypes.CDLL(path)
    for fn in (lib.pbr_ct_forward, lib.pbr_ct_backward):
        fn.restype = ctypes.c_int
    lib.pbr_ct_forward.argtypes = [ctypes.POINTER(_cabi.PbrCtDesc), ctypes.c_void_p]
    lib.pbr_ct_backward.argtypes = [ctypes.POINTER(_cabi.PbrCtDesc), ctypes.POINTER(_cabi.PbrCtGrads), ctypes.c_void_p]
    lib.pbr_ct_loss_fwd_bwd.restype = ctypes.c_int
    lib.pbr_ct_loss_fwd_bwd.argtypes = [ctypes.POINTER(_cabi.PbrCtDesc), ctypes.POINTER(_cabi.PbrCtLoss),
                                        ctypes.POINTER(_cabi.PbrCtGrads), ctypes.c_void_p]

    def fwd():
        rc = lib.pbr_ct_forward(ctypes.byref(d), stream)
        assert rc == 0, rc

    def bwd():
        if PER_LIGHT:
            rc = lib.pbr_ct_loss_fwd_bwd(ctypes.byref(d), ctypes.byref(ls), ctypes.byref(g), stream)
        else:
            rc = lib.pbr_ct_backward(ctypes.byref(d), ctypes.byref(g), stream)
        assert rc == 0, rc

    times = {}
    try:
        fwd(); bwd()
    except AssertionError as e:
        print(f"{name:20s} launch failed rc={e}", flush=True)
        continue
    for label, fn in (("fwd", fwd), ("bwd", bwd)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        times[label] = e0.elapsed_time(e1) / n
    if ref_out is None:
        ref_out = out.clone(); ref_g = {k: v.clone() for k, v in grads.items()}
        dout = 0.0; dg = 0.0
    else:
        dout = float((out - ref_out).abs().max())
        dg = max(float(((grads[k] - ref_g[k]).abs() / (ref_g[k].abs() + ref_g[k].abs().mean())).max()) for k in grads)
    fb, bb = texels * 44 / times["fwd"] / 1e6, texels * 76 / times["bwd"] / 1e6
    tot = texels * L / ((times["fwd"] + times["bwd"]) * 1e-3) / 1e9
    results[name] = dict(fwd_ms=times["fwd"], bwd_ms=times["bwd"], fwd_gbs=fb, bwd_gbs=bb, gtexel_lights=tot,
                         max_abs_out_vs_scalar=dout, max_rel_grad_vs_scalar=dg)
    print(f"{name:20s} fwd {times['fwd']:7.3f} ms {fb:7.0f} GB/s | bwd {times['bwd']:7.3f} ms {bb:7.0f} GB/s | "
          f"{tot:6.2f} Gtexel-lights/s | d_out {dout:.2e} d_grad {dg:.2e}", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"tune_B{B}_L{L}.json"), "w") as f:
    json.dump(results, f, indent=1)
