#!/usr/bin/env python
"""
bench.py — CookTorrance forward+backward throughput (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c5]

Workload (N=1): BASELINE.json configs[1] — batch 64 materials 1024x1024, CookTorrance forward+backward,
1 point light, fp32, metallic workflow, sRGB albedo in / sRGB colour out.  A step = one forward pass
(pbr_ct_forward) + one backward pass (pbr_ct_backward) over the whole batch through the public API
(CookTorranceBRDF.__call__ + torch.autograd.grad).  N>1 (torchrun): every rank owns its own 64-material
shard (weak scaling, no data-path collective - materials are independent); value = texel-lights of all
ranks / max-over-ranks device time.

  value     : Gtexel-lights/s, maps resident in HBM, timed with CUDA events on the launch stream
  e2e       : same metric through the public API with the maps in PINNED HOST memory: every step uploads
              the 8 map planes (H2D inside the timed region), runs the fused loss forward+backward
              (pbr_ct_loss_fwd_bwd) and reads the scalar loss back (D2H)
  roofline  : the backward kernel (dominant): algorithmic bytes (76 B/texel) / its own
              CUDA-event duration, against the measured HBM peak of MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the oracle port of the reference's PyTorch eager CPU path
              (oracle/pbr_oracle.py, bit-identical to the reference, tests/golden/make_golden.py) timed on
              the host cores on a bounded sample (one 1024x1024 material per step)
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIGS = {
    # name: (B per GPU, H, W, L, per_light)
    "c2": dict(B=64, H=1024, W=1024, L=1, per_light=False,
               workload="batch 64 materials 1024x1024, CookTorrance forward+backward, 1 point light, fp32"),
    "c3": dict(B=16, H=2048, W=2048, L=16, per_light=False,
               workload="batch 16 materials 2048x2048, 16 point lights, forward+backward shading"),
    "c5": dict(B=512, H=512, W=512, L=8, per_light=True,
               workload="SVBRDF fit step, 512 materials 512x512 per GPU, 8 lights, fused loss forward+backward"),
}
FWD_BYTES, BWD_BYTES = 44, 76  # algorithmic bytes per texel, metallic workflow, accumulate mode


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop_flag = True
        self.t.join(timeout=2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------- inputs
def synth_maps(B, H, W, device, seed, pin=False):
    """albedo U[0,1] sRGB, metallic U[0,1], roughness U[0.2,1], unit normals around +Z."""
    g = torch.Generator(device=device).manual_seed(seed)
    albedo = torch.rand(B, 3, H, W, generator=g, device=device)
    metallic = torch.rand(B, 1, H, W, generator=g, device=device)
    roughness = torch.rand(B, 1, H, W, generator=g, device=device) * 0.8 + 0.2
    n = torch.randn(B, 3, H, W, generator=g, device=device)
    n[:, 0:2] *= 0.3
    n[:, 2] = 1.0
    normal = torch.nn.functional.normalize(n, dim=1)
    del n
    maps = dict(albedo=albedo, normal=normal, roughness=roughness, metallic=metallic)
    if pin:
        maps = {k: v.pin_memory() for k, v in maps.items()}
    return maps


def lights_for(L):
    import math

    if L == 1:
        return torch.tensor([0.1, 0.1, 1.0]), torch.tensor([1.0, 1.0, 1.0])
    pts = [[0.4 * math.cos(2 * math.pi * l / L), 0.4 * math.sin(2 * math.pi * l / L), 1.0] for l in range(L)]
    return torch.tensor(pts), torch.ones(L, 3)


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_sample(H, W, L, reps=2, threads=None):
    """One material, fwd+bwd, through the oracle port of the reference's eager CPU path."""
    from oracle import pbr_oracle as O

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    maps = synth_maps(1, H, W, torch.device("cpu"), 7)
    maps = {k: v[0] for k, v in maps.items()}
    lights, inten = lights_for(min(L, 2))
    if L > 1:
        inten = inten / L
    view = torch.tensor([0.0, 0.0, 1.0])
    go = torch.rand(3, H, W)
    nl = lights.shape[0] if lights.dim() == 2 else 1
    best = None
    for it in range(reps + 1):
        leaves = {k: v.clone().requires_grad_(True) for k, v in maps.items()}
        t0 = time.perf_counter()
        out = O.render(leaves, view, lights, inten, 1.0, "point")
        out.backward(go)
        dt = time.perf_counter() - t0
        if it > 0:
            best = dt if best is None else min(best, dt)
    return (H * W * nl) / best / 1e9, threads, best, nl


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    H, W, L = cfg["H"], cfg["W"], cfg["L"]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    from oracle import pbr_oracle as O

    maps = {k: v[0] for k, v in synth_maps(1, H, W, torch.device("cpu"), 7).items()}
    lights, inten = lights_for(min(L, 2))
    nl = lights.shape[0] if lights.dim() == 2 else 1
    if L > 1:
        inten = inten / L
    view = torch.tensor([0.0, 0.0, 1.0])
    go = torch.rand(3, H, W)

    def step():
        leaves = {k: v.clone().requires_grad_(True) for k, v in maps.items()}
        out = O.render(leaves, view, lights, inten, 1.0, "point")
        out.backward(go)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = H * W * nl * args.steps / dt / 1e9
    sample = f"1 material {H}x{W} x {nl} light(s) per step, fwd+bwd (autograd), of the batch of {cfg['B']}"
    line = {
        "impl": "reference", "metric": "Gtexel-lights/sec CookTorrance fwd+bwd", "value": value, "unit": "Gtexel-lights/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Gtexel-lights/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Gtexel-lights/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(args, cfg):
    import torch.distributed as dist

    from pypbr_b200 import _cabi
    from pypbr_b200.fit import allreduce_loss_and_shared, fused_loss_step
    from pypbr_b200.materials import BasecolorMetallicMaterial
    from pypbr_b200.models import CookTorranceBRDF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the shading path has no CPU fallback)")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _cabi.load()

    B, H, W, L, per_light = cfg["B"], cfg["H"], cfg["W"], cfg["L"], cfg["per_light"]
    fused_fit = args.config == "c5"
    maps = synth_maps(B, H, W, dev, 1000 * 2 + rank)
    mat = BasecolorMetallicMaterial(albedo_is_srgb=True, device=dev)
    leaves = []
    for k, v in maps.items():
        v.requires_grad_(True)
        mat._maps[k] = v
        leaves.append(v)
    lights, inten = lights_for(L)
    if L > 1 and not per_light:
        inten = inten / L
    view = torch.tensor([0.0, 0.0, 1.0])
    brdf = CookTorranceBRDF("point", multi_light="per_light" if per_light else "accumulate")
    out_shape = (B, L, 3, H, W) if per_light else (B, 3, H, W)
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    if fused_fit:
        target = torch.rand(out_shape, generator=g, device=dev)
        bufs = {}
        grad_out = None
    else:
        grad_out = torch.rand(out_shape, generator=g, device=dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    fwd_ms, bwd_ms = [], []

    def step(record=False):
        if fused_fit:
            e0, e1 = ev(), ev()
            e0.record()
            buf, _ = fused_loss_step(mat, target, view, lights, inten, "point", 1.0, multi_light="per_light", out=bufs)
            allreduce_loss_and_shared(buf)
            e1.record()
            if record:
                bwd_ms.append((e0, e1))
            return
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        out = brdf(mat, view, lights, inten, 1.0)
        e1.record()
        torch.autograd.grad(out, leaves, grad_out)
        e2.record()
        if record:
            fwd_ms.append((e0, e1))
            bwd_ms.append((e1, e2))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = _cabi.launch_count()
    t_start, t_end = ev(), ev()
    with ClockSampler(local) as clk:
        t_start.record()
        for _ in range(args.steps):
            step(record=True)
        t_end.record()
        torch.cuda.synchronize()
    launches = _cabi.launch_count() - launches0
    elapsed_ms = t_start.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    barrier()

    texel_lights_per_step = B * H * W * L * world
    value = texel_lights_per_step * args.steps / (elapsed_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    texels = B * H * W
    bwd_avg = sum(a.elapsed_time(b) for a, b in bwd_ms) / len(bwd_ms)
    if fused_fit:
        kbytes = texels * (32 + 12 * L + 32)
        kname = "ct_backward_kernel (fused loss)"
        fwd_avg = None
    else:
        fwd_avg = sum(a.elapsed_time(b) for a, b in fwd_ms) / len(fwd_ms)
        kbytes = texels * (BWD_BYTES if not per_light else 64 + 12 * L)
        kname = "ct_backward_kernel"
    achieved = kbytes / (bwd_avg * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": kbytes, "kernel_ms": bwd_avg}
    if fwd_avg is not None:
        fb = texels * (FWD_BYTES if not per_light else 32 + 12 * L)
        roofline["forward"] = {"kernel": "ct_forward_kernel", "achieved": fb / (fwd_avg * 1e-3) / 1e9, "kernel_ms": fwd_avg,
                               "frac": fb / (fwd_avg * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": fb}
        roofline["fwd_plus_bwd_frac"] = (fb + kbytes) / ((fwd_avg + bwd_avg) * 1e-3) / 1e9 / peak

    # ---------------- e2e: host-resident (pinned) maps, upload + fused loss fwd+bwd + loss readback per step
    e2e = None
    if not args.no_e2e:
        del grad_out
        torch.cuda.empty_cache()
        host = synth_maps(B, H, W, torch.device("cpu"), 5 + rank, pin=True)
        host_mat = BasecolorMetallicMaterial(albedo_is_srgb=True, device=torch.device("cpu"))
        for k, v in host.items():
            host_mat._maps[k] = v
        with torch.no_grad():
            tgt = brdf(mat, view, lights, inten, 1.0).detach()
        bufs2 = {}
        h2d = sum(v.numel() * 4 for v in host.values())

        def e2e_step():
            dmat = BasecolorMetallicMaterial(albedo_is_srgb=True, device=dev)
            for k, v in host_mat._maps.items():
                dmat._maps[k] = v.to(dev, non_blocking=True)
            buf, _ = fused_loss_step(dmat, tgt, view, lights, inten, "point", 1.0,
                                     multi_light="per_light" if per_light else "accumulate", out=bufs2)
            return float(buf[0].item())  # D2H of the loss: 4 bytes, synchronises the step

        for _ in range(2):
            e2e_step()
        barrier()
        n_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": texel_lights_per_step * n_e2e / dt / 1e9, "unit": "Gtexel-lights/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4, "steps": n_e2e, "ms_per_step": dt / n_e2e * 1e3,
               "path": "pinned host maps -> .to(cuda) -> pypbr_b200.fit.fused_loss_step (pbr_ct_loss_fwd_bwd) -> loss.item()"}
        del host, host_mat

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, secs, nl = cpu_reference_sample(H, W, L)
        cpu_baseline = {"value": v, "unit": "Gtexel-lights/s", "cores": cores, "kind": "port",
                        "sample": f"1 material {H}x{W} x {nl} light(s), fwd+bwd, best of 2 after 1 warm-up ({secs:.2f} s)"}

    if rank == 0:
        line = {
            "metric": "Gtexel-lights/sec CookTorrance fwd+bwd", "value": value, "unit": "Gtexel-lights/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "per_gpu_batch": B, "H": H, "W": W, "lights": L,
                       "mode": "per_light fused loss" if fused_fit else "accumulate", "parallelism": f"batch-shard x{world}",
                       "l2": "inputs (2.1 GB per GPU) exceed the 126 MB L2; no flush needed"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
