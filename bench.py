#!/usr/bin/env python
"""
bench.py — CookTorrance forward+backward throughput (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5|aux]

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
               workload="SVBRDF inverse-rendering fit step, 512 materials 512x512 per GPU, 8 lights: fused render+MSE+backward, "
                        "loss all-reduce, fused Adam+projection"),
}
C4 = dict(H=4096, W=4096,
          workload="4096x4096 metallic->diffuse-specular conversion + mask/height blend_materials pipeline")
FWD_BYTES, BWD_BYTES = 44, 76  # algorithmic bytes per texel, metallic workflow, accumulate mode


def measured_traffic(kernel_regex_key):
    """Per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the committed
    `ncu --set full` capture of this same command (profiles/traffic.json, written by tools/ncu_traffic.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel_regex_key)
    except Exception:
        return None


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
def cpu_reference_sample(H, W, L, reps=8, threads=None, budget_s=12.0):
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
    best, n_timed, t_begin = None, 0, time.perf_counter()
    for it in range(reps + 1):
        leaves = {k: v.clone().requires_grad_(True) for k, v in maps.items()}
        t0 = time.perf_counter()
        out = O.render(leaves, view, lights, inten, 1.0, "point")
        out.backward(go)
        dt = time.perf_counter() - t0
        if it > 0:
            best = dt if best is None else min(best, dt)
            n_timed += 1
        if n_timed >= 2 and time.perf_counter() - t_begin > budget_s:
            break
    return (H * W * nl) / best / 1e9, threads, best, nl, n_timed


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
    from pypbr_b200.fit import FusedAdam, allreduce_loss_and_shared, fit_step, fused_loss_step
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
    one_launch = fused_fit and args.fit == "one-launch"
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
        for v in leaves:
            v.requires_grad_(False)
        adam = FusedAdam({k: mat._maps[k] for k in ("albedo", "normal", "roughness", "metallic")}, lr=1e-3)
        adam_ms = []
    else:
        grad_out = torch.rand(out_shape, generator=g, device=dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    fwd_ms, bwd_ms = [], []

    def step(record=False):
        if fused_fit:
            # one full step of the sharded fit: fused render + MSE + backward, ONE all-reduce, fused Adam + projection
            e0, e1, e2 = ev(), ev(), ev()
            e0.record()
            if one_launch:
                # ... all of it in ONE launch (pbr_ct_fit_step: Adam + projection in the loss kernel's epilogue)
                fit_step(mat, adam, target, view, lights, inten, "point", 1.0, scratch=bufs,
                         global_numel=target.numel() * world, fused=True)
                e1.record()
                if record:
                    bwd_ms.append((e0, e1))
                return
            buf, grads = fused_loss_step(mat, target, view, lights, inten, "point", 1.0, multi_light="per_light",
                                         loss_scale=1.0 / (target.numel() * world), out=bufs)
            e1.record()
            allreduce_loss_and_shared(buf)
            adam.step({k: grads[k] for k in adam.params})
            e2.record()
            if record:
                bwd_ms.append((e0, e1))
                adam_ms.append((e1, e2))
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
    if one_launch:
        # maps in (32) + per-light targets (12 L) + maps out (32) + both Adam moments of the 8 planes read (64) and
        # written (64); the epilogue's re-read of the parameters is served by L2 and not counted (ncu: 288 B per texel)
        kbytes = texels * (32 + 12 * L + 32 + 128)
        kname = "ct_backward_kernel<0,3,0> (fused loss + Adam epilogue, cached light geometry; includes the loss all-reduce when n_gpus > 1)"
        fwd_avg = None
    elif fused_fit:
        kbytes = texels * (32 + 12 * L + 32)
        kname = "ct_backward_kernel<0,3,0> (fused loss, cached light geometry)"
        fwd_avg = None
    else:
        fwd_avg = sum(a.elapsed_time(b) for a, b in fwd_ms) / len(fwd_ms)
        # accumulate mode with several lights: + the forward output the one-pass backward reads (12 B per texel)
        kbytes = texels * ((BWD_BYTES + (12 if L > 1 else 0)) if not per_light else 64 + 12 * L)
        kname = "ct_backward_stream<0,2,0> (TMA-fed)" if L == 1 else "ct_backward_kernel<0,3,0> (cached light geometry, one pass from the saved forward output)"
    achieved = kbytes / (bwd_avg * 1e-3) / 1e9
    tr = measured_traffic(f"{args.config}:backward" + (":two-kernel" if fused_fit and not one_launch else ""))
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (tr or {}).get("bytes"), "traffic_source": (tr or {}).get("source"),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": kbytes, "kernel_ms": bwd_avg}
    if tr and tr.get("fma_pipe_busy_pct") is not None and L > 1:
        # the multi-light kernels sit under the FP32 roof, not the HBM one: what the committed ncu capture
        # of this command says about the instruction side
        roofline["compute_side"] = {"fma_pipe_busy_pct": tr["fma_pipe_busy_pct"], "issue_slots_busy_pct": tr.get("issue_active_pct"),
                                    "warp_instructions": tr.get("warp_instructions"), "source": tr.get("source")}
    if fused_fit and not one_launch:
        a_avg = sum(a.elapsed_time(b) for a, b in adam_ms) / len(adam_ms)
        ab = texels * 8 * 28   # 8 parameter planes: read p, g, m, v and write p, m, v
        roofline["adam"] = {"kernel": "adam_kernel (+ the loss all-reduce when n_gpus > 1)", "kernel_ms": a_avg,
                            "algorithmic_bytes_per_launch": ab, "achieved": ab / (a_avg * 1e-3) / 1e9,
                            "frac": ab / (a_avg * 1e-3) / 1e9 / peak}
    if fwd_avg is not None:
        fb = texels * (FWD_BYTES if not per_light else 32 + 12 * L)
        roofline["forward"] = {"kernel": "ct_forward_stream<0,2> (TMA-fed)" if L == 1 else "ct_forward_kernel<0,3> (cached light geometry)", "achieved": fb / (fwd_avg * 1e-3) / 1e9, "kernel_ms": fwd_avg,
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
               "path": "pinned host float32 maps -> .to(cuda) -> pypbr_b200.fit.fused_loss_step (pbr_ct_loss_fwd_bwd) -> loss.item()",
               "bound": f"PCIe: {h2d / (dt / n_e2e) / 1e9:.1f} GB/s host->device"}
        del host, host_mat
        # the same step fed with 8-bit textures (what a dataset on disk holds): 8 bytes per texel cross PCIe instead of
        # 32 and pbr_ingest_image expands them on the device
        from pypbr_b200.materials._ingest import NORMAL3, PLAIN, ingest_uint

        g8 = torch.Generator().manual_seed(77 + rank)
        host8 = {k: torch.randint(0, 256, (B, H, W, c), dtype=torch.uint8, generator=g8).pin_memory()
                 for k, c in (("albedo", 3), ("normal", 3), ("roughness", 1), ("metallic", 1))}
        h2d8 = sum(v.numel() for v in host8.values())

        def e2e8_step():
            dmat = BasecolorMetallicMaterial(albedo_is_srgb=True, device=dev)
            for k, v in host8.items():
                dmat._maps[k] = ingest_uint(v, dev, NORMAL3 if k == "normal" else PLAIN)
            buf, _ = fused_loss_step(dmat, tgt, view, lights, inten, "point", 1.0,
                                     multi_light="per_light" if per_light else "accumulate", out=bufs2)
            return float(buf[0].item())

        for _ in range(2):
            e2e8_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e8_step()
        torch.cuda.synchronize()
        dt8 = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt8], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt8 = float(t.item())
        e2e["uint8_textures"] = {"value": texel_lights_per_step * n_e2e / dt8 / 1e9, "unit": "Gtexel-lights/s",
                                 "h2d_bytes_per_step": h2d8, "d2h_bytes_per_step": 4, "ms_per_step": dt8 / n_e2e * 1e3,
                                 "path": "pinned host uint8 (B,H,W,C) images -> .to(cuda) -> pbr_ingest_image x4 -> "
                                         "pbr_ct_loss_fwd_bwd -> loss.item()"}
        del host8

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, secs, nl, n_timed = cpu_reference_sample(H, W, L)
        cpu_baseline = {"value": v, "unit": "Gtexel-lights/s", "cores": cores, "kind": "port",
                        "sample": f"1 material {H}x{W} x {nl} light(s) of the batch, fwd+bwd through autograd, best of {n_timed} "
                                  f"after 1 warm-up ({secs:.2f} s each); oracle/pbr_oracle.py, bit-identical to the reference"}

    if rank == 0:
        line = {
            "metric": "Gtexel-lights/sec CookTorrance fwd+bwd", "value": value, "unit": "Gtexel-lights/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "per_gpu_batch": B, "H": H, "W": W, "lights": L,
                       "mode": ("per_light fused loss + Adam, " + args.fit) if fused_fit else "accumulate", "parallelism": f"batch-shard x{world}",
                       "l2": "inputs (2.1 GB per GPU) exceed the 126 MB L2; no flush needed"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- config 4
def run_c4(args):
    """
    BASELINE.json configs[3] (SURVEY.md §8d "C4"): two 4096x4096 metallic materials (albedo, normal, roughness,
    metallic, height) -> to_diffuse_specular_material() on each -> blend_materials(d1, d2, "mask", mask) ->
    blend_materials(m1, m2, "height").  Two measurements:
      pipeline : the four public-API calls back to back (host allocation, the normal-min probes and their 4-byte
                 readbacks included), CUDA-event time per pipeline
      kernels  : every kernel on its own through the C ABI on preallocated buffers (launches back to back between
                 two CUDA events), algorithmic bytes / time against the measured HBM peak
    """
    from pypbr_b200 import _cabi
    from pypbr_b200.blending import blend_materials
    from pypbr_b200.materials import BasecolorMetallicMaterial

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the conversion / blend path has no CPU fallback)")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    lib = _cabi.load()
    H, W = C4["H"], C4["W"]
    texels = H * W
    peak, peak_src = measured_peak()
    st = _cabi.stream_ptr(dev)

    def material(seed):
        m = synth_maps(1, H, W, dev, seed)
        g = torch.Generator(device=dev).manual_seed(seed + 50)
        mat = BasecolorMetallicMaterial(albedo_is_srgb=True, device=dev)
        for k, v in m.items():
            mat._maps[k] = v[0]
        mat._maps["height"] = torch.rand(1, H, W, generator=g, device=dev)
        return mat

    m1, m2 = material(4001), material(4002)
    mask = torch.rand(1, H, W, generator=torch.Generator(device=dev).manual_seed(4003), device=dev)

    def pipeline():
        d1 = m1.to_diffuse_specular_material()
        d2 = m2.to_diffuse_specular_material()
        b1, _ = blend_materials(d1, d2, "mask", mask=mask)
        b2, mk = blend_materials(m1, m2, "height", blend_width=0.1)
        return b1, b2, mk

    ev = lambda: torch.cuda.Event(enable_timing=True)
    for _ in range(max(args.warmup, 3)):
        pipeline()
    torch.cuda.synchronize()
    l0 = _cabi.launch_count()
    e0, e1 = ev(), ev()
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            pipeline()
        e1.record()
        torch.cuda.synchronize()
    launches = _cabi.launch_count() - l0
    pipe_ms = e0.elapsed_time(e1) / args.steps

    # kernel-only legs through the C ABI
    d1 = m1.to_diffuse_specular_material()
    d2 = m2.to_diffuse_specular_material()

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    kernels = {}

    def add(name, ms, bytes_per_texel):
        ach = bytes_per_texel * texels / (ms * 1e-3) / 1e9
        kernels[name] = {"ms": ms, "bytes_per_texel": bytes_per_texel, "achieved": ach, "frac": ach / peak}

    o0, o1 = torch.empty(3, H, W, device=dev), torch.empty(3, H, W, device=dev)
    cd = _cabi.PbrConvDesc(1, H, W, 1, _cabi.plane(m1.albedo), _cabi.plane(m1.metallic), _cabi.plane(o0), _cabi.plane(o1))
    add("convert_m2s (K4)", timed(lambda: _cabi.check(lib.pbr_convert_m2s(_cabi.byref(cd), st), "m2s")), 16 + 24)
    cs = _cabi.PbrConvDesc(1, H, W, 0, _cabi.plane(d1.albedo), _cabi.plane(d1.specular), _cabi.plane(o0), _cabi.plane(o1))
    add("convert_s2m (K5)", timed(lambda: _cabi.check(lib.pbr_convert_s2m(_cabi.byref(cs), st), "s2m")), 24 + 24)
    res = torch.full((1,), float("inf"), device=dev)
    nd = _cabi.PbrNormalDesc(1, H, W, 3, _cabi.plane(m1.normal), _cabi.plane(None))
    add("normal_min probe (K7)", timed(lambda: _cabi.check(lib.pbr_normal_min(_cabi.byref(nd), res.data_ptr(), st), "nmin")), 12)

    def blend_desc(a, b, names, mode):
        d = _cabi.PbrBlendDesc()
        d.B, d.H, d.W, d.mask_mode = 1, H, W, mode
        d.blend_width, d.shift, d.apply_shift = 0.1, 0.0, 1
        outs, ch_total = [], 0
        for i, n in enumerate(names):
            ta, tb = a._maps[n], b._maps[n]
            out = torch.empty_like(ta)
            outs.append(out)
            d.maps[i] = _cabi.PbrBlendMap(_cabi.plane(ta), _cabi.plane(tb), _cabi.plane(out), ta.shape[0], int(n == "normal"))
            ch_total += ta.shape[0]
        d.n_maps = len(names)
        d.normal_min = res.data_ptr()
        return d, outs, ch_total

    bd, keep1, ch = blend_desc(d1, d2, ["albedo", "normal", "roughness", "specular"], _cabi.MASK_GIVEN)
    bd.mask = _cabi.plane(mask)
    add("blend, given mask, diffuse-specular pair (K6)", timed(lambda: _cabi.check(lib.pbr_blend(_cabi.byref(bd), st), "blend")),
        4 * (2 * ch + 1) + 4 * ch)
    bh, keep2, ch2 = blend_desc(m1, m2, ["albedo", "height", "metallic", "normal", "roughness"], _cabi.MASK_SIGMOID)
    mask_out = torch.empty(1, H, W, device=dev)
    bh.prop1, bh.prop2, bh.mask_out = _cabi.plane(m1.height), _cabi.plane(m2.height), _cabi.plane(mask_out)
    # the two height planes are read twice by one thread (mask builder + the height map's own lerp): second read is L1/L2
    add("blend, height sigmoid, metallic pair (K6)", timed(lambda: _cabi.check(lib.pbr_blend(_cabi.byref(bh), st), "blend")),
        4 * (2 * ch2) + 4 * ch2 + 4)

    total_bytes = texels * (2 * (40 + 12) + kernels["blend, given mask, diffuse-specular pair (K6)"]["bytes_per_texel"] + 12
                            + kernels["blend, height sigmoid, metallic pair (K6)"]["bytes_per_texel"] + 12)
    dom = max(kernels.items(), key=lambda kv: kv[1]["ms"])
    line = {
        "metric": "Gtexel/s conversion + blend pipeline", "value": texels / (pipe_ms * 1e-3) / 1e9, "unit": "Gtexel/s",
        "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": pipe_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": C4["workload"], "H": H, "W": W,
                   "pipeline": "2 x to_diffuse_specular_material + blend_materials(mask) + blend_materials(height), public API",
                   "l2": "every kernel streams 0.6-2 GB, far above the 126 MB L2; no flush needed"},
        "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": dom[1]["achieved"], "peak": peak, "unit": "GB/s",
                     "frac": dom[1]["frac"], "traffic": None, "peak_source": peak_src,
                     "pipeline_algorithmic_bytes": total_bytes,
                     "pipeline_frac": total_bytes / (pipe_ms * 1e-3) / 1e9 / peak, "kernels": kernels},
        "cpu_baseline": None, "e2e": None, "gpu_launches": int(launches), "clocks": clk.summary(),
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- aux rows
def run_aux(args):
    """Kernel-only timings of the rows either side of the shading path: image ingestion, index transforms, Adam."""
    from pypbr_b200 import _cabi
    from pypbr_b200.fit import FusedAdam
    from pypbr_b200.materials import BasecolorMetallicMaterial
    from pypbr_b200.materials._ingest import NORMAL3, PLAIN, ingest_uint

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    _cabi.load()
    B, H, W = 16, 2048, 2048
    texels = B * H * W
    peak, peak_src = measured_peak()
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    kernels = {}

    def add(name, ms, nbytes):
        ach = nbytes / (ms * 1e-3) / 1e9
        kernels[name] = {"ms": ms, "algorithmic_bytes": nbytes, "achieved": ach, "frac": ach / peak}

    l0 = _cabi.launch_count()
    with ClockSampler(local) as clk:
        g = torch.Generator(device=dev).manual_seed(8)
        rgb8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev, generator=g)
        add("ingest uint8 RGB -> float32 planar", timed(lambda: ingest_uint(rgb8, dev, PLAIN)), texels * (3 + 12))
        add("ingest uint8 RGB -> unit normal", timed(lambda: ingest_uint(rgb8, dev, NORMAL3)), texels * (3 + 12))
        del rgb8
        maps = synth_maps(B, H, W, dev, 9)
        mat = BasecolorMetallicMaterial(albedo_is_srgb=True, device=dev)

        def reset():
            for k, v in maps.items():
                mat._maps[k] = v

        def flip():
            reset()
            mat.flip_horizontal()

        def rollit():
            reset()
            mat.roll((17, -33))

        add("flip_horizontal, 4 maps (8 channels) in one gather", timed(flip), texels * 8 * 8)
        add("roll, 4 maps (8 channels) in one gather", timed(rollit), texels * 8 * 8)
        reset()
        opt = FusedAdam({k: v for k, v in maps.items()}, lr=1e-3)
        grads = {k: torch.rand_like(v) for k, v in maps.items()}
        add("Adam + projection, 4 maps (8 channels)", timed(lambda: opt.step(grads)), texels * 8 * 28)
        del opt, grads
        from pypbr_b200.utils import compute_normal_from_height, rotate_normals

        height = torch.rand(B, 1, H, W, device=dev, generator=g)
        add("compute_normal_from_height (1 ch in, 3 ch out)", timed(lambda: compute_normal_from_height(height, 2.0)), texels * 16)
        nrm = maps["normal"].clone()
        add("rotate_normals in place (3 ch)", timed(lambda: rotate_normals(nrm, 33.0)), texels * 24)
    launches = _cabi.launch_count() - l0
    dom = max(kernels.items(), key=lambda kv: kv[1]["ms"])
    line = {
        "metric": "GB/s of the ingestion / index-transform / optimiser kernels", "value": dom[1]["achieved"], "unit": "GB/s",
        "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": dom[1]["ms"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SURVEY 8f rows on {B} materials {H}x{W}: image ingestion, index transforms, Adam step, normal utilities",
                   "l2": "every kernel streams 1-15 GB, far above the 126 MB L2; no flush needed"},
        "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": dom[1]["achieved"], "peak": peak, "unit": "GB/s",
                     "frac": dom[1]["frac"], "traffic": None, "peak_source": peak_src, "kernels": kernels},
        "cpu_baseline": None, "e2e": None, "gpu_launches": int(launches), "clocks": clk.summary(),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(list(CONFIGS) + ["c4", "aux"]))
    ap.add_argument("--fit", default="one-launch", choices=["one-launch", "two-kernel"],
                    help="c5: pbr_ct_fit_step (Adam in the loss kernel's epilogue) or pbr_ct_loss_fwd_bwd + pbr_adam_step")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.config == "c4":
        if args.impl == "reference":
            raise SystemExit("bench.py: the reference arm is defined for the shading configs (c2/c3/c5) only")
        return run_c4(args)
    if args.config == "aux":
        return run_aux(args)
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
