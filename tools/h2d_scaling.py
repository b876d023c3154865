"""Aggregate host->device copy rate with N ranks copying at once (torchrun): the ceiling the e2e legs of bench.py can reach
on this host.  Each rank: pinned host buffer -> cudaMemcpyAsync in chunks, timed between barriers, max over ranks."""
import os, sys, time, json
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device(f"cuda:{local}"); torch.cuda.set_device(dev)
numa = bench.bind_to_gpu_numa(local) if int(os.environ.get("BIND", "1")) else None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
out = {}
for chunk_mb in (64, 512):
    n = 16 if chunk_mb == 64 else 4
    host = torch.empty(n * chunk_mb * (1 << 20), dtype=torch.uint8).pin_memory()
    host.fill_(1)
    devb = torch.empty(chunk_mb * (1 << 20), dtype=torch.uint8, device=dev)
    def run():
        for i in range(n):
            devb.copy_(host[i * chunk_mb * (1 << 20):(i + 1) * chunk_mb * (1 << 20)], non_blocking=True)
    for direction in ("h2d",):
        run(); torch.cuda.synchronize()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4): run()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gbs = 4 * host.numel() / float(t.item()) / 1e9
        out[f"h2d_{chunk_mb}MB_chunks_per_gpu_GBs"] = gbs
    del host, devb
if rank == 0:
    out["world"] = world; out["aggregate_GBs"] = {k: v * world for k, v in out.items() if k.startswith("h2d")}; out["numa"] = numa
    print(json.dumps(out))
if world > 1: dist.destroy_process_group()
