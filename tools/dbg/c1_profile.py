"""Where does the host time of one 256x256 brdf() call go?  (GPU box)"""
import cProfile, pstats, os, sys, io
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from pypbr_b200.models import CookTorranceBRDF
dev = torch.device("cuda:0")
maps, view, lights, inten = bench.config1_fixture()
mat, leaves = bench.make_material({k: v.to(dev) for k, v in maps.items()}, dev, True)
brdf = CookTorranceBRDF("point")
go = torch.rand(3, 256, 256, device=dev)
vd, ld, idv = view.to(dev), lights.to(dev), inten.to(dev)
def fwd():
    with torch.no_grad():
        return brdf(mat, vd, ld, idv, 1.0)
def fb():
    out = brdf(mat, vd, ld, idv, 1.0)
    torch.autograd.grad(out, leaves, go)
for fn in (fwd, fb):
    for _ in range(200): fn()
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(2000): fn()
    torch.cuda.synchronize(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(fn.__name__); print(s.getvalue()[:3500])
