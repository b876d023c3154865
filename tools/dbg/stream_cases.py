import subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CASE = r'''
import sys, torch
sys.path.insert(0, %r)
from pypbr_b200.materials import BasecolorMetallicMaterial
from pypbr_b200.models import CookTorranceBRDF
B,H,W,lt = %s
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
shp = (lambda c: (B,c,H,W)) if B else (lambda c: (c,H,W))
m = BasecolorMetallicMaterial(device=dev)
m._maps["albedo"] = torch.rand(shp(3), generator=g).to(dev)
m._maps["roughness"] = (torch.rand(shp(1), generator=g)*0.8+0.2).to(dev)
m._maps["metallic"] = torch.rand(shp(1), generator=g).to(dev)
m._maps["normal"] = torch.nn.functional.normalize(torch.randn(shp(3), generator=g)*0.2 + torch.tensor([0,0,1.0]).view(3,1,1), dim=-3).to(dev)
out = CookTorranceBRDF(lt)(m, torch.tensor([0,0,1.0]), torch.tensor([0.1,0.1,1.0]), torch.ones(3), 1.0)
torch.cuda.synchronize()
print("ok", float(out.mean()))
'''
for case, bx in [((None,16,64,"point"),1), ((None,16,64,"point"),32), ((None,16,128,"point"),16), ((None,32,32,"point"),1), ((None,32,32,"point"),32),((None,16,64,"point"),64)]:
    r = subprocess.run([sys.executable, "-c", CASE % (ROOT, repr(case))], capture_output=True, text=True, env=dict(os.environ, PBR_STREAM_MIN_BX=str(bx)))
    print(case, bx, r.stdout.strip()[-60:], "|", r.stderr.strip()[-120:].replace("\n", " "))
