"""Find the column of the small multirank problem whose fp32 code departs from the fp64 code; dump its inputs."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from onmf_ontf_ndl_b200 import OnmfEngine, _lib

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
d, k, n = 64, 32, 1001
X = torch.from_numpy(rng.random((n, d)).astype(np.float32)).to(dev)
W0 = rng.random((d, k))
e32 = OnmfEngine(d, k, alpha=0.5, dtype=torch.float32, device=dev, collect_stats=True)
e32.set_state(W0)
for t in (1, 2):
    e32.step(X, float(t))
W32 = e32.state()[0].clone()
H32 = e32.sparse_code(X, W32).clone()
Ct32 = e32.Ct[:n].clone()
e64 = OnmfEngine(d, k, alpha=0.5, dtype=torch.float64, device=dev, collect_stats=True)
H64 = e64.sparse_code(X.double(), W32.double()).clone()
Ct64 = e64.Ct[:n].clone()
# fp32 coder fed with the fp64 covariances rounded to fp32
H32b = torch.empty_like(H32)
_lib.lasso_lars(e32._G_scratch, Ct64.float().contiguous(), d, 0.5, H32b, e32._ws_lars)
torch.cuda.synchronize()
for nm, H in (("fp32", H32), ("fp32 with exact cov", H32b)):
    e = (H.double() - H64).norm(dim=1) / H64.norm(dim=1).clamp_min(1e-30)
    j = int(e.argmax())
    print(nm, "worst col", j, float(e[j]), "cols>1e-3", int((e > 1e-3).sum()), "cov err", float((Ct32.double() - Ct64).abs().max()))
    print(" H32", H.double()[j].cpu().numpy().round(6))
    print(" H64", H64[j].cpu().numpy().round(6))
e = (H32.double() - H64).norm(dim=1) / H64.norm(dim=1).clamp_min(1e-30)
j = int(e.argmax())
np.savez(os.path.join(ROOT, "gpurun_out", "mr_col.npz"), W=W32.cpu().numpy(), x=X[j].cpu().numpy(), j=j, H32=H32[j].cpu().numpy(),
         H64=H64[j].cpu().numpy(), c32=Ct32[j].cpu().numpy(), c64=Ct64[j].cpu().numpy(), G64=e32._G_scratch.cpu().numpy())
print(e32.read_stats())
