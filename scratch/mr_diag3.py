import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from onmf_ontf_ndl_b200 import _lib
dev = torch.device("cuda", 0)
z = np.load(os.path.join(ROOT, "scratch", "mr_col.npz"))
G64 = torch.from_numpy(z["G64"]).to(dev)
G32 = G64.float().contiguous()
d, k = 64, 32
for label, c in (("c32", z["c32"]), ("c64->32", z["c64"].astype(np.float32))):
    for reps in (1, 4, 8):
        Ct = torch.from_numpy(np.tile(c[None, :], (reps, 1))).to(dev).contiguous()
        for gl, G in (("G64", G64), ("G32", G32)):
            Ht = torch.zeros(reps, k, device=dev)
            ws = torch.zeros(_lib.lasso_lars_workspace(torch.float32, k, reps), dtype=torch.uint8, device=dev)
            st = torch.zeros(len(_lib.STATS_FIELDS), dtype=torch.int64, device=dev)
            _lib.lasso_lars(G, Ct, d, 0.5, Ht, ws, stats=st)
            torch.cuda.synchronize()
            h = Ht[0].cpu().numpy()
            print(label, "reps", reps, gl, "nz", np.nonzero(h)[0], h[np.nonzero(h)[0]].round(5), dict(zip(_lib.STATS_FIELDS, st.cpu().tolist())))
# fp64 coder on the same covariances
Ct = torch.from_numpy(z["c32"].astype(np.float64)[None, :]).to(dev).contiguous()
Ht = torch.zeros(1, k, device=dev, dtype=torch.float64)
ws = torch.zeros(_lib.lasso_lars_workspace(torch.float64, k, 1), dtype=torch.uint8, device=dev)
_lib.lasso_lars(G64, Ct, d, 0.5, Ht, ws)
print("fp64 coder on c32:", np.nonzero(Ht[0].cpu().numpy())[0])
