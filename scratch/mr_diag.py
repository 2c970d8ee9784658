"""Sensitivity of the small multirank test problem (d=64,k=32,n=1001,alpha=.5): fp32 vs fp64, and fp32 with permuted rows."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from onmf_ontf_ndl_b200 import OnmfEngine

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
d, k, n = 64, 32, 1001
X = torch.from_numpy(rng.random((n, d)).astype(np.float32)).to(dev)
W0 = rng.random((d, k))
perm = torch.from_numpy(np.random.default_rng(1).permutation(n)).to(dev)


def run(dt, use_tc=None, permute=False, steps=4):
    eng = OnmfEngine(d, k, alpha=0.5, dtype=dt, device=dev, use_tc=use_tc, collect_stats=True)
    eng.set_state(W0)
    Xs = (X[perm] if permute else X).to(dt).contiguous()
    out = []
    for t in range(1, steps + 1):
        H = eng.step(Xs, float(t)).double().clone()
        if permute:
            Hf = torch.empty_like(H); Hf[perm] = H; H = Hf
        W, A, B, _ = eng.state()
        out.append((H.cpu().numpy(), W.double().cpu().numpy().copy(), A.double().cpu().numpy().copy(), B.double().cpu().numpy().copy()))
    return out, eng.read_stats()


ref, st = run(torch.float64)
print("fp64 stats", st)
for label, kw in (("fp32 tc", {}), ("fp32 simt", dict(use_tc=False)), ("fp32 tc perm", dict(permute=True)), ("fp64 perm", None)):
    o, st = run(torch.float64, permute=True) if kw is None else run(torch.float32, **kw)
    print(label, "flagged", st["flagged"])
    for t, ((H, W, A, B), (Hr, Wr, Ar, Br)) in enumerate(zip(o, ref), 1):
        ecol = np.linalg.norm(H - Hr, axis=1) / np.maximum(np.linalg.norm(Hr, axis=1), 1e-30)
        print("  t=%d  H rel %.2e (worst col %.2e, #cols>1e-3: %d)  W %.2e  A %.2e  B %.2e  nnz/col %.1f" % (
            t, np.linalg.norm(H - Hr) / np.linalg.norm(Hr), ecol.max(), int((ecol > 1e-3).sum()),
            np.abs(W - Wr).max() / np.abs(Wr).max(), np.abs(A - Ar).max() / np.abs(Ar).max(), np.abs(B - Br).max() / np.abs(Br).max(),
            (Hr > 0).sum(1).mean()))
Gr = ref[-1][1].T @ ref[-1][1]
print("cond(G) after 4 steps %.3e ; cond(A) %.3e" % (np.linalg.cond(Gr), np.linalg.cond(ref[-1][2])))
