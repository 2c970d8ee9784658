"""Diagnose fp32 drift on full_cfg2 (d=300,k=49,batch 4000, 500 it): fp32 variants vs the fp64 engine trajectory."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from onmf_ontf_ndl_b200 import OnmfEngine, _lib, _host

name = sys.argv[1] if len(sys.argv) > 1 else "full_cfg2"
g = np.load(os.path.join(ROOT, "tests/golden/%s.npz" % name))
scale = 255.0 if name in ("full_cfg1", "full_cfg2") else 1.0
X = g["pool_u8"].astype(np.float64) / scale
ntr, k, iters, batch, alpha = int(g["n_train"]), int(g["k"]), int(g["iters"]), int(g["batch"]), float(g["alpha"])
X = X[:, :ntr]
d = X.shape[0]
dev = torch.device("cuda", 0)


def per_atom(a, b):
    return float(np.max(np.linalg.norm(a - b, axis=0) / np.maximum(np.linalg.norm(b, axis=0), 1e-300)))


_orig_gram = _lib.gram
def exact_gram(W, G, stream=None, workspace=None):
    with torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream()):
        G.copy_((W.double().T @ W.double()).to(G.dtype))
    return G


def run(dtype, use_tc=None, gram_ws=True, every=3, reserve=None, exact=False):
    _lib.gram = exact_gram if exact else _orig_gram
    np.random.seed(int(g["seed"]))
    W = np.random.rand(d, k)
    pool = _host.to_sample_major(X, dtype, dev)
    eng = OnmfEngine(d, k, alpha=alpha, dtype=dtype, device=dev, use_tc=use_tc, reserve_sms=reserve, collect_stats=True)
    if not gram_ws:
        orig = eng._derive
        eng._derive = lambda W_, G_, hi, lo, st, use_ws=True: orig(W_, G_, hi, lo, st, use_ws=False)
    eng.set_state(W)
    Xb = torch.empty(batch, d, dtype=dtype, device=dev)
    out = {}
    for i in range(1, iters + 1):
        idx = np.random.randint(ntr, size=batch)
        _lib.gather_rows(pool, torch.from_numpy(idx.astype(np.int64)).to(dev), Xb)
        eng.step(Xb, float(i))
        if i % every == 0 or i == iters:
            out[i] = eng.state()[0].double().cpu().numpy().copy()
    return out, eng.read_stats()


ref, st = run(torch.float64)
print("fp64 vs golden:", per_atom(ref[iters], g["W"]), st)
for label, kw in (("fp32 default", {}), ("fp32 simt", dict(use_tc=False))):
    o, st = run(torch.float32, **kw)
    print(label, " ".join("%d:%.1e" % (i, per_atom(o[i], ref[i])) for i in sorted(o)), "| vs golden %.2e" % per_atom(o[iters], g["W"]),
          "flagged", st["flagged"], flush=True)
    e = np.linalg.norm(o[iters] - ref[iters], axis=0) / np.linalg.norm(ref[iters], axis=0)
    print("   worst atoms", np.argsort(e)[-3:], np.sort(e)[-3:], "atom norms", np.linalg.norm(ref[iters], axis=0)[np.argsort(e)[-3:]])

