"""Time the coder alone at a given shape after six online steps (steady-state, unit-norm learned dictionary).
    python profiles/tools/prof_lars.py d k n [reps]
ONMF_B200_LIB=<variant .so> selects an experimental build of the same library (make variant ...)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from onmf_ontf_ndl_b200 import _lib, OnmfEngine
d, k, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device('cuda:0'); dt = torch.float32
g = torch.Generator(device=dev); g.manual_seed(0)
Xt = torch.rand(n, d, dtype=dt, device=dev, generator=g); W = torch.rand(d, k, dtype=dt, device=dev, generator=g)
eng = OnmfEngine(d, k, alpha=1.0, dtype=dt, device=dev, collect_stats=True)
eng.set_state(W)
for t in range(1, 7):
    eng.step(Xt, float(t))
torch.cuda.synchronize()
ms = []
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.stats.zero_()
    e0.record(); _lib.lasso_lars(eng.G, eng.Ct[:n], d, 1.0, eng.Ht[:n], eng._ws_lars, stats=eng.stats); e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
H = eng.Ht[:n]
print('lib', os.path.basename(_lib.LIB_PATH), 'lars ms min %.3f med %.3f' % (min(ms), sorted(ms)[len(ms) // 2]),
      'checksum %.9e nnz %d' % (float(H.double().sum()), int((H != 0).sum())), eng.read_stats())
