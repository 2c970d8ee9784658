"""Fast first tier vs general kernel on the prof_lars workload: time of both and equality of the codes.
    python profiles/tools/cmp_fast.py d k n [reps]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from onmf_ontf_ndl_b200 import _lib, OnmfEngine
d, k, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
alpha = float(os.environ.get('CMP_ALPHA', '1.0'))      # (the engine trains with alpha 1; the compared coder calls use this one)
dev = torch.device('cuda:0'); dt = torch.float32
g = torch.Generator(device=dev); g.manual_seed(0)
Xt = torch.rand(n, d, dtype=dt, device=dev, generator=g); W = torch.rand(d, k, dtype=dt, device=dev, generator=g)
eng = OnmfEngine(d, k, alpha=1.0, dtype=dt, device=dev, collect_stats=True)
eng.set_state(W)
for t in range(1, 7):
    eng.step(Xt, float(t))
torch.cuda.synchronize()
out = {}
for fast in (0, 1):
    _lib.set_option(_lib.OPT_LARS_FAST_TIER, fast)
    ms = []
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.stats.zero_()
        eng.Ht.fill_(float('nan'))
        e0.record(); _lib.lasso_lars(eng.G, eng.Ct[:n], d, alpha, eng.Ht[:n], eng._ws_lars, stats=eng.stats); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    H = eng.Ht[:n].clone()
    out[fast] = H
    print('fast=%d lars ms min %.3f med %.3f' % (fast, min(ms), sorted(ms)[len(ms) // 2]),
          'checksum %.9e nnz %d nan %d' % (float(H.double().nan_to_num().sum()), int((H != 0).sum()), int(torch.isnan(H).sum())), eng.read_stats())
dif = (out[0] != out[1])
print('columns differing: %d of %d, max abs diff %.3e' % (int(dif.any(1).sum()), n, float((out[0] - out[1]).abs().max())))
