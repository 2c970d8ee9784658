import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onmf_ontf_ndl_b200 import _lib, OnmfEngine
d, k, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device('cuda:0'); dt = torch.float32
g = torch.Generator(device=dev); g.manual_seed(0)
Xt = torch.rand(n, d, dtype=dt, device=dev, generator=g); W = torch.rand(d, k, dtype=dt, device=dev, generator=g)
eng = OnmfEngine(d, k, alpha=1.0, dtype=dt, device=dev, collect_stats=True)
eng.set_state(W)
for t in range(1, 7):
    eng.step(Xt, float(t))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.stats.zero_()
e0.record(); _lib.lasso_lars(eng.G, eng.Ct[:n], d, 1.0, eng.Ht[:n], eng._ws_lars, stats=eng.stats); e1.record()
torch.cuda.synchronize()
print('lars ms', e0.elapsed_time(e1), eng.read_stats())
