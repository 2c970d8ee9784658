"""coder time of every step of a bench-like run (cfg5, pool resampled each step)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from onmf_ontf_ndl_b200 import OnmfEngine
d, k, n = 1024, 256, 262144
dev = torch.device('cuda:0')
g = torch.Generator(device=dev); g.manual_seed(1234)
pool = torch.rand(n, d, device=dev, generator=g)
W0 = torch.rand(d, k, device=dev, generator=g)
eng = OnmfEngine(d, k, alpha=1.0, dtype=torch.float32, device=dev, lars_timing=True, collect_stats=True)
eng.set_state(W0)
prev = None
for t in range(1, 27):
    idx = torch.randint(0, n, (n,), device=dev, generator=g)
    eng.stats.zero_()
    eng.step_pool(pool, idx, float(t))
    st = eng.read_stats()
    lm = eng.read_lars_ms()
    print('step %2d coder %.3f ms  knots/col %.2f  mean_active %.2f  max_active %d  overflow %d flagged %d drops/col %.2f' % (
        t, lm[-1] if lm else -1, st['knots'] / max(st['columns'], 1), st['sum_active'] / max(st['knots'], 1), st['max_active'], st['overflow'], st['flagged'], st['drops'] / max(st['columns'], 1)))
