"""Where does the host-minibatch loop lose time against device-resident steps?  uint8 storage, cfg5.
    python profiles/tools/prof_e2e.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from onmf_ontf_ndl_b200 import OnmfEngine
d, k, n = 1024, 256, 262144
dev = torch.device('cuda:0')
g = torch.Generator(device=dev); g.manual_seed(0)
W = torch.rand(d, k, device=dev, generator=g)
p8 = torch.randint(0, 256, (n, d), dtype=torch.uint8, device=dev, generator=g)
host8 = [p8.cpu().pin_memory(), torch.randint(0, 256, (n, d), dtype=torch.uint8).pin_memory()]
dev8 = [h.to(dev) for h in host8]
W_host = torch.empty(d, k).pin_memory()
for mode, kw in (('resident', {}), ('host', {}), ('host', dict(graph=False)), ('host', dict(graph=False, collect_stats=True)), ('host_noW', dict(graph=False, collect_stats=True)), ('resident', dict(graph=False, collect_stats=True))):
    eng = OnmfEngine(d, k, alpha=1.0, dtype=torch.float32, device=dev, lars_timing=True, **kw)
    eng.set_state(W)
    t = 0
    def one(i):
        global t
        t += 1
        if mode == 'resident':
            eng.step_pool(dev8[i & 1], None, float(t), n=n, scale=1 / 255.0)
        elif mode == 'host':
            eng.step_host(host8[i & 1], float(t), W_host)
        else:
            eng.step_host(host8[i & 1], float(t), None)
    for i in range(8):
        one(i)
    eng.flush(); torch.cuda.synchronize()
    eng.reset_lars_timing()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.main)
    for i in range(20):
        one(i)
    eng.flush(); e1.record(eng.main); torch.cuda.synchronize()
    lm = eng._plan.lars_ms()
    print('%-9s %-45s %.3f ms/step; coder ms: mean %.3f min %.3f max %.3f (%d)' % (mode, str(kw), e0.elapsed_time(e1) / 20, sum(lm) / max(len(lm), 1), min(lm), max(lm), len(lm)))
