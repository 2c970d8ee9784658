"""Device-resident steps with the pool stored as fp32 / uint8 / fp16: are the narrow-storage loaders of the fused kernels as fast as the fp32 ones?
    python profiles/tools/prof_storage.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from onmf_ontf_ndl_b200 import OnmfEngine
d, k = 1024, 256
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = torch.device('cuda:0')
g = torch.Generator(device=dev); g.manual_seed(0)
W = torch.rand(d, k, device=dev, generator=g)
p8 = torch.randint(0, 256, (n, d), dtype=torch.uint8, device=dev, generator=g)
pools = {'u8': (p8, 1 / 255.0), 'f32': (p8.float() / 255.0, 1.0), 'f16': ((p8.float() / 255.0).half(), 1.0)}
for name, (pool, sc) in pools.items():
    eng = OnmfEngine(d, k, alpha=1.0, dtype=torch.float32, device=dev, lars_timing=True)
    eng.set_state(W)
    t = 0
    for _ in range(6):
        t += 1; eng.step_pool(pool, None, float(t), n=n, scale=sc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng.main)
    for _ in range(10):
        t += 1; eng.step_pool(pool, None, float(t), n=n, scale=sc)
    eng.flush(); e1.record(eng.main); torch.cuda.synchronize()
    Wf = eng.state()[0]
    print('%s pool: %.3f ms/step, checksum %.6e' % (name, e0.elapsed_time(e1) / 10, float(Wf.double().sum())))
