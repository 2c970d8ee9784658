"""time the dictionary update alone: python profiles/tools/prof_bcd.py d k [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from onmf_ontf_ndl_b200 import _lib
d, k = int(sys.argv[1]), int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
W = torch.rand(d, k, device=dev, generator=g); W /= W.norm(dim=0, keepdim=True)
H = torch.rand(4096, k, device=dev, generator=g) * (torch.rand(4096, k, device=dev, generator=g) < 0.1)
X = torch.rand(4096, d, device=dev, generator=g)
A = (H.T @ H).contiguous(); B = (H.T @ X).contiguous()
out = torch.empty_like(W)
for _ in range(3):
    _lib.update_dict(W, A, B, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    _lib.update_dict(W, A, B, out)
e1.record(); torch.cuda.synchronize()
print("bcd d=%d k=%d TPR=%s: %.4f ms per sweep, checksum %.9e" % (d, k, os.environ.get("ONMF_BCD_TPR", "auto"), e0.elapsed_time(e1) / reps, float(out.double().sum())))
