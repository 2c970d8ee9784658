import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onmf_ontf_ndl_b200 import _lib
dev = torch.device('cuda:0')
def rel(a, b): return float((a.double() - b).norm() / b.norm())
def split(x):
    hi = torch.empty_like(x); lo = torch.empty_like(x); _lib.split_tf32(x, hi, lo); return hi, lo
g = torch.Generator(device=dev); g.manual_seed(0)
for (n, d, k) in [(256, 64, 64), (300, 128, 96), (1000, 400, 100), (4096, 1024, 256), (513, 300, 52), (70000, 1024, 256)]:
    X = torch.rand(n, d, device=dev, generator=g); W = torch.rand(d, k, device=dev, generator=g)
    H = torch.rand(n, k, device=dev, generator=g) * (torch.rand(n, k, device=dev, generator=g) < 0.2)
    Xh, Xl = split(X); Wh, Wl = split(W); Hh, Hl = split(H)
    assert torch.equal(Xh + Xl, X)
    Ct = torch.full((n, k), float('nan'), device=dev)
    try:
        _lib.cov_tc(Xh, Xl, Wh, Wl, Ct); torch.cuda.synchronize()
        ref = X.double() @ W.double()
        Cs = torch.empty(n, k, device=dev); _lib.cov(X, W, Cs)
        print(n, d, k, 'cov_tc rel %.2e (simt fp32 %.2e)  nan=%d' % (rel(Ct, ref), rel(Cs, ref), int(torch.isnan(Ct).sum())))
    except Exception as e:
        print(n, d, k, 'cov_tc FAILED', e)
    P = torch.full((k, k + d), float('nan'), device=dev)
    ws = torch.empty(_lib.surrogate_tc_workspace(n, k, d), dtype=torch.uint8, device=dev)
    try:
        _lib.surrogate_partial_tc(Hh, Hl, Xh, Xl, P, ws); torch.cuda.synchronize()
        r1 = H.double().T @ H.double(); r2 = H.double().T @ X.double()
        Ps = torch.empty(k, k + d, device=dev); ws2 = torch.empty(_lib.surrogate_workspace(torch.float32, n, k, d), dtype=torch.uint8, device=dev)
        _lib.surrogate_partial(H, X, Ps, ws2)
        print(n, d, k, 'HtH rel %.2e (simt %.2e)  HtX rel %.2e (simt %.2e) nan=%d' % (rel(P[:, :k], r1), rel(Ps[:, :k], r1), rel(P[:, k:], r2), rel(Ps[:, k:], r2), int(torch.isnan(P).sum())))
    except Exception as e:
        print(n, d, k, 'surrogate_tc FAILED', e)
# timing at cfg5
n, d, k = 262144, 1024, 256
X = torch.rand(n, d, device=dev, generator=g); W = torch.rand(d, k, device=dev, generator=g)
H = torch.rand(n, k, device=dev, generator=g) * (torch.rand(n, k, device=dev, generator=g) < 0.1)
Xh, Xl = split(X); Wh, Wl = split(W); Hh, Hl = split(H)
Ct = torch.empty(n, k, device=dev); P = torch.empty(k, k + d, device=dev)
ws = torch.empty(_lib.surrogate_tc_workspace(n, k, d), dtype=torch.uint8, device=dev)
def timeit(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
print('cfg5 split X ms %.3f' % timeit(lambda: _lib.split_tf32(X, Xh, Xl)))
print('cfg5 cov_tc ms %.3f' % timeit(lambda: _lib.cov_tc(Xh, Xl, Wh, Wl, Ct)))
print('cfg5 surrogate_tc ms %.3f' % timeit(lambda: _lib.surrogate_partial_tc(Hh, Hl, Xh, Xl, P, ws)))
