import sys, time, traceback, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from onmf_ontf_ndl_b200 import _lib, OnmfEngine, Online_NTF
from oracle import onmf_oracle as O
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
dev = torch.device('cuda:0')
def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
def section(name, fn):
    print('=== ' + name, flush=True)
    try:
        fn(); torch.cuda.synchronize()
    except Exception:
        traceback.print_exc()
def t_(x, dt): return torch.from_numpy(np.ascontiguousarray(x)).to(dev, dt)

def test_gemm():
    rng = np.random.default_rng(0)
    for dt in (torch.float64, torch.float32):
        for (n, d, k) in [(300, 100, 25), (1000, 441, 25), (257, 300, 49), (513, 400, 100), (640, 1024, 256)]:
            X = rng.random((n, d)); W = rng.random((d, k)); H = rng.random((n, k)) * (rng.random((n, k)) < 0.2)
            Xt, Wd, Ht = t_(X, dt), t_(W, dt), t_(H, dt)
            Gm = torch.empty(k, k, dtype=dt, device=dev); Ct = torch.empty(n, k, dtype=dt, device=dev)
            _lib.gram(Wd, Gm); _lib.cov(Xt, Wd, Ct)
            P = torch.empty(k, k + d, dtype=dt, device=dev)
            ws = torch.empty(_lib.surrogate_workspace(dt, n, k, d), dtype=torch.uint8, device=dev)
            _lib.surrogate_partial(Ht, Xt, P, ws)
            Pn = P.cpu().numpy().astype(np.float64)
            print(dt, n, d, k, 'gram %.1e cov %.1e HtH %.1e HtX %.1e' % (rel(Gm.cpu().numpy(), W.T @ W), rel(Ct.cpu().numpy(), X @ W), rel(Pn[:, :k], H.T @ H), rel(Pn[:, k:], H.T @ X)))

def test_bcd():
    rng = np.random.default_rng(1)
    for dt in (torch.float64, torch.float32):
        for (d, k) in [(100, 25), (300, 49), (441, 25), (400, 100), (1024, 256), (2700, 25), (77, 3)]:
            W = rng.random((d, k)); H = rng.random((k, 200)); A = H @ H.T / 7; B = H @ rng.random((200, d)) / 7
            Wd, Ad, Bd = t_(W, dt), t_(A, dt), t_(B, dt); out = torch.empty_like(Wd)
            _lib.update_dict(Wd, Ad, Bd, out)
            ref = O.update_dict(W, A, B)
            print(dt, d, k, 'update_dict rel %.2e  per-atom max %.2e' % (rel(out.cpu().numpy(), ref), np.max(np.linalg.norm(out.cpu().numpy() - ref, axis=0) / np.maximum(np.linalg.norm(ref, axis=0), 1e-30))))

def test_lars():
    for name in ['cfg1_renoir_gray', 'cfg1_alpha0', 'cfg2_renoir_color_tensor', 'cfg3_binary_motif', 'cfg4_ising_pm1']:
        g = np.load(os.path.join(G, name + '.npz'))
        X = g['X']; alpha = float(g['alpha']); ns = int(g['n_steps'])
        for step in (0, ns - 1):
            W = g['W0'] if step == 0 else g['W_%d' % (step - 1)]
            Xb = X[:, g['idx'][step]]; Href = g['H_%d' % step]
            for dt in (torch.float64, torch.float32):
                eng = OnmfEngine(W.shape[0], W.shape[1], alpha=alpha, dtype=dt, device=dev, collect_stats=True)
                Ht = eng.sparse_code(t_(Xb.T, dt), t_(W, dt))
                H = Ht.cpu().numpy().astype(np.float64).T
                ce = np.linalg.norm(H - Href, axis=0) / np.maximum(np.linalg.norm(Href, axis=0), 1e-30)
                print(name, 'step', step, dt, 'rel %.2e worstcol %.2e' % (rel(H, Href), ce.max()), eng.read_stats())

def test_lars5():
    g = np.load(os.path.join(G, 'cfg5_synthetic.npz'))
    X = np.random.RandomState(int(g['x_seed'])).rand(1024, 160); W0 = np.random.RandomState(int(g['w0_seed'])).rand(1024, 256)
    Xb = X[:, g['idx'][0]]; Href = g['H_0']
    for dt in (torch.float64, torch.float32):
        eng = OnmfEngine(1024, 256, alpha=1.0, dtype=dt, device=dev, collect_stats=True)
        Ht = eng.sparse_code(t_(Xb.T, dt), t_(W0, dt))
        H = Ht.cpu().numpy().astype(np.float64).T
        print('cfg5 step0', dt, 'rel %.2e' % rel(H, Href), eng.read_stats())

def test_train():
    for name in ['cfg1_renoir_gray', 'cfg4_ising_pm1']:
        g = np.load(os.path.join(G, name + '.npz'))
        X = g['X']; ns = int(g['n_steps']); k = g['W0'].shape[1]
        for dt in (torch.float64, torch.float32):
            eng = OnmfEngine(X.shape[0], k, alpha=float(g['alpha']), dtype=dt, device=dev)
            eng.set_state(g['W0'])
            pool = t_(X.T, dt)
            for i in range(ns):
                idx = torch.from_numpy(g['idx'][i].astype(np.int64)).to(dev)
                Xb = torch.empty(len(idx), X.shape[0], dtype=dt, device=dev)
                _lib.gather_rows(pool, idx, Xb)
                eng.step(Xb, float(i + 1))
            W, A, B, _ = eng.state(); torch.cuda.synchronize()
            Wn = W.cpu().numpy().astype(np.float64)
            pa = np.linalg.norm(Wn - g['W_final'], axis=0) / np.maximum(np.linalg.norm(g['W_final'], axis=0), 1e-30)
            print(name, dt, 'W rel %.2e per-atom max %.2e A rel %.2e B rel %.2e' % (rel(Wn, g['W_final']), pa.max(), rel(A.cpu().numpy(), g['A_final']), rel(B.cpu().numpy(), g['B_final'])))

def test_speed():
    for (d, k, n) in [(100, 25, 100000), (400, 100, 65536), (1024, 256, 32768)]:
        dt = torch.float32
        g = torch.Generator(device=dev); g.manual_seed(0)
        Xt = torch.rand(n, d, dtype=dt, device=dev, generator=g); W = torch.rand(d, k, dtype=dt, device=dev, generator=g)
        eng = OnmfEngine(d, k, alpha=1.0, dtype=dt, device=dev, collect_stats=True)
        eng.set_state(W)
        for t in range(1, 4):
            eng.step(Xt, float(t))
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        evs[0].record()
        _lib.gram(eng.W, eng.G); evs[1].record()
        _lib.cov(Xt, eng.W, eng.Ct[:n]); evs[2].record()
        _lib.lasso_lars(eng.G, eng.Ct[:n], d, 1.0, eng.Ht[:n], eng._ws_lars); evs[3].record()
        _lib.surrogate_partial(eng.Ht[:n], Xt, eng.P[0], eng._ws_sur); evs[4].record()
        _lib.update_dict(eng.W, eng.A, eng.B, eng.W_next); evs[5].record()
        torch.cuda.synchronize()
        ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(5)]
        t0 = time.time()
        for t in range(4, 9):
            eng.step(Xt, float(t))
        torch.cuda.synchronize(); el = (time.time() - t0) / 5
        print('d,k,n', d, k, n, 'ms: gram %.3f cov %.3f lars %.3f partial %.3f bcd %.3f | step %.3f ms -> %.3g samples/s' % (*ts, el * 1e3, n / el), eng.read_stats())

for nm, fn in [('gemm', test_gemm), ('bcd', test_bcd), ('lars', test_lars), ('lars5', test_lars5), ('train', test_train), ('speed', test_speed)]:
    if len(sys.argv) > 1 and nm not in sys.argv[1:]: continue
    section(nm, fn)
