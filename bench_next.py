"""`python bench.py --workload next [--quick] [--only name,name]`: the K1 gathers, K5 and the SURVEY §8(f) 'next' rows on one
B200 -- one JSON line per kernel: device time (CUDA events on the launching stream, after warm-up, L2 flushed between
repetitions), algorithmic bytes / time against the measured HBM copy peak (MEASURED_PEAKS.json), and, as each line's
cpu_baseline leg, the reference's CPU form of the same operation (the oracle's numpy restatement) timed beside it on a
bounded sample.  Not part of the headline contract line (that is `bench.py` without --workload next).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from onmf_ontf_ndl_b200 import _lib, patches, reconstruct  # noqa: E402
from oracle import onmf_oracle as O  # noqa: E402  (checker / CPU arm only)

dev = torch.device("cuda", 0)


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6500.0


_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    _flush.zero_()


def gpu_ms(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def cpu_s(fn, min_s=0.5):
    fn()
    n, t0 = 0, time.perf_counter()
    while True:
        fn()
        n += 1
        el = time.perf_counter() - t0
        if el > min_s:
            return el / n


ROWS = []


def report(name, shape, ms, byts, cpu_unit_s=None, cpu_note="", units=None, unit_name="samples"):
    med, best = ms
    gbs = byts / (med * 1e-3) / 1e9
    row = {"kernel": name, "shape": shape, "ms_median": med, "ms_best": best, "algorithmic_MB": byts / 1e6, "GB/s": gbs,
           "frac_hbm_peak": gbs / hbm_peak()}
    if units:
        row["M_%s/s" % unit_name] = units / (med * 1e-3) / 1e6
    if cpu_unit_s is not None:
        row["cpu_%s/s" % unit_name] = 1.0 / cpu_unit_s
        row["cpu_note"] = cpu_note
        if units:
            row["gpu_over_cpu"] = (units / (med * 1e-3)) * cpu_unit_s
    ROWS.append(row)
    print(json.dumps(row), flush=True)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    args, _ = ap.parse_known_args(argv)
    only = set(x for x in args.only.split(",") if x)
    want = lambda nm: (not only) or nm in only
    rng = np.random.RandomState(0)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    f32 = torch.float32

    # ------------------------------------------------------------------ K1 gather_rows: X_unfold[:, idx] (src/ontf.py:231)
    if want("gather_rows"):
        for (n, d) in ((262144, 1024), (16384, 400), (65536, 300)):
            pool = torch.rand(n, d, dtype=f32, device=dev, generator=g)
            idx = torch.randint(0, n, (n,), device=dev, generator=g)
            out = torch.empty(n, d, dtype=f32, device=dev)
            ms = gpu_ms(lambda: _lib.gather_rows(pool, idx, out))
            assert torch.equal(out, pool[idx])
            ns = min(n, 8192)
            Xc = np.asarray(rng.rand(d, n if n * d < 5e7 else 32768))
            ic = rng.randint(Xc.shape[1], size=ns)
            c = cpu_s(lambda: Xc[:, ic]) / ns
            report("gather_rows", "n=%d d=%d fp32" % (n, d), ms, 2.0 * n * d * 4 + 8 * n, c, "numpy X[:, idx], float64, 1 core", n)
            del pool, idx, out

    # ------------------------------------------------------------------ K1 gather_patches (image_reconstruction.py:173-206)
    if want("gather_patches"):
        for (Hh, Ww, C, p, n, tag) in ((512, 512, 1, 10, 1000000, "gray 10x10 (cfg1 patches)"),
                                       (512, 512, 3, 10, 400000, "colour 10x10x3 (cfg2)"),
                                       (200, 200, 1, 20, 262144, "Ising 20x20 (cfg4)"),
                                       (1024, 1024, 1, 32, 262144, "32x32 (cfg5 patches)")):
            img = torch.rand(Hh, Ww, C, dtype=f32, device=dev, generator=g)
            co = torch.stack([torch.randint(0, Hh - p, (n,), device=dev, generator=g),
                              torch.randint(0, Ww - p, (n,), device=dev, generator=g)], 1).to(torch.int32).contiguous()
            d = p * p * C
            out = torch.empty(n, d, dtype=f32, device=dev)
            ms = gpu_ms(lambda: _lib.gather_patches(img, co, p, out))
            fill = gpu_ms(lambda: out.fill_(1.0))              # calibration: a write-only stream of the same size (torch fill)
            _lib.gather_patches(img, co, p, out)
            # checker: the oracle's loop on a few patches
            imc, coc = img.cpu().numpy().astype(np.float64), co[:64].cpu().numpy()
            ref = O.gather_patches_gray(imc[:, :, 0], coc, p) if C == 1 else O.gather_patches_color_tensor(imc, coc, p).reshape(d, -1)
            assert np.array_equal(out[:64].cpu().numpy().astype(np.float64).T, ref.astype(np.float32).astype(np.float64))
            ns = 2000
            coc = co[:ns].cpu().numpy()
            fn = (lambda: O.gather_patches_gray(imc[:, :, 0], coc, p)) if C == 1 else (lambda: O.gather_patches_color_tensor(imc, coc, p))
            c = cpu_s(fn) / ns
            # HBM bytes: every output byte once + the image once + the corners (the n*d source reads are L1/L2 hits)
            report("gather_patches", "%s n=%d d=%d" % (tag, n, d), ms, 1.0 * n * d * 4 + Hh * Ww * C * 4 + 8 * n, c,
                   "oracle numpy patch slicing per patch (the reference appends per patch, O(N^2): faster than the reference)", n, "patches")
            ROWS[-1]["write_only_fill_ms"] = fill[0]
            ROWS[-1]["frac_of_fill_rate"] = fill[0] / ms[0]
            print(json.dumps({"kernel": "gather_patches", "shape": tag, "write_only_fill_ms": fill[0], "frac_of_fill_rate": fill[0] / ms[0]}), flush=True)
            del img, co, out

    # ------------------------------------------------------------------ K1 transpose / matricization (src/ontf.py:203-208)
    if want("transpose"):
        for (r, c_) in ((1024, 262144), (300, 65536)):
            src = torch.rand(r, c_, dtype=f32, device=dev, generator=g)
            dst = torch.empty(c_, r, dtype=f32, device=dev)
            ms = gpu_ms(lambda: _lib.transpose(src, dst))
            assert torch.equal(dst, src.t())
            Xc = rng.rand(r, min(c_, 32768))
            c = cpu_s(lambda: np.ascontiguousarray(Xc.T)) / Xc.shape[1]
            report("transpose", "(%d x %d) -> sample-major fp32" % (r, c_), ms, 2.0 * r * c_ * 4, c, "numpy ascontiguousarray(X.T) float64", c_)
            del src, dst

    # ------------------------------------------------------------------ widen (uint8 / fp16 storage -> fp32)
    if want("widen"):
        n, d = 262144, 1024
        src = torch.randint(0, 256, (n, d), dtype=torch.uint8, device=dev, generator=g)
        dst = torch.empty(n, d, dtype=f32, device=dev)
        ms = gpu_ms(lambda: _lib.widen(src, 1.0 / 255.0, dst))
        sc = src[:4096].cpu().numpy()
        c = cpu_s(lambda: sc / 255) / 4096
        report("widen", "uint8 -> fp32 n=%d d=%d" % (n, d), ms, n * d * 5.0, c, "numpy data / 255 (image_reconstruction.py:88)", n)
        del src, dst

    # ------------------------------------------------------------------ (f.4) motif patches (network_reconstruction_nx.py:302-305)
    if want("motif_patches"):
        import scipy.sparse as sp
        nn, deg, kk = 20000, 50, 21
        n = 20000 if args.quick else 200000
        rows = np.repeat(np.arange(nn), deg)
        cols = rng.randint(nn, size=nn * deg)
        M = sp.coo_matrix((np.ones(nn * deg), (rows, cols)), shape=(nn, nn)).tocsr()
        M = ((M + M.T) > 0).astype(np.float64).tocsr()
        M.sort_indices()
        rowptr = torch.from_numpy(M.indptr.astype(np.int64)).to(dev)
        colidx = torch.from_numpy(M.indices.astype(np.int32)).to(dev)
        emb = torch.randint(0, nn, (n, kk), dtype=torch.int32, device=dev, generator=g)
        out = torch.empty(n, kk * kk, dtype=f32, device=dev)
        ms = gpu_ms(lambda: _lib.motif_patches(rowptr, colidx, emb, out))
        adj = [set(M.indices[M.indptr[i]:M.indptr[i + 1]].tolist()) for i in range(nn)]
        ec = emb[:200].cpu().numpy()
        ref = O.motif_patches(adj, ec)
        assert np.array_equal(out[:200].cpu().numpy().astype(np.float64).T, ref)
        c = cpu_s(lambda: O.motif_patches(adj, ec)) / 200
        report("motif_patches", "k=21 (d=441) n=%d, graph %d nodes mean degree %.0f" % (n, nn, M.nnz / nn), ms,
               n * kk * kk * 4.0 + n * kk * 4.0, c, "oracle python loop over (q, r) with set lookups (networkx has_edge is slower)", n, "patches")
        del out, emb

    # ------------------------------------------------------------------ (f.1) patch_grid_mean (image_reconstruction.py:389-392)
    if want("patch_grid_mean"):
        for (Hh, Ww, C, p, s) in ((512, 512, 1, 10, 1), (512, 512, 3, 10, 2)):
            ny, nx = reconstruct.grid_shape(Hh, Ww, p, s)
            n, d = ny * nx, p * p * C
            R = torch.rand(n, d, dtype=f32, device=dev, generator=g)
            canvas = torch.empty(Hh, Ww, C, dtype=f32, device=dev)
            count = torch.empty(Hh, Ww, dtype=f32, device=dev)
            ms = gpu_ms(lambda: _lib.patch_grid_mean(R, ny, nx, p, s, C, Hh, Ww, canvas, count))
            report("patch_grid_mean", "%dx%dx%d p=%d stride=%d (%d patches)" % (Hh, Ww, C, p, s, n), ms, n * d * 4.0 + Hh * Ww * (C + 1) * 4.0,
                   None, "", n, "patches")
            del R

    # ------------------------------------------------------------------ (f.1) whole image reconstruction vs the reference loop
    if want("reconstruct_image"):
        Hh = Ww = 128 if args.quick else 512
        p, r = 10, 25
        A = rng.rand(Hh, Ww)
        W = rng.rand(p * p, r)
        W /= np.maximum(1, np.linalg.norm(W, axis=0))
        ny, nx = reconstruct.grid_shape(Hh, Ww, p, 1)
        H0 = rng.rand(r, ny * nx)
        for coder in ("pgd", "lasso_lars"):
            reconstruct.reconstruct_image(A[:64, :64], W, p, 1, coder=coder, H0=H0[:, :54 * 54] if coder == "pgd" else None)
            walls = []
            for _ in range(5):                                 # host wall clock (uploads from pageable numpy arrays): best of 5
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out = reconstruct.reconstruct_image(A, W, p, 1, coder=coder, H0=H0 if coder == "pgd" else None)
                torch.cuda.synchronize()
                walls.append(time.perf_counter() - t0)
            wall = min(walls)
            # reference loop on a crop (python loop, one coder call per patch)
            crop = 40
            nyc, nxc = reconstruct.grid_shape(crop, crop, p, 1)
            t0 = time.perf_counter()
            O.reconstruct_image_loop(A[:crop, :crop], W, p, 1, 1, 10, 0.01, H0[:, :nyc * nxc],
                                     coder=None if coder == "pgd" else (lambda patch, h0: O.sparse_code_lars(patch, W, 1)))
            cpu = (time.perf_counter() - t0) / (nyc * nxc)
            row = {"kernel": "reconstruct_image(%s)" % coder, "shape": "%dx%d gray, p=10, r=25, stride 1: %d patches" % (Hh, Ww, ny * nx),
                   "wall_ms_host_to_host": wall * 1e3, "wall_ms_median_of_5": float(np.median(walls)) * 1e3, "M_patches/s": ny * nx / wall / 1e6, "cpu_patches/s": 1.0 / cpu,
                   "cpu_note": "oracle restatement of the reference loop (image_reconstruction.py:375-392) on a %dx%d crop, 1 core" % (crop, crop),
                   "gpu_over_cpu": ny * nx / wall * cpu}
            ROWS.append(row)
            print(json.dumps(row), flush=True)

    # ------------------------------------------------------------------ (f.1) network reconstruction: scatter-mean of the edge weights
    if want("edge_scatter"):
        kk = 21
        n = 5000 if args.quick else 100000
        nn = 20000
        R = torch.rand(n, kk * kk, dtype=f32, device=dev, generator=g)
        emb = torch.randint(0, nn, (n, kk), dtype=torch.int32, device=dev, generator=g)
        cap = 1
        while cap < 2 * n * kk * kk + 2:
            cap *= 2
        keys = torch.empty(cap, dtype=torch.int64, device=dev)
        sums = torch.empty(cap, dtype=torch.float64, device=dev)
        cnts = torch.empty(cap, dtype=torch.int32, device=dev)
        failed = torch.zeros(1, dtype=torch.int32, device=dev)

        def run():
            keys.fill_(-1)
            sums.zero_()
            cnts.zero_()
            _lib.edge_scatter_add(R, emb, keys, sums, cnts, failed)
        ms = gpu_ms(run, reps=5)
        assert int(failed.item()) == 0
        report("edge_scatter_add (+ table reset)", "k=21 n=%d states, %d entries, table %d slots" % (n, n * kk * kk, cap), ms,
               n * kk * kk * 4.0 + cap * 20.0 * 2, None, "", n, "states")

    # ------------------------------------------------------------------ (f.3) surrogate-error read-out, K5 dictionary update
    if want("bcd"):
        for (d, k) in ((1024, 256), (400, 100), (100, 25)):
            Wt = torch.rand(d, k, dtype=f32, device=dev, generator=g)
            H = torch.rand(k, 2 * k, dtype=f32, device=dev, generator=g)
            A = H @ H.t()
            B = torch.rand(k, d, dtype=f32, device=dev, generator=g) * k
            Wo = torch.empty_like(Wt)
            ms = gpu_ms(lambda: _lib.update_dict(Wt, A, B, Wo))
            Wc, Ac, Bc = Wt.cpu().numpy().astype(np.float64), A.cpu().numpy().astype(np.float64), B.cpu().numpy().astype(np.float64)
            c = cpu_s(lambda: O.update_dict(Wc, Ac, Bc))
            row = {"kernel": "update_dict (K5)", "shape": "d=%d k=%d" % (d, k), "ms_median": ms[0], "ms_best": ms[1],
                   "algorithmic_MB": (2.0 * d * k + k * k + k * d) * 4 / 1e6, "us_per_atom": ms[0] * 1e3 / k, "cpu_ms": c * 1e3,
                   "cpu_note": "oracle numpy loop over atoms (src/ontf.py:109-113), 1 core", "gpu_over_cpu": c * 1e3 / ms[0]}
            ROWS.append(row)
            print(json.dumps(row), flush=True)

    # ------------------------------------------------------------------ secondary coder: batched projected gradient
    if want("pgd"):
        d, k, n = 100, 25, 253009
        Wt = torch.rand(d, k, dtype=f32, device=dev, generator=g)
        Wt /= Wt.norm(dim=0).clamp_min(1.0)
        Xt = torch.rand(n, d, dtype=f32, device=dev, generator=g)
        G = torch.empty(k, k, dtype=f32, device=dev)
        Ct = torch.empty(n, k, dtype=f32, device=dev)
        _lib.gram(Wt, G)
        _lib.cov(Xt, Wt, Ct)
        H0 = torch.rand(n, k, dtype=f32, device=dev, generator=g)
        Ht = torch.empty_like(H0)

        def run():
            Ht.copy_(H0)
            _lib.pgd_code_columns(G, Ct, 1.0, 10, 0.01, Ht)
        ms = gpu_ms(run)
        Wc, Xc, Hc = Wt.cpu().numpy().astype(np.float64), Xt[:200].cpu().numpy().astype(np.float64).T, H0[:200].cpu().numpy().astype(np.float64).T
        t0 = time.perf_counter()
        for j in range(200):
            O.update_code_within_radius(Xc[:, j:j + 1], Wc, Hc[:, j:j + 1].copy(), None, 1, 10, 0.01)
        c = (time.perf_counter() - t0) / 200
        report("pgd_code_columns (+ H0 copy)", "d=100 k=25 n=%d, sub_iter 10" % n, ms, n * k * 4 * 3.0, c,
               "oracle update_code_within_radius per patch (src/onmf.py:233-271), 1 core", n, "patches")

    out = os.path.join("gpurun_out", "prof_next.json")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(ROWS, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
