"""Batched patch reconstruction.

The reference reconstructs an image by looping over a stride grid of patches in Python, coding ONE patch per call
and painting a running-mean canvas pixel by pixel (image_reconstruction.py:358-406: `update_code_within_radius(patch,
W, H0=None, r=None, alpha=1, sub_iter=10, stopping_diff=0.01)` at :384, running mean at :389-392).  Here the whole
grid is one K1 gather, one Gram/covariance product, one batched coder launch (the same projected-gradient iteration
with the reference's per-patch stopping test, or the positive LARS-lasso coder), one product Ht W^T and one
overlap-mean kernel.  Results equal the reference loop's (same H0 stream) up to summation order.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _host, _lib, patches
from .engine import OnmfEngine


def grid_shape(H, W, patch_size, recons_resolution):
    """number of grid rows / columns of `for i in range(0, H - k, res)` (image_reconstruction.py:375-376)."""
    k, s = patch_size, recons_resolution
    return len(range(0, H - k, s)), len(range(0, W - k, s))


def reconstruct_image(A, W, patch_size, recons_resolution=1, alpha=1, sub_iter=10, stopping_diff=0.01, coder="pgd",
                      H0=None, precision=None, return_code=False):
    """A: image (H x W) or (H x W x C); W: dictionary (patch_size^2 * C, r).
    Returns (A_recons, overlap_count[, code (r x n_patches)]) as numpy float64.
    coder="pgd": the reference's projected-gradient coder, one independent run per patch; H0 (r x n_patches) defaults to
    np.random.rand(r, 1) per patch in loop order, like the reference.  coder="lasso_lars": the positive lasso (the
    commented-out alternative at image_reconstruction.py:380-383)."""
    dev = _host.device()
    dtype = _host.torch_dtype(precision)
    A = np.asarray(A, dtype=np.float64)
    squeeze = A.ndim == 2
    A3 = A[:, :, None] if squeeze else A
    Hh, Ww, C = A3.shape
    k = int(patch_size)
    W = np.asarray(W, dtype=np.float64)
    d, r = W.shape
    if d != k * k * C:
        raise ValueError("dictionary has %d rows, expected patch_size^2 * channels = %d" % (d, k * k * C))
    ny, nx = grid_shape(Hh, Ww, k, recons_resolution)
    n = ny * nx
    canvas = torch.zeros(Hh, Ww, C, dtype=dtype, device=dev)
    count = torch.zeros(Hh, Ww, dtype=dtype, device=dev)
    if n == 0:
        out = canvas.cpu().numpy().astype(np.float64)
        return (out[:, :, 0] if squeeze else out), count.cpu().numpy().astype(np.float64)
    gy, gx = np.meshgrid(np.arange(ny) * recons_resolution, np.arange(nx) * recons_resolution, indexing="ij")
    coords = torch.from_numpy(np.stack([gy.reshape(-1), gx.reshape(-1)], 1).astype(np.int32)).to(dev)
    img = _host.to_device(A3, dtype, dev)
    Xt = torch.empty(n, d, dtype=dtype, device=dev)
    _lib.gather_patches(img, coords, k, Xt)                                  # K1
    Wd = _host.to_device(W, dtype, dev)
    if coder == "lasso_lars":
        eng = OnmfEngine(d, r, alpha=alpha, dtype=dtype, device=dev)
        Ht = eng.sparse_code(Xt, Wd, alpha=alpha).clone()                    # K2 + K3
    elif coder == "pgd":
        if H0 is None:
            # same stream as n calls of np.random.rand(r, 1); already sample-major
            Ht = _host.to_device(np.random.rand(n, r), dtype, dev)
        else:
            H0 = np.asarray(H0, dtype=np.float64)
            if H0.shape != (r, n):
                raise ValueError("H0 must have shape (r, n_patches) = (%d, %d)" % (r, n))
            Ht = _host.to_sample_major(H0, dtype, dev)                        # uploaded as it is, transposed on the device
        G = torch.empty(r, r, dtype=dtype, device=dev)
        Ct = torch.empty(n, r, dtype=dtype, device=dev)
        _lib.gram(Wd, G)
        _lib.cov(Xt, Wd, Ct)
        _lib.pgd_code_columns(G, Ct, alpha, sub_iter, stopping_diff, Ht)
    else:
        raise ValueError("coder must be 'pgd' or 'lasso_lars'")
    Wt = torch.empty(r, d, dtype=dtype, device=dev)
    _lib.transpose(Wd, Wt)
    R = torch.empty(n, d, dtype=dtype, device=dev)
    _lib.cov(Ht, Wt, R)                                                      # patch reconstructions (n x d) = Ht W^T
    _lib.patch_grid_mean(R, ny, nx, k, recons_resolution, C, Hh, Ww, canvas, count)
    out = canvas.cpu().numpy().astype(np.float64)
    res = (out[:, :, 0] if squeeze else out, count.cpu().numpy().astype(np.float64))
    if return_code:
        res = res + (_host.from_sample_major(Ht),)
    return res


def reconstruct_from_patches_2d(patches_or_W, image_size, code=None, precision=None):
    """Overlap-averaged image from ALL its patches in sklearn's order -- the last two lines of the drivers' one-shot
    reconstruction (image_reconstruction.py:351-354, image_reconstruction_tensor.py:280-283, ising_reconstruction.py:196-199):

        patches_recons = np.dot(W, code).T.reshape(N, k, k);  img = reconstruct_from_patches_2d(patches_recons, dims)

    Two call forms: `reconstruct_from_patches_2d(patches (N, k, k[, C]), (H, W))` -- sklearn's signature, the patches are
    uploaded -- or `reconstruct_from_patches_2d(W (k*k*C, r), (H, W), code=code (r, N))`, which forms W code on the device
    (K2-style product) and never materialises the N x k x k array on the host.  Both end in the deterministic gather-form
    mean kernel onmf_patch_grid_mean.  Returns numpy float64 (H, W[, C])."""
    dev = _host.device()
    dtype = _host.torch_dtype(precision)
    Hh, Ww = int(image_size[0]), int(image_size[1])
    if code is None:
        P = np.asarray(patches_or_W, dtype=np.float64)
        if P.ndim not in (3, 4) or P.shape[1] != P.shape[2]:
            raise ValueError("patches must have shape (N, k, k) or (N, k, k, C)")
        n, k = P.shape[0], P.shape[1]
        C = 1 if P.ndim == 3 else P.shape[3]
        R = _host.to_device(P.reshape(n, -1), dtype, dev)
    else:
        W = np.asarray(patches_or_W, dtype=np.float64)
        d, r = W.shape
        C = int(image_size[2]) if len(image_size) > 2 else 1
        k = int(round(np.sqrt(d // C)))
        if k * k * C != d:
            raise ValueError("dictionary rows %d are not patch_size^2 * channels (channels = %d)" % (d, C))
        Ht = _host.to_sample_major(np.asarray(code, dtype=np.float64), dtype, dev)
        n = Ht.shape[0]
        if Ht.shape[1] != r:
            raise ValueError("code must have shape (r, N) = (%d, N)" % r)
        Wt = torch.empty(r, d, dtype=dtype, device=dev)
        _lib.transpose(_host.to_device(W, dtype, dev), Wt)
        R = torch.empty(n, d, dtype=dtype, device=dev)
        if n:
            _lib.cov(Ht, Wt, R)
    ny, nx = Hh - k + 1, Ww - k + 1
    if n != ny * nx:
        raise ValueError("expected (H-k+1)*(W-k+1) = %d patches, got %d" % (ny * nx, n))
    canvas = torch.empty(Hh, Ww, C, dtype=dtype, device=dev)
    _lib.patch_grid_mean(R, ny, nx, k, 1, C, Hh, Ww, canvas)
    out = canvas.cpu().numpy().astype(np.float64)
    return out[:, :, 0] if (C == 1 and len(image_size) == 2) else out


def reconstruct_network(G, W, embs, alpha=0, precision=None):
    """Batched form of Network_Reconstructor.reconstruct_network (network_reconstruction_nx.py:444-511).

    The reference walks the motif MCMC and, per step, builds ONE k x k adjacency patch (`get_single_patch_glauber`,
    :331-340), codes it with `SparseCoder(transform_alpha=0, 'lasso_lars', positive_code=True)` (:466-473), forms
    patch_recons = W code (:474-475) and folds every entry into a running-mean weight of the directed edge
    (emb[q], emb[r]) (:477-491).  The walk is sequential and stays with the caller; given its states `embs`
    (recons_iter x k node labels, the embedding AFTER each update, i.e. what get_single_patch_glauber returns), this
    runs all steps at once: K1 motif patches -> K2 Gram/covariances -> K3 positive LARS at alpha -> Ht W^T -> one
    scatter kernel accumulating sum / count per directed pair.

    G: networkx graph (or scipy sparse adjacency, node labels = indices); W: (k*k, r) dictionary.
    Returns (pairs (E x 2 node labels), weight (E,), count (E,)) sorted by (a, b) position in G.nodes -- the content
    of the reference's G_recons / G_overlap_count DiGraphs."""
    dev = _host.device()
    dtype = _host.torch_dtype(precision)
    W = np.asarray(W, dtype=np.float64)
    d, r = W.shape
    embs = np.asarray(embs)
    n, kk = embs.shape
    if kk * kk != d:
        raise ValueError("dictionary has %d rows, expected k^2 = %d" % (d, kk * kk))
    if n == 0:
        return np.empty((0, 2), dtype=object), np.empty(0), np.empty(0, dtype=np.int64)
    rowptr, colidx, nodes = patches.graph_to_csr(G)
    pos = {u: i for i, u in enumerate(nodes)}
    emb_i = np.asarray([[pos[u] for u in row] for row in embs.tolist()], dtype=np.int32).reshape(n, kk)
    e = torch.from_numpy(emb_i).to(dev)
    Xt = torch.empty(n, d, dtype=dtype, device=dev)
    if n:
        _lib.motif_patches(torch.from_numpy(np.asarray(rowptr, dtype=np.int64)).to(dev),
                           torch.from_numpy(np.asarray(colidx, dtype=np.int32)).to(dev), e, Xt)      # K1
    Wd = _host.to_device(W, dtype, dev)
    eng = OnmfEngine(d, r, alpha=alpha, dtype=dtype, device=dev)
    Ht = eng.sparse_code(Xt, Wd, alpha=alpha)                                                        # K2 + K3
    Wt = torch.empty(r, d, dtype=dtype, device=dev)
    _lib.transpose(Wd, Wt)
    R = torch.empty(n, d, dtype=dtype, device=dev)
    if n:
        _lib.cov(Ht, Wt, R)                                                                          # patch_recons rows
    # Table size: a table of 2 * n * k^2 slots always suffices (every entry a distinct pair), but consecutive states of the
    # Glauber walk differ in one node, so a trajectory visits about n * (2k - 1) distinct directed pairs: start there (20 bytes
    # per slot -- the worst-case table of a 10^6-state trajectory at k = 21 would be 20 GB) and grow on the kernel's
    # "table full" report.
    worst = 2 * max(n * d, 1) + 2
    cap = 1024
    while cap < min(worst, 2 * n * 4 * kk + 2):
        cap *= 2
    while True:
        keys = torch.full((cap,), -1, dtype=torch.int64, device=dev)
        sums = torch.zeros(cap, dtype=torch.float64, device=dev)
        cnts = torch.zeros(cap, dtype=torch.int32, device=dev)
        failed = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.edge_scatter_add(R, e, keys, sums, cnts, failed)
        nf = int(failed.item())
        if nf == 0:
            break
        if cap >= worst:
            raise _lib.OnmfKernelError("edge_scatter_add could not place %d entries" % nf)
        del keys, sums, cnts
        cap *= 4
    used = torch.nonzero(cnts > 0).flatten()
    kz = keys[used].cpu().numpy().astype(np.uint64)
    sm = sums[used].cpu().numpy()
    ct = cnts[used].cpu().numpy().astype(np.int64)
    order = np.argsort(kz, kind="stable")
    kz, sm, ct = kz[order], sm[order], ct[order]
    a, b = (kz >> np.uint64(32)).astype(np.int64), (kz & np.uint64(0xffffffff)).astype(np.int64)
    lab = np.asarray(nodes, dtype=object)
    pairs = np.stack([lab[a], lab[b]], axis=1) if len(kz) else np.empty((0, 2), dtype=object)
    return pairs, sm / np.maximum(ct, 1), ct


def simple_graph_edges(pairs, weight):
    """The reference's final rounding (network_reconstruction_nx.py:499-507): undirected edge {a, b} when the mean
    weight of the directed pair rounds to a positive integer.  Returns a set of frozensets."""
    out = set()
    for (a, b), w in zip(pairs.tolist(), np.asarray(weight).tolist()):
        if np.round(w) > 0:
            out.add(frozenset((a, b)))
    return out
