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

from . import _host, _lib
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
            H0 = np.random.rand(n, r).T                                      # same stream as n calls of np.random.rand(r, 1)
        H0 = np.asarray(H0, dtype=np.float64)
        if H0.shape != (r, n):
            raise ValueError("H0 must have shape (r, n_patches) = (%d, %d)" % (r, n))
        Ht = _host.to_device(np.ascontiguousarray(H0.T), dtype, dev)
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
