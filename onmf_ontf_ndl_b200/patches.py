"""On-device patch pipelines: the drivers build their d x N data matrices with O(N^2)
`np.append` loops; here the image / lattice / graph lives on the GPU, patch positions are drawn on the host in the
reference's RNG order, and one K1 kernel gathers every patch.  Outputs are numpy float64 in the reference's shapes
(or device tensors in the sample-major layout with device=True)."""
from __future__ import annotations

import numpy as np
import torch

from . import _host, _lib


def sample_patch_coords(shape, patch_size, num_patches):
    """Top-left corners in the reference's draw order: a = np.random.choice(H - k), b = np.random.choice(W - k),
    interleaved per patch (image_reconstruction.py:184-186, image_reconstruction_tensor.py:103-104,
    ising_reconstruction.py:57-58).  The last valid offset is excluded, like the reference."""
    k = patch_size
    out = np.empty((num_patches, 2), dtype=np.int32)
    for i in range(num_patches):
        out[i, 0] = np.random.choice(shape[0] - k)
        out[i, 1] = np.random.choice(shape[1] - k)
    return out


def gather_patches(img, coords, patch_size, precision=None, device=False):
    """img (H x W) or (H x W x C), coords (N x 2) -> data matrix X (k*k*C, N), feature f = (row*k + col)*C + ch.
    For a gray image this is `extract_random_patches` of image_reconstruction.py:173-206; for a colour image it is the
    mode-2 joint unfolding of the (k*k, 3, N) tensor of image_reconstruction_tensor.py:87-124 (see patches_to_tensor)."""
    dev = _host.device()
    dtype = _host.torch_dtype(precision)
    A = np.asarray(img, dtype=np.float64)
    A3 = A[:, :, None] if A.ndim == 2 else A
    C = A3.shape[2]
    co = torch.from_numpy(np.ascontiguousarray(np.asarray(coords, dtype=np.int32))).to(dev)
    n = co.shape[0]
    out = torch.empty(n, patch_size * patch_size * C, dtype=dtype, device=dev)
    if n:
        _lib.gather_patches(_host.to_device(A3, dtype, dev), co, patch_size, out)
    return out if device else _host.from_sample_major(out)


def extract_patches_2d(img, patch_size, precision=None, device=False):
    """ALL patches of an image in sklearn's order -- the data matrix the drivers build for reconstruction with
    `extract_patches_2d(data, (k, k)).reshape(N, -1).T` (image_reconstruction.py:163-166, ising_reconstruction.py:185-186):
    corners (i, j), i in 0..H-k, j in 0..W-k (the last offset INCLUDED, unlike the random sampler), row-major.
    Returns X (k*k*C, N) numpy float64, or the sample-major device tensor (N x k*k*C) with device=True."""
    A = np.asarray(img)
    return gather_patches(A, all_patch_coords(A.shape, patch_size), int(patch_size), precision, device)


def all_patch_coords(shape, patch_size, stride=1, include_last=True):
    """Top-left corners of a patch grid, row-major, as an (N x 2) int32 array.  include_last=True: every offset 0..H-k
    (sklearn's extract_patches_2d order); False: `range(0, H - k, stride)` -- the grid of the drivers' reconstruction loop
    (image_reconstruction.py:375-376), which leaves the last offset out."""
    k, s = int(patch_size), int(stride)
    if k > shape[0] or k > shape[1]:
        raise ValueError("patch_size %d larger than the image %s" % (k, tuple(shape[:2])))
    ys = np.arange(0, shape[0] - k + (1 if include_last else 0), s)
    xs = np.arange(0, shape[1] - k + (1 if include_last else 0), s)
    gy, gx = np.meshgrid(ys, xs, indexing="ij")
    return np.stack([gy.reshape(-1), gx.reshape(-1)], 1).astype(np.int32)


def extract_random_patches(img, patch_size, num_patches, precision=None):
    """Drop-in for the drivers' extract_random_patches: gray -> (k*k, N); colour -> tensor (k*k, C, N)."""
    A = np.asarray(img)
    co = sample_patch_coords(A.shape, patch_size, num_patches)
    X = gather_patches(A, co, patch_size, precision)
    return X if A.ndim == 2 else patches_to_tensor(X, A.shape[2])


def patches_to_tensor(X, channels):
    """(k*k*C, N) HWC data matrix -> the reference's (k*k, C, N) patch tensor (a view, no arithmetic)."""
    d, n = X.shape
    return X.reshape(d // channels, channels, n)


def graph_to_csr(G):
    """networkx Graph (or anything with .nodes / .neighbors) or a scipy sparse matrix -> (rowptr int64, colidx int32,
    node list); neighbour lists sorted, both directions present for undirected graphs."""
    if hasattr(G, "tocsr"):
        M = G.tocsr()
        M.sort_indices()
        return M.indptr.astype(np.int64), M.indices.astype(np.int32), list(range(M.shape[0]))
    nodes = list(G.nodes())
    pos = {u: i for i, u in enumerate(nodes)}
    rowptr = np.zeros(len(nodes) + 1, dtype=np.int64)
    cols = []
    for i, u in enumerate(nodes):
        nb = sorted(pos[v] for v in G.neighbors(u))
        cols.extend(nb)
        rowptr[i + 1] = len(cols)
    return rowptr, np.asarray(cols, dtype=np.int32), nodes


def motif_patches(csr, emb, precision=None, device=False):
    """emb (N x kk) node positions (indices into the CSR) of N motif embeddings -> X (kk*kk, N) with
    X[q*kk + r, j] = has_edge(emb[j, q], emb[j, r])   (network_reconstruction_nx.py:302-305, 315-329)."""
    dev = _host.device()
    dtype = _host.torch_dtype(precision)
    rowptr, colidx = csr[0], csr[1]
    e = torch.from_numpy(np.ascontiguousarray(np.asarray(emb, dtype=np.int32))).to(dev)
    n, kk = e.shape
    out = torch.empty(n, kk * kk, dtype=dtype, device=dev)
    if n:
        _lib.motif_patches(torch.from_numpy(np.asarray(rowptr, dtype=np.int64)).to(dev),
                           torch.from_numpy(np.asarray(colidx, dtype=np.int32)).to(dev), e, out)
    return out if device else _host.from_sample_major(out)
