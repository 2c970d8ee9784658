"""Host-side helpers shared by Online_NMF / Online_NTF: numpy <-> device staging."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

DEFAULT_PRECISION = "fp32"          # production mode; "fp64" = parity mode (the reference computes in float64)


def torch_dtype(precision):
    p = DEFAULT_PRECISION if precision is None else precision
    if p in ("fp32", "float32", torch.float32, np.float32):
        return torch.float32
    if p in ("fp64", "float64", torch.float64, np.float64):
        return torch.float64
    raise ValueError("precision must be 'fp32' or 'fp64', got %r" % (precision,))


def device():
    if not torch.cuda.is_available():
        raise _lib.OnmfKernelError("a CUDA device is required: onmf_ontf_ndl_b200 has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def to_sample_major(X, dtype, dev):
    """numpy (d x n) data matrix (any float dtype / layout) -> device tensor (n x d) of `dtype`,
    transposed + converted by the K1 transpose kernel."""
    if isinstance(X, torch.Tensor):
        src = X.to(dev)
        if src.dtype not in (torch.float32, torch.float64):
            src = src.to(torch.float64)
        src = src.contiguous()
    else:
        a = np.ascontiguousarray(np.asarray(X))
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        src = torch.from_numpy(a).to(dev)
    rows, cols = src.shape
    out = torch.empty(cols, rows, dtype=dtype, device=dev)
    if rows > 0 and cols > 0:
        _lib.transpose(src, out)
    return out


def from_sample_major(Ht):
    """device (n x k) -> numpy float64 (k x n); transposed + widened on the device (K1 transpose), one D2H copy."""
    Ht = Ht.detach()
    n, k = Ht.shape
    if n == 0 or k == 0 or not Ht.is_cuda or Ht.dtype not in (torch.float32, torch.float64):
        return Ht.to(torch.float64).cpu().numpy().T.copy()
    out = torch.empty(k, n, dtype=torch.float64, device=Ht.device)
    _lib.transpose(Ht.contiguous(), out)
    return out.cpu().numpy()


def to_device(M, dtype, dev):
    if isinstance(M, torch.Tensor):
        return M.to(dev, dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(M, dtype=np.float64))).to(dev, dtype)


def to_numpy(T):
    return T.detach().to(torch.float64).cpu().numpy()


def upload_indices(idx, dev):
    """int64 index array (any shape) -> device, through pinned memory with a non-blocking copy."""
    a = np.ascontiguousarray(idx, dtype=np.int64)
    h = torch.from_numpy(a).pin_memory()
    return h.to(dev, non_blocking=True)


def broadcast_run_inputs(group, dev, d, r, W, A, B, idx_all):
    """Multi-GPU training through the reference-facing classes: rank 0's initial state and minibatch index sequence are
    what every rank uses (the reference draws them from numpy's global RNG, which is per process)."""
    import torch.distributed as dist
    src = dist.get_global_rank(group, 0) if group is not None else 0
    flags = torch.tensor([0 if A is None else 1, 0 if idx_all is None else 1,
                          0 if idx_all is None else idx_all.shape[0], 0 if idx_all is None else idx_all.shape[1]],
                         dtype=torch.int64, device=dev)
    dist.broadcast(flags, src, group=group)
    has_ab, has_idx, s0, s1 = [int(v) for v in flags.tolist()]
    Wd = to_device(W, torch.float64, dev)
    dist.broadcast(Wd, src, group=group)
    out = [Wd.cpu().numpy()]
    for M, shape in ((A, (r, r)), (B, (r, d))):
        if not has_ab:
            out.append(None)
            continue
        Md = to_device(M, torch.float64, dev) if M is not None else torch.empty(shape, dtype=torch.float64, device=dev)
        dist.broadcast(Md, src, group=group)
        out.append(Md.cpu().numpy())
    if has_idx:
        I = torch.from_numpy(np.ascontiguousarray(idx_all, dtype=np.int64)).to(dev) if idx_all is not None and \
            idx_all.shape == (s0, s1) else torch.empty(s0, s1, dtype=torch.int64, device=dev)
        dist.broadcast(I, src, group=group)
        out.append(I.cpu().numpy())
    else:
        out.append(None)
    return tuple(out)


def broadcast_indices(group, dev, idx):
    import torch.distributed as dist
    src = dist.get_global_rank(group, 0) if group is not None else 0
    I = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)).to(dev)
    dist.broadcast(I, src, group=group)
    return I.cpu().numpy()
