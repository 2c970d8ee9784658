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
    """device (n x k) -> numpy float64 (k x n)."""
    return Ht.detach().to(torch.float64).cpu().numpy().T.copy()


def to_device(M, dtype, dev):
    if isinstance(M, torch.Tensor):
        return M.to(dev, dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(M, dtype=np.float64))).to(dev, dtype)


def to_numpy(T):
    return T.detach().to(torch.float64).cpu().numpy()
