"""Online_NTF -- drop-in for the reference's src/ontf.py:19-244 on B200.

Same constructor arguments, methods, return shapes and numpy-float64 in/out conventions:

    Online_NTF(X, n_components=100, iterations=500, sub_iterations=10, batch_size=20, ini_dict=None,
               ini_A=None, ini_B=None, history=0, mode=0, learn_joint_dict=False, alpha=None, beta=None,
               subsample=True)
      .train_dict_single()              -> (W, A, B, code)            src/ontf.py:156-244
      .joint_sparse_code_tensor(X, W)   -> H  (n x r)                 src/ontf.py:59-89
      .update_dict(W, A, B)             -> W1 (d x r)                 src/ontf.py:91-115
      .step(X, A, B, W, t)              -> (H1, A1, B1, W1)           src/ontf.py:117-154
      .history, .code

The random draws (W0 = np.random.rand(d, r), idx = np.random.randint(n, size=batch)) come from numpy's
global RNG on the host in the reference's order (src/ontf.py:213, :230), so a seeded run sees the same
initial dictionary and minibatch sequence as the reference.  Everything else runs on the GPU.

Extra keywords (not in the reference):
  precision     = "fp32" (production, default) | "fp64" (parity mode)
  process_group = a torch.distributed group (default: the WORLD group when torch.distributed is initialised with more
                  than one rank, else single GPU).  Every rank constructs the object with the same X and calls
                  train_dict_single(); rank 0's random draws (W0 and the minibatch indices) are broadcast, each rank
                  codes its contiguous shard of every minibatch (parallel.shard_range), the packed partial sums are
                  all-reduced, and every rank returns the identical (W, A, B)  [src/ontf.py:156-244 data-parallel].
"""
from __future__ import annotations

import numpy as np
import torch

from . import _host, _lib
from .engine import OnmfEngine
from .parallel import default_group, shard_range

DEBUG = False


class Online_NTF():

    def __init__(self,
                 X, n_components=100,
                 iterations=500,
                 sub_iterations=10,
                 batch_size=20,
                 ini_dict=None,
                 ini_A=None,
                 ini_B=None,
                 history=0,
                 mode=0,
                 learn_joint_dict=False,
                 alpha=None,
                 beta=None,
                 subsample=True,
                 precision=None,
                 process_group=None):
        self.X = X
        self.n_components = n_components
        self.batch_size = batch_size
        self.iterations = iterations
        self.sub_iterations = sub_iterations
        self.initial_dict = ini_dict
        self.initial_A = ini_A
        self.initial_B = ini_B
        self.history = history
        self.alpha = alpha
        self.beta = beta
        self.mode = mode
        self.learn_joint_dict = learn_joint_dict
        self.code = np.zeros(shape=(X.shape[1], n_components))     # src/ontf.py:56 (never filled there either)
        self.subsample = subsample
        self.precision = precision
        self._dtype = _host.torch_dtype(precision)
        self.lars_stats = None
        self.process_group = process_group
        self._engines = {}

    # -- helpers ---------------------------------------------------------------------------------
    def _alpha(self):
        return 2 if self.alpha is None else self.alpha       # src/ontf.py:79-81

    def _engine(self, d, r, collect_stats=False, group=None):
        """One engine (streams, step plan, workspaces) per (d, r, group) is kept on the object: drivers call
        step() / joint_sparse_code_tensor() / update_dict() once per patch or per epoch."""
        key = (int(d), int(r), id(group) if group is not None else None)
        eng = self._engines.get(key)
        if eng is None or eng.alpha != float(self._alpha()) or eng.beta != (1.0 if self.beta is None else float(self.beta)):
            eng = OnmfEngine(d, r, alpha=self._alpha(), beta=self.beta, dtype=self._dtype,
                             device=_host.device(), collect_stats=True, process_group=group)
            self._engines[key] = eng
        eng.stats.zero_()
        return eng

    def _unfold_sample_major(self, dev):
        """Matricize self.X (src/ontf.py:203-208) directly into the sample-major device layout (n x d).

        tl_unfold(X, mode) = reshape(moveaxis(X, mode, 0), (X.shape[mode], -1)); with learn_joint_dict the
        data matrix is its transpose, so its sample-major form IS the unfolding (no transpose needed);
        otherwise the sample-major form is the unfolding transposed (one K1 transpose kernel)."""
        X = np.asarray(self.X)
        U = np.reshape(np.moveaxis(X, self.mode, 0), (X.shape[self.mode], -1))   # host view/copy, no arithmetic
        if self.learn_joint_dict:
            # data matrix = U.T (d = prod(other dims), n = X.shape[mode]); sample-major = U
            return _host.to_device(U, self._dtype, dev)
        return _host.to_sample_major(U, self._dtype, dev)

    # -- reference API ---------------------------------------------------------------------------
    def joint_sparse_code_tensor(self, X, W):
        """H (n x r): positive lasso_lars codes of the columns of X (d x n) against W (d x r)."""
        dev = _host.device()
        Xt = _host.to_sample_major(X, self._dtype, dev)
        Wd = _host.to_device(W, self._dtype, dev)
        eng = self._engine(Wd.shape[0], Wd.shape[1])
        Ht = eng.sparse_code(Xt, Wd)
        H = Ht.detach().to(torch.float64).cpu().numpy()
        self.lars_stats = eng.read_stats()
        return H

    def update_dict(self, W, A, B):
        dev = _host.device()
        Wd = _host.to_device(W, self._dtype, dev)
        Ad = _host.to_device(A, self._dtype, dev)
        Bd = _host.to_device(B, self._dtype, dev)
        out = torch.empty_like(Wd)
        _lib.update_dict(Wd, Ad, Bd, out)
        return _host.to_numpy(out)

    def step(self, X, A, B, W, t):
        """One online step on host arrays (src/ontf.py:117-154)."""
        dev = _host.device()
        Xt = _host.to_sample_major(X, self._dtype, dev)
        d, r = np.shape(W)
        eng = self._engine(d, r)
        eng.set_state(W, A, B)
        Ht = eng.step(Xt, float(t))
        Wd, Ad, Bd, _ = eng.state()
        H1 = Ht.detach().to(torch.float64).cpu().numpy()
        self.history = np.float64(t) + 1
        return H1, _host.to_numpy(Ad), _host.to_numpy(Bd), _host.to_numpy(Wd)

    def train_dict_single(self):
        r = self.n_components
        code = self.code
        dev = _host.device()
        pool = self._unfold_sample_major(dev)          # (n x d) resident minibatch pool
        n, d = pool.shape

        if self.initial_dict is None:
            W = np.random.rand(d, r)                   # src/ontf.py:213 (host RNG, reference order)
            print('W.shape', W.shape)
            A = B = None
        else:
            W, A, B = self.initial_dict, self.initial_A, self.initial_B
        t0 = self.history

        group, world, rank = default_group(self.process_group)
        steps = max(int(self.iterations) - 1, 0)
        # The reference draws idx = np.random.randint(n, size=batch) once per step (src/ontf.py:230) and nothing else
        # touches the global RNG inside the loop (lasso_lars draws nothing), so drawing all steps up front is the same
        # stream; the indices go to the device in ONE pinned, asynchronous copy instead of one blocking copy per step.
        idx_all = None
        if self.subsample and steps > 0:
            idx_all = np.stack([np.random.randint(n, size=self.batch_size) for _ in range(steps)]).astype(np.int64)
        if world > 1:
            W, A, B, idx_all = _host.broadcast_run_inputs(group, dev, d, r, W, A, B, idx_all)
        eng = self._engine(d, r, group=group)
        eng.set_state(W, A, B)
        m = self.batch_size if self.subsample else n
        lo, hi = shard_range(m, world, rank)           # this rank's columns of every minibatch (contiguous block)
        idx_d = _host.upload_indices(idx_all[:, lo:hi], dev) if idx_all is not None else None
        for s_ in range(steps):
            i = s_ + 1
            if self.subsample:
                # X_batch = X_unfold[:, idx] (src/ontf.py:231) by reference: the kernels read the pool rows in place
                eng.step_pool(pool, idx_d[s_], float(t0 + i))
            else:
                eng.step(pool[lo:hi], float(t0 + i))
            self.history = np.float64(t0 + i) + 1
        Wd, Ad, Bd, _ = eng.state()
        self.lars_stats = eng.read_stats()
        return _host.to_numpy(Wd), _host.to_numpy(Ad), _host.to_numpy(Bd), code
