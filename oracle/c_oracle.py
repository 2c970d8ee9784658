"""ctypes loader of oracle/liblars_oracle.so (the plain-C restatement in lars_oracle.c).
TEST INFRASTRUCTURE ONLY -- see the header of lars_oracle.c."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblars_oracle.so")
_lib = None


def build():
    src = os.path.join(_HERE, "lars_oracle.c")
    if not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


def load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.oracle_sparse_code.restype = ctypes.c_int
        _lib.oracle_sparse_code.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                            ctypes.c_int, dp]
        _lib.oracle_update_dict.restype = None
        _lib.oracle_update_dict.argtypes = [dp, dp, dp, ctypes.c_int, ctypes.c_int]
        _lib.oracle_aggregate.restype = None
        _lib.oracle_aggregate.argtypes = [dp, dp, dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double]
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def sparse_code(X, W, alpha, max_iter=1000):
    """H (k x n), positive lasso_lars codes (C restatement)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    W = np.ascontiguousarray(W, dtype=np.float64)
    d, n = X.shape
    k = W.shape[1]
    H = np.zeros((k, n))
    load().oracle_sparse_code(_p(X), _p(W), d, n, k, float(alpha), int(max_iter), _p(H))
    return H


def update_dict(W, A, B):
    W1 = np.array(W, dtype=np.float64, order="C", copy=True)
    A = np.ascontiguousarray(A, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    d, k = W1.shape
    load().oracle_update_dict(_p(W1), _p(A), _p(B), d, k)
    return W1


def step(X, A, B, W, t, alpha, beta=None):
    """reference src/ontf.py:117-154 with the C restatement: returns H (k x n), A1, B1, W1."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    a = 2 if alpha is None else alpha
    H = sparse_code(X, W, a)
    w = float(t) ** (-(1.0 if beta is None else beta))
    A1 = np.array(A, dtype=np.float64, order="C", copy=True)
    B1 = np.array(B, dtype=np.float64, order="C", copy=True)
    d, n = X.shape
    load().oracle_aggregate(_p(A1), _p(B1), _p(H), _p(X), d, n, W.shape[1], w)
    W1 = update_dict(W, A, B)
    return H, A1, B1, W1
