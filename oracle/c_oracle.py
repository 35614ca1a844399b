"""ctypes binding of oracle/libfsw_oracle.so (C restatement, OpenMP).  Test / baseline infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libfsw_oracle.so")
_lib = None

c_i64, c_vp, c_dbl = ctypes.c_int64, ctypes.c_void_p, ctypes.c_double


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(["make", "-s", "-C", HERE])
        _lib = ctypes.CDLL(LIB)
        for sfx in ("_f32", "_f64"):
            getattr(_lib, "fsw_oracle_project" + sfx).argtypes = [c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp]
            getattr(_lib, "fsw_oracle_embed_forward" + sfx).argtypes = [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_dbl, c_vp, c_vp, c_i64]
            getattr(_lib, "fsw_oracle_embed_backward" + sfx).argtypes = [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_dbl, c_vp, c_vp, c_vp, c_i64]
            getattr(_lib, "fsw_oracle_project_backward" + sfx).argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp]
            for fn in ("project", "embed_forward", "embed_backward", "project_backward"):
                getattr(_lib, "fsw_oracle_%s%s" % (fn, sfx)).restype = None
            getattr(_lib, "fsw_oracle_threads" + sfx).restype = ctypes.c_int
            getattr(_lib, "fsw_oracle_set_threads" + sfx).argtypes = [ctypes.c_int]
            getattr(_lib, "fsw_oracle_set_threads" + sfx).restype = None
    return _lib


def threads():
    return int(load().fsw_oracle_threads_f64())


def set_threads(n=None):
    """use n OpenMP threads (default: every host core) whatever OMP_NUM_THREADS says"""
    n = int(n or os.cpu_count() or 1)
    load().fsw_oracle_set_threads_f64(n)
    load().fsw_oracle_set_threads_f32(n)
    return threads()


def _p(a):
    return None if a is None else a.ctypes.data_as(c_vp)


def embed_forward_backward(X, rowptr, col, W, theta, xi, g=None, thresh=1.0, dtype=np.float64):
    """Core embedding [S, K] (+ gradients dX, dtheta, dxi when g is given) in the C oracle.
    theta [K, d] (no edge features in the C oracle)."""
    lib = load()
    sfx = "_f64" if dtype == np.float64 else "_f32"
    X = np.ascontiguousarray(X, dtype=dtype)
    theta = np.ascontiguousarray(theta, dtype=dtype)
    xi = np.ascontiguousarray(xi, dtype=dtype)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = None if col is None else np.ascontiguousarray(col, dtype=np.int32)
    W = None if W is None else np.ascontiguousarray(W, dtype=dtype)
    N, d = X.shape
    K = theta.shape[0]
    S = len(rowptr) - 1
    max_n = int(np.max(np.diff(rowptr))) if S > 0 else 0
    Xp = np.empty((N, K), dtype=dtype)
    getattr(lib, "fsw_oracle_project" + sfx)(_p(X), N, d, _p(theta), theta.shape[1], K, _p(Xp))
    out = np.empty((S, K), dtype=dtype)
    mass = np.empty(S, dtype=np.float64)
    getattr(lib, "fsw_oracle_embed_forward" + sfx)(_p(Xp), K, _p(rowptr), _p(col), _p(W), S, _p(xi), float(thresh), _p(out), _p(mass), max_n)
    if g is None:
        return out, mass
    g = np.ascontiguousarray(g, dtype=dtype)
    dXp = np.zeros((N, K), dtype=dtype)
    dxi = np.zeros(K, dtype=np.float64)
    getattr(lib, "fsw_oracle_embed_backward" + sfx)(_p(Xp), K, _p(rowptr), _p(col), _p(W), S, _p(xi), float(thresh), _p(g), _p(dXp), _p(dxi), max_n)
    dX = np.empty((N, d), dtype=dtype)
    dtheta = np.zeros_like(theta)
    getattr(lib, "fsw_oracle_project_backward" + sfx)(_p(dXp), _p(X), N, d, _p(theta), theta.shape[1], K, _p(dX), _p(dtheta))
    return out, mass, dX, dtheta, dxi.astype(dtype)
