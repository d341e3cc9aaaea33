"""ctypes loader for oracle/_build/liboracle_rk4.so (the C restatement oracle).
TEST INFRASTRUCTURE ONLY -- see the header of oracle/rk4_oracle.c."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_rk4.so")
_lib = None
_dp = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rk4_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_evaluate.restype = ctypes.c_int
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def evaluate(params, tspan, B, x0, v, mode=1, tspan_extra=None, uStar=None,
             want_grids=False, nthreads=0, problem_id=0):
    """Batched single_shooting.m:137-150 evaluation on the CPU.
    x0: (n,Bt); v: (nc*K,) shared or (nc*K,Bt); params: (np,) or (np,Bt); B: (K,G).
    Returns dict of arrays in MATLAB (column-major) shapes."""
    L = lib()
    n, nc = 1, 1
    m = n + 1
    x0 = np.asfortranarray(np.atleast_2d(np.asarray(x0, dtype=np.float64)))
    Bt = x0.shape[1]
    tspan = np.ascontiguousarray(tspan, dtype=np.float64)
    N = tspan.size - 1
    G = 2 * N + 1
    Bm = np.asfortranarray(B, dtype=np.float64)
    K = Bm.shape[0]
    v = np.asfortranarray(v, dtype=np.float64)
    vb = int(v.ndim == 2)
    params = np.asfortranarray(params, dtype=np.float64)
    pb = int(params.ndim == 2)
    np_ = params.shape[0]
    N2 = 0
    te = us = None
    if tspan_extra is not None:
        te = np.ascontiguousarray(tspan_extra, dtype=np.float64)
        N2 = te.size - 1
        us = np.ascontiguousarray(np.atleast_1d(uStar), dtype=np.float64)
    out = dict(J=np.empty(Bt), xT=np.empty((n, Bt), order="F"))
    if mode:
        out["dJdv"] = np.empty((nc * K, Bt), order="F")
        out["lam0"] = np.empty((m, Bt), order="F")
    if want_grids:
        out["x"] = np.empty((m, N + 1, Bt), order="F")
        if mode:
            out["lam"] = np.empty((m, N + 1, Bt), order="F")
            out["dJdu"] = np.empty((nc, G, Bt), order="F")
    rc = L.oracle_evaluate(
        ctypes.c_int(problem_id), _p(params), ctypes.c_int(np_), ctypes.c_int(pb),
        ctypes.c_int(N), _p(tspan), ctypes.c_int(N2), _p(te), _p(us),
        ctypes.c_int(K), _p(Bm), _p(x0), _p(v), ctypes.c_int(vb), ctypes.c_long(Bt),
        ctypes.c_int(mode), _p(out["J"]), _p(out["xT"]), _p(out.get("dJdv")), _p(out.get("lam0")),
        _p(out.get("x")), _p(out.get("lam")), _p(out.get("dJdu")), ctypes.c_int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle_evaluate failed rc={rc}")
    return out


def num_threads() -> int:
    return int(lib().oracle_num_threads())
