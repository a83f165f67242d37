"""ctypes front end of oracle/pursuit_oracle.c (liboracle_c.so)  --  TEST INFRASTRUCTURE ONLY.

The plain-C restatement of the reference's `mp` / `omp` / `gomp` (Float64), one signal per host thread.  Used by
tests/test_oracle_c.py (cross-check against the NumPy oracle) and by bench.py's CPU legs (`cpu_baseline`,
`--impl reference`).  Nothing under compressedsensing.jl_b200/ may import it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_int, c_int64
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle_c.so")
_lib = None


def build() -> str:
    """`make -C oracle` (gcc); returns the library path."""
    r = subprocess.run(["make", "-C", HERE], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("building oracle/liboracle_c.so failed:\n" + r.stdout[-2000:])
    return LIB_PATH


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        L = ctypes.CDLL(LIB_PATH)
        i64p, f64p = POINTER(c_int64), POINTER(c_double)
        L.cs_oracle_solve_batch.restype = c_int
        L.cs_oracle_solve_batch.argtypes = [c_int, f64p, c_int64, c_int64, c_int64, f64p, c_int64, c_int64, c_int64,
                                            c_int64, c_double, c_int64, i64p, i64p, f64p, i64p, f64p, i64p, c_int,
                                            POINTER(c_int)]
        _lib = L
    return _lib


ALGO = {"omp": 0, "gomp": 1, "mp": 2}


def solve_batch(algo: str, A: np.ndarray, B: np.ndarray, k: int, l: int = 1, eps: Optional[float] = None,
                threads: int = 0):
    """Run `algo` for every column of B.  Returns a dict of arrays with one row per signal:
    order (atoms in the order they were appended / selected, -1 padded), nzind / nzval (the SparseVector, first nnz
    entries), nnz, resnorm, iters, and threads (the team size used)."""
    A = np.asfortranarray(A, dtype=np.float64)
    B = np.asfortranarray(np.asarray(B, dtype=np.float64).reshape(A.shape[0], -1))
    if not (np.all(np.isfinite(A)) and np.all(np.isfinite(B))):
        raise ValueError("non-finite input")
    M, N = A.shape
    nsig = B.shape[1]
    eps = float(np.finfo(np.float64).eps) if eps is None else float(eps)
    if not eps >= 0:
        raise ValueError(f"ε = {eps} has to be non-negative")
    cap = max(1, int(k))
    order = np.full((nsig, cap), -1, dtype=np.int64)
    nzind = np.full((nsig, cap), -1, dtype=np.int64)
    nzval = np.zeros((nsig, cap), dtype=np.float64)
    nnz = np.zeros(nsig, dtype=np.int64)
    res = np.zeros(nsig, dtype=np.float64)
    its = np.zeros(nsig, dtype=np.int64)
    used = c_int(0)
    i64p, f64p = POINTER(c_int64), POINTER(c_double)
    rc = lib().cs_oracle_solve_batch(ALGO[algo], A.ctypes.data_as(f64p), M, N, M, B.ctypes.data_as(f64p), M, nsig,
                                     int(k), int(l), eps, cap, order.ctypes.data_as(i64p), nzind.ctypes.data_as(i64p),
                                     nzval.ctypes.data_as(f64p), nnz.ctypes.data_as(i64p), res.ctypes.data_as(f64p),
                                     its.ctypes.data_as(i64p), int(threads), ctypes.byref(used))
    if rc != 0:
        raise RuntimeError(f"cs_oracle_solve_batch failed: {rc}")
    return {"order": order, "nzind": nzind, "nzval": nzval, "nnz": nnz, "resnorm": res, "iters": its,
            "threads": used.value}
