"""Host-side mirror of CompressedSensing.jl's greedy-pursuit call surface over libcsb200.so.

Julia is not installed in this image, so the host layer that the north star asks for in Julia
(`julia/CompressedSensingB200.jl`, written against the same C ABI) is
mirrored here 1:1 in Python/ctypes so that it can be exercised end to end on the GPU box.
Same names, argument meaning and error behaviour as the reference:

    omp(A, b, k)            /root/reference/src/matchingpursuit.jl:84-86
    omp(A, b, eps, k=M)     :73-82        (eps < 0 raises, as the reference throws)
    omp(A, b; max_residual, sparsity)     :88-91
    gomp(A, b, l, k) / gomp(A, b, l, eps, k=M) / gomp(A, b, l; max_residual, sparsity=N)   :126-148
    mp(A, b, k, x=spzeros(N))             :34-40

Every call returns a `SparseVector` (length N, `nzind` strictly ascending, Float64 `nzval`
even for a Float32 dictionary -- `spzeros(N)` at :76).  Indices are 0-based on the Python side;
the Julia shim adds 1.  Additive: `b` may be an M x B matrix (one signal per column), in which
case a list of B SparseVectors is returned -- the column-wise map of the single-signal call.

There is no CPU fallback: all arithmetic happens in the CUDA library; if it is missing or no
sm_100 device is present the call raises.  Nothing here imports `oracle/`.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int64, c_void_p
from dataclasses import dataclass
from typing import List, Optional, Sequence, Union

import numpy as np

__all__ = [
    "SparseVector", "Dictionary", "Batch", "omp", "gomp", "mp", "lib", "LIB_PATH", "CSB200Error",
    "device_count", "F64", "F32", "ShardComm", "omp_sharded", "shard_range", "owner_of", "pick_global",
    "exchange_unique_id", "assemble_csc", "fr", "ols", "oomp", "ormp", "sp", "oblivious", "colnorms", "normalize",
    "cumbabel", "babel", "coherence",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcsb200.so")

F64, F32 = 0, 1
NCCL_ID_BYTES = 128

_STATUS_NAMES = {
    0: "OK", -1: "INVALID_ARG", -2: "NEGATIVE_EPS", -3: "NONFINITE_INPUT", -4: "CUDA", -5: "OOM",
    -6: "UNSUPPORTED_ARCH", -7: "UNSUPPORTED", -8: "NCCL", -9: "DIM_MISMATCH",
}


class CSB200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C compressedsensing.jl_b200/csrc`.  There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    i64p, f64p = POINTER(c_int64), POINTER(c_double)
    sig = {
        "csb200_version": (c_int, []),
        "csb200_strerror": (c_char_p, [c_int]),
        "csb200_last_error": (c_char_p, []),
        "csb200_device_count": (c_int, []),
        "csb200_dict_create": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, POINTER(c_void_p)]),
        "csb200_dict_create_shard": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int64, c_int64,
                                             POINTER(c_void_p)]),
        "csb200_dict_create_multi": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, POINTER(c_int), c_int,
                                             POINTER(c_void_p)]),
        "csb200_dict_devices": (c_int, [c_void_p, POINTER(c_int), c_int]),
        "csb200_dict_destroy": (c_int, [c_void_p]),
        "csb200_dict_trim": (c_int, [c_void_p]),
        "csb200_dict_shape": (c_int, [c_void_p, i64p, i64p, POINTER(c_int), POINTER(c_int)]),
        "csb200_batch_create": (c_int, [c_void_p, c_int64, c_int64, POINTER(c_void_p)]),
        "csb200_batch_destroy": (c_int, [c_void_p]),
        "csb200_batch_upload": (c_int, [c_void_p, c_void_p, c_int64, c_int64]),
        "csb200_batch_upload_device": (c_int, [c_void_p, c_void_p, c_int64, c_int64]),
        "csb200_batch_omp": (c_int, [c_void_p, c_int64, c_double]),
        "csb200_batch_gomp": (c_int, [c_void_p, c_int64, c_int64, c_double]),
        "csb200_batch_mp": (c_int, [c_void_p, c_int64, i64p, f64p, i64p, c_int64]),
        "csb200_batch_fr": (c_int, [c_void_p, c_int64, c_double, c_double]),
        "csb200_batch_sp": (c_int, [c_void_p, c_int64, c_double, c_int64]),
        "csb200_batch_oblivious": (c_int, [c_void_p, c_int64]),
        "csb200_sp": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_double, c_int64, i64p, f64p, i64p, f64p,
                              i64p]),
        "csb200_oblivious": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, i64p, f64p, i64p, f64p]),
        "csb200_fr": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_double, c_double, i64p, f64p, i64p, f64p,
                              i64p]),
        "csb200_batch_download": (c_int, [c_void_p, c_int64, i64p, f64p, i64p, f64p, i64p]),
        "csb200_batch_flags": (c_int, [c_void_p, POINTER(ctypes.c_int32)]),
        "csb200_batch_profile": (c_int, [c_void_p, c_int]),
        "csb200_batch_corr_time": (c_int, [c_void_p, f64p, i64p, i64p]),
        "csb200_batch_last_solve_ms": (c_int, [c_void_p, f64p]),
        "csb200_omp": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_double, i64p, f64p, i64p, f64p, i64p]),
        "csb200_gomp": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_double, i64p, f64p, i64p,
                                f64p, i64p]),
        "csb200_mp": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, i64p, f64p, i64p, c_int64, i64p, f64p,
                              f64p]),
        "csb200_dict_colnorms": (c_int, [c_void_p, f64p]),
        "csb200_dict_cumbabel": (c_int, [c_void_p, c_int64, f64p]),
        "csb200_assemble_csc": (c_int, [c_int64, c_int64, i64p, f64p, i64p, c_int64, i64p, i64p, f64p]),
        "csb200_comm_unique_id": (c_int, [c_void_p]),
        "csb200_comm_create": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_void_p)]),
        "csb200_comm_destroy": (c_int, [c_void_p]),
        "csb200_comm_exchange_mode": (c_int, [c_void_p]),
        "csb200_comm_last_timing": (c_int, [c_void_p, f64p, i64p]),
        "csb200_omp_sharded": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_double, i64p, f64p, i64p, f64p, i64p,
                                       f64p]),
        "csb200_debug_corr_topk": (c_int, [c_void_p, c_int, c_int64, i64p, f64p]),
        "csb200_debug_get_residual": (c_int, [c_void_p, c_void_p]),
        "csb200_debug_screen_pass": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_int64), POINTER(c_double)]),
        "csb200_batch_screen_stats": (c_int, [c_void_p, POINTER(c_int64), c_void_p, c_int]),
        "csb200_debug_graph_replays": (c_int, [c_void_p, POINTER(c_int64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)          # AttributeError here = the library does not export what csb200.h declares
        fn.restype, fn.argtypes = res, args
    return L


lib = _load()
EXPORTED_SYMBOLS = [
    "csb200_version", "csb200_strerror", "csb200_last_error", "csb200_device_count", "csb200_dict_create",
    "csb200_dict_create_shard", "csb200_dict_create_multi", "csb200_dict_devices", "csb200_dict_destroy", "csb200_dict_trim", "csb200_dict_shape", "csb200_batch_create",
    "csb200_batch_destroy", "csb200_batch_upload", "csb200_batch_upload_device", "csb200_batch_omp",
    "csb200_batch_gomp", "csb200_batch_mp", "csb200_batch_download", "csb200_batch_flags", "csb200_batch_profile",
    "csb200_batch_corr_time", "csb200_batch_last_solve_ms", "csb200_omp", "csb200_gomp", "csb200_mp",
    "csb200_assemble_csc", "csb200_comm_unique_id", "csb200_batch_fr", "csb200_fr",
    "csb200_batch_sp", "csb200_batch_oblivious", "csb200_sp", "csb200_oblivious", "csb200_dict_colnorms",
    "csb200_dict_cumbabel",
    "csb200_comm_create", "csb200_comm_destroy", "csb200_comm_exchange_mode", "csb200_comm_last_timing", "csb200_omp_sharded", "csb200_debug_corr_topk",
    "csb200_debug_get_residual", "csb200_debug_graph_replays", "csb200_debug_screen_pass", "csb200_batch_screen_stats",
]


def _check(rc: int, eps: Optional[float] = None) -> None:
    if rc == 0:
        return
    if rc == -2:
        # the reference throws the String "ε = $ε has to be non-negative" (matchingpursuit.jl:74,127)
        raise ValueError(f"ε = {eps} has to be non-negative")
    detail = lib.csb200_last_error().decode() if rc in (-4, -5, -6, -7, -8, -1) else ""
    msg = f"csb200: {lib.csb200_strerror(rc).decode()} [{_STATUS_NAMES.get(rc, rc)}]" + (f": {detail}" if detail else "")
    if rc == -9:
        raise ValueError(msg)          # Julia: DimensionMismatch
    raise CSB200Error(rc, msg)


def device_count() -> int:
    n = lib.csb200_device_count()
    if n < 0:
        _check(n)
    return n


def _i64p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(POINTER(c_int64))


def _f64p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(POINTER(c_double))


# ------------------------------------------------------------------------------------------------
@dataclass
class SparseVector:
    """Stand-in for Julia's `SparseVector{Float64,Int64}` (0-based `nzind`, ascending)."""
    n: int
    nzind: np.ndarray
    nzval: np.ndarray

    def nnz(self) -> int:
        return int(self.nzind.size)

    def dense(self) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.float64)
        out[self.nzind] = self.nzval
        return out

    def __repr__(self) -> str:
        return f"SparseVector(n={self.n}, nzind={self.nzind.tolist()}, nzval={self.nzval.tolist()})"


def _as_matrix(A) -> np.ndarray:
    A = np.asarray(A)
    if A.ndim != 2:
        raise ValueError("A must be a matrix")
    if A.dtype not in (np.float64, np.float32):
        A = A.astype(np.float64)                     # the shim copies anything that is not a strided Float32/64 matrix
    return np.asfortranarray(A)


class Dictionary:
    """A dictionary resident in HBM (`csb200_dict`).  Reuse it across solves to keep A in HBM/L2.

    `devices=[...]` replicates it onto several GPUs (`csb200_dict_create_multi`): the one-shot calls (`omp(D, B, k)`
    etc.) then split the columns of B over those GPUs inside the library, with results bit-identical to one GPU."""

    def __init__(self, A, device: int = 0, n_offset: int = 0, n_total: Optional[int] = None,
                 devices: Optional[Sequence[int]] = None):
        A = _as_matrix(A)
        self.M, self.N = int(A.shape[0]), int(A.shape[1])
        self.dtype = A.dtype
        self.n_offset = int(n_offset)
        self.n_total = int(self.N if n_total is None else n_total)
        h = c_void_p()
        lda = A.strides[1] // A.itemsize if self.N > 1 else self.M
        dt = F32 if A.dtype == np.float32 else F64
        if devices is not None:
            devs = [int(d) for d in devices]
            if not devs or n_offset or self.n_total != self.N:
                raise ValueError("devices= needs at least one device and an unsharded dictionary")
            arr = (c_int * len(devs))(*devs)
            _check(lib.csb200_dict_create_multi(A.ctypes.data, self.M, self.N, max(lda, self.M), dt, arr, len(devs),
                                                byref(h)))
            self.device, self.devices = devs[0], devs
        else:
            _check(lib.csb200_dict_create_shard(A.ctypes.data, self.M, self.N, max(lda, self.M), dt, device,
                                                self.n_offset, self.n_total, byref(h)))
            self.device, self.devices = device, [device]
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.csb200_dict_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class Batch:
    """Device-resident solver state for up to `max_signals` right-hand sides (`csb200_batch`)."""

    def __init__(self, dictionary: Dictionary, max_signals: int, max_sparsity: int):
        self.dict = dictionary
        self.max_signals, self.max_sparsity = int(max_signals), int(max_sparsity)
        h = c_void_p()
        _check(lib.csb200_batch_create(dictionary._h, self.max_signals, self.max_sparsity, byref(h)))
        self._h = h
        self.nsig = 0

    def upload(self, B: np.ndarray) -> None:
        B = self._signals(B)
        self.nsig = B.shape[1]
        _check(lib.csb200_batch_upload(self._h, B.ctypes.data, max(B.strides[1] // B.itemsize, B.shape[0]), self.nsig))

    def upload_device(self, ptr: int, ldb: int, nsig: int) -> None:
        self.nsig = int(nsig)
        _check(lib.csb200_batch_upload_device(self._h, c_void_p(ptr), ldb, nsig))

    def _signals(self, B) -> np.ndarray:
        B = np.asarray(B)
        if B.ndim == 1:
            B = B.reshape(-1, 1)
        if B.shape[0] != self.dict.M:
            raise ValueError(f"DimensionMismatch: signal length {B.shape[0]} != {self.dict.M} rows of A")
        return np.asfortranarray(B.astype(self.dict.dtype, copy=False))

    def omp(self, k: int, eps: float) -> None:
        _check(lib.csb200_batch_omp(self._h, int(k), float(eps)), eps)

    def gomp(self, l: int, k: int, eps: float) -> None:
        _check(lib.csb200_batch_gomp(self._h, int(l), int(k), float(eps)), eps)

    def fr(self, k: int, max_eps: float = 0.0, min_delta: float = 0.0) -> None:
        _check(lib.csb200_batch_fr(self._h, int(k), float(max_eps), float(min_delta)))

    def sp(self, k: int, delta: float = 1e-12, maxiter: Optional[int] = None) -> None:
        _check(lib.csb200_batch_sp(self._h, int(k), float(delta), int(16 * k if maxiter is None else maxiter)))

    def oblivious(self, k: int) -> None:
        _check(lib.csb200_batch_oblivious(self._h, int(k)))

    def mp(self, iters: int, x0: Optional[Sequence[SparseVector]] = None) -> None:
        if x0 is None:
            _check(lib.csb200_batch_mp(self._h, int(iters), None, None, None, 0))
            return
        stride = max(1, max(v.nnz() for v in x0))
        idx = np.zeros((self.nsig, stride), dtype=np.int64)
        val = np.zeros((self.nsig, stride), dtype=np.float64)
        nnz = np.zeros(self.nsig, dtype=np.int64)
        for s, v in enumerate(x0):
            nnz[s] = v.nnz()
            idx[s, :v.nnz()] = v.nzind
            val[s, :v.nnz()] = v.nzval
        _check(lib.csb200_batch_mp(self._h, int(iters), _i64p(idx), _f64p(val), _i64p(nnz), stride))

    def download(self, stride: int, want_coef: bool = True):
        ns, stride = self.nsig, max(int(stride), 1)
        sel = np.empty((ns, stride), dtype=np.int64)
        coef = np.empty((ns, stride), dtype=np.float64) if want_coef else None
        nnz = np.empty(ns, dtype=np.int64)
        res = np.empty(ns, dtype=np.float64)
        its = np.empty(ns, dtype=np.int64)
        _check(lib.csb200_batch_download(self._h, stride, _i64p(sel), _f64p(coef), _i64p(nnz), _f64p(res), _i64p(its)))
        return sel, coef, nnz, res, its

    def flags(self) -> np.ndarray:
        """Per-signal diagnostic bits of the last solve (`csb200_batch_flags`): 1 dependent atom skipped, 2 no
        candidate, 16 ill-conditioned support (coefficients refined)."""
        out = np.zeros(self.nsig, dtype=np.int32)
        _check(lib.csb200_batch_flags(self._h, out.ctypes.data_as(POINTER(ctypes.c_int32))))
        return out

    def profile(self, enable: bool) -> None:
        _check(lib.csb200_batch_profile(self._h, 1 if enable else 0))

    def corr_time(self):
        ms, n, other = c_double(), c_int64(), c_int64()
        _check(lib.csb200_batch_corr_time(self._h, byref(ms), byref(n), byref(other)))
        return ms.value, n.value, other.value

    def last_solve_ms(self) -> float:
        ms = c_double()
        _check(lib.csb200_batch_last_solve_ms(self._h, byref(ms)))
        return ms.value

    def debug_corr_topk(self, s: int, impl: int = 0):
        idx = np.empty((self.nsig, s), dtype=np.int64)
        val = np.empty((self.nsig, s), dtype=np.float64)
        _check(lib.csb200_debug_corr_topk(self._h, impl, s, _i64p(idx), _f64p(val)))
        return idx, val

    def debug_screen_pass(self):
        """One TF32 screening pass over the current residuals: (|c~| [nsig, chunks*8], atom [nsig, chunks*8], bound)."""
        val = np.empty((self.nsig, 128), dtype=np.float32)
        idx = np.empty((self.nsig, 128), dtype=np.int32)
        chunks = c_int64()
        bound = c_double()
        _check(lib.csb200_debug_screen_pass(self._h, val.ctypes.data, idx.ctypes.data, byref(chunks), byref(bound)))
        nc = int(chunks.value) * 8
        return (val.reshape(-1)[: self.nsig * nc].reshape(self.nsig, nc).copy(),
                idx.reshape(-1)[: self.nsig * nc].reshape(self.nsig, nc).copy(), bound.value)

    def screen_stats(self, reset: bool = False) -> dict:
        """Path of the last omp solve on this batch and the screening counters (`csb200_batch_screen_stats`)."""
        path = c_int64()
        st = np.zeros(3, dtype=np.uint64)
        _check(lib.csb200_batch_screen_stats(self._h, byref(path), st.ctypes.data, 1 if reset else 0))
        names = {0: "few-signal / gemv", 1: "dmma", 2: "dmma two-half overlap", 3: "tf32 screening + exact re-evaluation",
                 4: "fp16 screening + exact re-evaluation"}
        return {"path": names.get(int(path.value), str(path.value)), "path_id": int(path.value), "signal_updates": int(st[0]),
                "candidates_reevaluated": int(st[1]), "exact_scans": int(st[2])}

    def residual(self) -> np.ndarray:
        out = np.empty((self.dict.M, self.nsig), dtype=self.dict.dtype, order="F")
        _check(lib.csb200_debug_get_residual(self._h, out.ctypes.data))
        return out

    def graph_replays(self) -> int:
        """Solves of this batch that ran as a CUDA-graph launch (few-signal paths; debug hook)."""
        n = c_int64()
        _check(lib.csb200_debug_graph_replays(self._h, byref(n)))
        return n.value

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.csb200_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ------------------------------------------------------------------------------------------------
def _eps_of(dtype) -> float:
    return float(np.finfo(np.dtype(dtype)).eps)


def _dictionary(A, device):
    if isinstance(A, Dictionary):
        return A, False
    if isinstance(device, (list, tuple, range)):          # omp(A, B, k, device=range(8)): replicate for this call
        return Dictionary(A, devices=list(device)), True
    return Dictionary(A, device=device), True


def _to_sparse(n: int, sel_row: np.ndarray, coef_row: np.ndarray, nnz: int) -> SparseVector:
    """Sort (index, coefficient) pairs ascending: the order `x[i] = NaN` keeps (`src/util.jl:120-122`)."""
    idx, val = sel_row[:nnz], coef_row[:nnz]
    order = np.argsort(idx, kind="stable")
    return SparseVector(n, idx[order].astype(np.int64), val[order].astype(np.float64))


def _signals_for(D: "Dictionary", b) -> np.ndarray:
    B = np.asarray(b)
    if B.ndim == 1:
        B = B.reshape(-1, 1)
    if B.ndim != 2 or B.shape[0] != D.M:
        raise ValueError(f"DimensionMismatch: signal length {B.shape[0]} != {D.M} rows of A")
    return np.asfortranarray(B.astype(D.dtype, copy=False))


def assemble_csc(n: int, sel: np.ndarray, coef: np.ndarray, nnz: np.ndarray):
    """(colptr, rowval, nzval) of the n x nsig coefficient matrix, 0-based, rows ascending within each column --
    the batched result format (what the Julia shim wraps in a SparseMatrixCSC)."""
    nsig, stride = sel.shape
    colptr = np.empty(nsig + 1, dtype=np.int64)
    total = int(nnz.sum())
    rowval = np.empty(max(total, 1), dtype=np.int64)
    nzval = np.empty(max(total, 1), dtype=np.float64)
    sel = np.ascontiguousarray(sel, dtype=np.int64); coef = np.ascontiguousarray(coef, dtype=np.float64)
    nnz = np.ascontiguousarray(nnz, dtype=np.int64)
    _check(lib.csb200_assemble_csc(nsig, stride, _i64p(sel), _f64p(coef), _i64p(nnz), 0, _i64p(colptr), _i64p(rowval),
                                   _f64p(nzval)))
    return colptr, rowval[:total], nzval[:total]


def _solve(A, b, device, call, stride, eps=None, merge=None, result="vectors"):
    """One-shot solve through the host-buffer C entry points (csb200_omp / csb200_gomp / csb200_mp): upload,
    solve and download happen inside one C call on a workspace cached on the dictionary handle.
    result = "vectors": SparseVector (or a list of them); "csc": scipy.sparse.csc_matrix (N x nsig), assembled in C."""
    single = np.asarray(b).ndim == 1
    D, owned = _dictionary(A, device)
    try:
        B = _signals_for(D, b)
        nsig = B.shape[1]
        ldb = max(B.strides[1] // B.itemsize, D.M) if nsig > 1 else D.M
        sel = np.empty((nsig, stride), dtype=np.int64)
        coef = np.empty((nsig, stride), dtype=np.float64)
        nnz = np.full(nsig, stride, dtype=np.int64)
        res = np.empty(nsig, dtype=np.float64)
        its = np.empty(nsig, dtype=np.int64)
        _check(call(D, B, ldb, nsig, sel, coef, nnz, res, its), eps)
        if result == "csc" and merge is None:
            import scipy.sparse as sp
            colptr, rowval, nzval = assemble_csc(D.n_total, sel, coef, nnz)
            return sp.csc_matrix((nzval, rowval, colptr), shape=(D.n_total, nsig))
        if merge is not None:
            out = [merge(D.n_total, sel[s], coef[s], int(nnz[s]), s) for s in range(nsig)]
        else:
            out = [_to_sparse(D.n_total, sel[s], coef[s], int(nnz[s])) for s in range(nsig)]
    finally:
        if owned:
            D.close()
    return out[0] if single else out


def _is_int(v) -> bool:
    return isinstance(v, (int, np.integer)) and not isinstance(v, bool)


def omp(A, b, *args, max_residual=None, sparsity=None, device: int = 0, result: str = "vectors"):
    """Orthogonal matching pursuit -- `omp` (`src/matchingpursuit.jl:73-91`).

    omp(A, b, k) | omp(A, b, eps[, k]) | omp(A, b, max_residual=..., sparsity=...)
    """
    M, N = (A.M, A.n_total) if isinstance(A, Dictionary) else np.shape(A)
    dt = A.dtype if isinstance(A, Dictionary) else (np.float32 if np.asarray(A).dtype == np.float32 else np.float64)
    if len(args) == 1 and _is_int(args[0]):
        eps, k = _eps_of(dt), int(args[0])                           # :84-86
    elif len(args) >= 1:
        eps, k = float(args[0]), (int(args[1]) if len(args) > 1 else M)   # :73
    else:
        eps = _eps_of(dt) if max_residual is None else float(max_residual)   # :88-91
        k = min(M, N) if sparsity is None else int(sparsity)
    if not eps >= 0:
        raise ValueError(f"ε = {eps} has to be non-negative")
    if k < 0:
        raise ValueError("k must be non-negative")
    stride = max(k, 1)

    def call(D, B, ldb, nsig, sel, coef, nnz, res, its):
        return lib.csb200_omp(D._h, B.ctypes.data, ldb, nsig, k, eps, _i64p(sel), _f64p(coef), _i64p(nnz), _f64p(res),
                              _i64p(its))

    return _solve(A, b, device, call, stride, eps, result=result)


def gomp(A, b, l: int, *args, max_residual=None, sparsity=None, device: int = 0, result: str = "vectors"):
    """Generalized OMP, l atoms per update -- `gomp` (`src/matchingpursuit.jl:126-148`)."""
    M, N = (A.M, A.n_total) if isinstance(A, Dictionary) else np.shape(A)
    dt = A.dtype if isinstance(A, Dictionary) else (np.float32 if np.asarray(A).dtype == np.float32 else np.float64)
    if len(args) == 1 and _is_int(args[0]):
        eps, k = _eps_of(dt), int(args[0])                           # :141-143
    elif len(args) >= 1:
        eps, k = float(args[0]), (int(args[1]) if len(args) > 1 else M)   # :126
    else:
        eps = _eps_of(dt) if max_residual is None else float(max_residual)   # :145-148
        k = N if sparsity is None else int(sparsity)
    if not eps >= 0:
        raise ValueError(f"ε = {eps} has to be non-negative")
    if k < 0:
        raise ValueError("k must be non-negative")
    stride = max(k, 1)

    def call(D, B, ldb, nsig, sel, coef, nnz, res, its):
        return lib.csb200_gomp(D._h, B.ctypes.data, ldb, nsig, int(l), k, eps, _i64p(sel), _f64p(coef), _i64p(nnz),
                               _f64p(res), _i64p(its))

    return _solve(A, b, device, call, stride, eps, result=result)


def fr(A, b, *args, max_residual: float = 0.0, min_decrease: float = 0.0, sparsity=None, device: int = 0,
       result: str = "vectors"):
    """Forward regression a.k.a. OLS / OOMP / ORMP -- `fr` (`src/forward.jl:33-54`).

    fr(A, b, max_eps, min_delta[, k = M])  |  fr(A, b, max_residual=0, min_decrease=0, sparsity=N)
    """
    M, N = (A.M, A.n_total) if isinstance(A, Dictionary) else np.shape(A)
    if len(args) >= 2:
        max_eps, min_delta = float(args[0]), float(args[1])              # :44
        k = int(args[2]) if len(args) > 2 else M
    elif len(args) == 0:
        max_eps, min_delta = float(max_residual), float(min_decrease)    # :33-36
        k = N if sparsity is None else int(sparsity)
    else:
        raise TypeError("fr(A, b, max_eps, min_delta[, k]) or fr(A, b, max_residual=, min_decrease=, sparsity=)")
    if k < 0:
        raise ValueError("k must be non-negative")
    stride = max(min(k, M, N), 1)

    def call(D, B, ldb, nsig, sel, coef, nnz, res, its):
        return lib.csb200_fr(D._h, B.ctypes.data, ldb, nsig, min(k, M, N), max_eps, min_delta, _i64p(sel), _f64p(coef),
                             _i64p(nnz), _f64p(res), _i64p(its))

    return _solve(A, b, device, call, stride, result=result)


ols = oomp = ormp = fr            # `const ols = fr` etc. (src/forward.jl:52-54)


def sp(A, b, k: int, delta: float = 1e-12, maxiter: Optional[int] = None, device: int = 0, result: str = "vectors"):
    """Subspace pursuit -- `sp(A, b, k, δ = 1e-12; maxiter = 16k)` (`src/twostage.jl:105-117`)."""
    M = A.M if isinstance(A, Dictionary) else np.shape(A)[0]
    k = int(k)
    if 2 * k > M:
        raise ValueError(f"2k = {2 * k} > {M} = length(b) is invalid for Subspace Pursuit")     # twostage.jl:62
    mi = 16 * k if maxiter is None else int(maxiter)

    def call(D, B, ldb, nsig, sel, coef, nnz, res, its):
        return lib.csb200_sp(D._h, B.ctypes.data, ldb, nsig, k, float(delta), mi, _i64p(sel), _f64p(coef), _i64p(nnz),
                             _f64p(res), _i64p(its))

    return _solve(A, b, device, call, max(k, 1), result=result)


def oblivious(A, b, k: int, device: int = 0, result: str = "vectors"):
    """Oblivious selection -- `oblivious(A, b, k)` (`src/oblivious.jl:3-8`): the k atoms most correlated with b.
    (The reference allocates its result with `spzeros(size(b))`, i.e. length M; this returns length N.)"""
    k = int(k)

    def call(D, B, ldb, nsig, sel, coef, nnz, res, its):
        its[:] = 1
        return lib.csb200_oblivious(D._h, B.ctypes.data, ldb, nsig, k, _i64p(sel), _f64p(coef), _i64p(nnz), _f64p(res))

    return _solve(A, b, device, call, max(k, 1), result=result)


def mp(A, b, k: int, x=None, device: int = 0):
    """Matching pursuit, exactly k updates, optional warm start x -- `mp` (`src/matchingpursuit.jl:34-40`)."""
    b_arr = np.asarray(b)
    single = b_arr.ndim == 1
    x0 = None
    if x is not None:
        x0 = [x] if isinstance(x, SparseVector) else list(x)

    def merge(n, sel_row, coef_row, nnz, s):
        # x[i] += <a_i, r> in iteration order (`:29`), starting from the warm start
        acc = {}
        if x0 is not None:
            for i, v in zip(x0[s].nzind.tolist(), x0[s].nzval.tolist()):
                acc[i] = v
        for i, c in zip(sel_row[:nnz].tolist(), coef_row[:nnz].tolist()):
            if i >= 0:
                acc[i] = acc.get(i, 0.0) + c
        idx = np.array(sorted(acc), dtype=np.int64)
        return SparseVector(n, idx, np.array([acc[i] for i in idx.tolist()], dtype=np.float64))

    k = int(k)
    stride = max(k, 1)

    def call(D, B, ldb, nsig, sel, coef, nnz, res, its):
        if x0 is None:
            rc = lib.csb200_mp(D._h, B.ctypes.data, ldb, nsig, k, None, None, None, 0, _i64p(sel), _f64p(coef), _f64p(res))
        else:
            if len(x0) != nsig:
                raise ValueError("one warm-start vector per signal is required")
            st0 = max(1, max(v.nnz() for v in x0))
            xi = np.zeros((nsig, st0), dtype=np.int64)
            xv = np.zeros((nsig, st0), dtype=np.float64)
            xn = np.zeros(nsig, dtype=np.int64)
            for s_, v in enumerate(x0):
                xn[s_] = v.nnz(); xi[s_, :v.nnz()] = v.nzind; xv[s_, :v.nnz()] = v.nzval
            rc = lib.csb200_mp(D._h, B.ctypes.data, ldb, nsig, k, _i64p(xi), _f64p(xv), _i64p(xn), st0, _i64p(sel),
                               _f64p(coef), _f64p(res))
        nnz[:] = k                     # mp: one (atom, increment) record per iteration
        return rc

    return _solve(A, b, device, call, stride, merge=merge)


# ------------------------------------------------------------------------------------------------
# Dictionary analysis (`src/util.jl:2, 59-61, 96-117`)
def colnorms(A, device: int = 0) -> np.ndarray:
    """`colnorms(A)` (`src/util.jl:2`): 2-norm of every column."""
    D, owned = _dictionary(A, device)
    try:
        out = np.empty(D.N, dtype=np.float64)
        _check(lib.csb200_dict_colnorms(D._h, _f64p(out)))
    finally:
        if owned:
            D.close()
    return out


def normalize(A: np.ndarray, device: int = 0) -> np.ndarray:
    """`normalize!(A)` (`src/util.jl:59-61`): `A ./= colnorms(A)'`, in place on the caller's (host) matrix."""
    A /= colnorms(A, device).astype(A.dtype)[None, :]
    return A


def cumbabel(A, k: int, device: int = 0) -> np.ndarray:
    """`cumbabel(A, k)` (`src/util.jl:106-117`): Babel function values mu_1(1..k)."""
    D, owned = _dictionary(A, device)
    try:
        mu = np.empty(int(k), dtype=np.float64)
        _check(lib.csb200_dict_cumbabel(D._h, int(k), _f64p(mu)))
    finally:
        if owned:
            D.close()
    return mu.astype(D.dtype)             # `zeros(eltype(A), k)` (:107)


def babel(A, k: int, device: int = 0):
    """`babel(A, k) = cumbabel(A, k)[k]` (`src/util.jl:101`)."""
    return cumbabel(A, k, device)[k - 1]


def coherence(A, device: int = 0):
    """Mutual coherence, `coherence(A) = babel(A, 1)` (`src/util.jl:98`)."""
    return babel(A, 1, device)


# ------------------------------------------------------------------------------------------------
# Multi-GPU host logic (one process per GPU).  Independent signals are split with no communication;
# a single huge dictionary is column-sharded.
def shard_range(n: int, nranks: int, rank: int):
    """Contiguous [lo, hi) slice of n units owned by `rank`: sizes differ by at most one, lower ranks first."""
    if not (0 <= rank < nranks):
        raise ValueError("rank out of range")
    base, extra = divmod(n, nranks)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_of(index: int, n: int, nranks: int) -> int:
    """Rank whose shard_range contains `index`."""
    base, extra = divmod(n, nranks)
    cut = extra * (base + 1)
    return index // (base + 1) if index < cut else extra + (index - cut) // max(base, 1)


def pick_global(records):
    """The per-iteration exchange of the column-sharded solve, on the host: given every rank's best
    (|c|, global atom index) choose the winner exactly as global_pick_kernel does -- largest |c|, lowest
    index on ties (Julia `argmax`).  Returns (rank, value, index); index -1 if no rank has a candidate."""
    best = (-1, -1.0, -1)
    for g, (v, i) in enumerate(records):
        if i < 0:
            continue
        if best[2] < 0 or v > best[1] or (v == best[1] and i < best[2]):
            best = (g, float(v), int(i))
    return best


def exchange_unique_id(dist, rank: int, make_id=None) -> bytes:
    """Rank 0 creates the 128-byte NCCL unique id, everyone receives it through torch.distributed
    (any backend; the tests use gloo)."""
    import torch
    make_id = make_id or ShardComm.unique_id
    buf = torch.zeros(NCCL_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        buf = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone()
    if buf.device.type == "cpu" and dist.get_backend() == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        buf = buf.to(dev)
        dist.broadcast(buf, src=0)
        buf = buf.cpu()
    else:
        dist.broadcast(buf, src=0)
    return bytes(buf.numpy().tobytes())


class ShardComm:
    """NCCL communicator owned by the library (`csb200_comm`), one per process/GPU."""

    @staticmethod
    def unique_id() -> bytes:
        try:
            import torch  # noqa: F401  (see __init__)
        except ImportError:
            pass
        buf = ctypes.create_string_buffer(NCCL_ID_BYTES)
        _check(lib.csb200_comm_unique_id(buf))
        return buf.raw

    def __init__(self, unique_id: bytes, rank: int, nranks: int, device: int):
        try:                      # let PyTorch load the NCCL it was built against first: the library then reuses that one
            import torch  # noqa: F401
        except ImportError:
            pass
        h = c_void_p()
        buf = ctypes.create_string_buffer(unique_id, NCCL_ID_BYTES)
        _check(lib.csb200_comm_create(buf, rank, nranks, device, byref(h)))
        self._h, self.rank, self.nranks = h, rank, nranks

    def last_timing(self) -> dict:
        """Device-time breakdown (ms) of the last sharded solve on this rank (`csb200_comm_last_timing`)."""
        ms = np.zeros(5, dtype=np.float64)
        it = c_int64()
        _check(lib.csb200_comm_last_timing(self._h, _f64p(ms), byref(it)))
        return {"solve_ms": float(ms[0]), "corr_ms": float(ms[1]), "exchange_ms": float(ms[2]), "update_ms": float(ms[3]),
                "gap_ms": float(ms[4]), "iters": int(it.value)}

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.csb200_comm_destroy(self._h)
            self._h = None


def omp_sharded(shard: Dictionary, comm: ShardComm, b, k: int, eps: Optional[float] = None):
    """Single-signal `omp` on a column-sharded dictionary; every rank returns the same SparseVector."""
    eps = _eps_of(shard.dtype) if eps is None else float(eps)
    b = np.ascontiguousarray(np.asarray(b, dtype=shard.dtype))
    if b.shape != (shard.M,):
        raise ValueError(f"DimensionMismatch: signal length {b.shape} != {shard.M} rows of A")
    sel = np.empty(max(k, 1), dtype=np.int64)
    coef = np.empty(max(k, 1), dtype=np.float64)
    nnz, its = np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int64)
    res, ms = np.zeros(1, dtype=np.float64), np.zeros(1, dtype=np.float64)
    _check(lib.csb200_omp_sharded(shard._h, comm._h, b.ctypes.data, int(k), eps, _i64p(sel), _f64p(coef), _i64p(nnz),
                                  _f64p(res), _i64p(its), _f64p(ms)), eps)
    x = _to_sparse(shard.n_total, sel, coef, int(nnz[0]))
    return x, {"resnorm": float(res[0]), "iters": int(its[0]), "corr_ms": float(ms[0]), "order": sel[:int(nnz[0])].copy(),
               "exchange": "peer-memory" if lib.csb200_comm_exchange_mode(comm._h) == 1 else "nccl"}
