"""CPU oracle for the greedy-pursuit hot path of CompressedSensing.jl  --  TEST INFRASTRUCTURE ONLY.

This file is a line-by-line NumPy restatement of the reference algorithm
(`/root/reference/src/matchingpursuit.jl:10-193`, `/root/reference/src/util.jl:118-134`).
It is the *checker* for the CUDA product path and is imported only by `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs.
Nothing under `compressedsensing.jl_b200/` may import it.

PARITY STATUS: **value-level parity is unpinned.**  The reference ships no golden
vectors, no seeds and no fixtures for this path, and Julia is not
installed in this image, so the oracle cannot be compared with outputs of the reference
itself.  What *is* pinned (tests/test_oracle.py): every property the reference's own tests
assert for this path (`test/matchingpursuit.jl:15-45`, `test/forward.jl:24-28`) at the
reference's shapes, plus the Julia stdlib semantics restated below (first-index `argmax`,
`partialsortperm(..., rev=true)` ordering, ascending-index sparse AXPY in `residual!`).

The one piece of arithmetic that lives in an un-vendored dependency is
UpdatableQRFactorizations v1.0.0 (git-tree-sha1 dd1d0589f29fcac6f29bbeff2e3c698a4299c3db,
`Manifest.toml:446-450`): `UpdatableQR(T,n,k)`, `add_column!(F,a,pos)`, `ldiv!(F,r)`.
Its published contract is "QR of the active columns under column insertion, least-squares
solve returned in logical (sorted) order".  Three interchangeable engines restate it here:
  * ``ls="lapack"``  -- dense Householder LS on ``A[:, sort(S)]`` (ground truth),
  * ``ls="givens"``  -- an updatable thin QR with Givens-rotation column insertion
                        (oracle/updatable_qr.py), the scheme the call-site comment at
                        `src/util.jl:121` names, and
  * ``ls="scipy"``   -- a full M x M Q (what the package's call sites reveal it keeps) updated
                        by `scipy.linalg.qr_insert`, an independent published implementation of the same
                        Givens column-insertion primitive.
All three give the same answer to a few ulps; tests assert that (the property
`test/forward.jl:24-28` pins).

Indices are 0-based here; the reference is 1-based.  "first index on ties" is preserved.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from .updatable_qr import UpdatableQR

__all__ = [
    "SparseVec", "Trace", "mp", "omp", "gomp", "residual", "argmaxinner", "argmaxinner_k",
    "sparse_vector", "sparse_data", "perturb", "eps_of", "fr", "ols", "oomp", "ormp", "forward_delta", "ols_rescaling",
    "findmax_first", "sp", "oblivious", "colnorms", "normalize", "cumbabel", "babel", "coherence",
]


def eps_of(dtype) -> float:
    """Julia `eps(T)` (`src/matchingpursuit.jl:85,142`)."""
    return float(np.finfo(np.dtype(dtype)).eps)


# ----------------------------------------------------------------------------------------
# SparseVector{Float64,Int64} stand-in (what `spzeros(N)` returns, `src/matchingpursuit.jl:76`)
# ----------------------------------------------------------------------------------------
@dataclass
class SparseVec:
    """Sorted-index sparse vector; values are ALWAYS float64, as in the reference."""
    n: int
    nzind: List[int] = field(default_factory=list)
    nzval: List[float] = field(default_factory=list)

    def nnz(self) -> int:
        return len(self.nzind)

    def __contains__(self, i: int) -> bool:
        return i in self.nzind

    def setindex(self, i: int, v: float) -> None:
        """`x[i] = v`: sorted insert, or overwrite when already stored (SparseArrays semantics)."""
        lo = int(np.searchsorted(np.asarray(self.nzind, dtype=np.int64), i))
        if lo < len(self.nzind) and self.nzind[lo] == i:
            self.nzval[lo] = float(v)
        else:
            self.nzind.insert(lo, int(i))
            self.nzval.insert(lo, float(v))

    def getindex(self, i: int) -> float:
        lo = int(np.searchsorted(np.asarray(self.nzind, dtype=np.int64), i))
        if lo < len(self.nzind) and self.nzind[lo] == i:
            return self.nzval[lo]
        return 0.0

    def dense(self) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.float64)
        if self.nzind:
            out[np.asarray(self.nzind)] = np.asarray(self.nzval)
        return out

    def copy(self) -> "SparseVec":
        return SparseVec(self.n, list(self.nzind), list(self.nzval))


@dataclass
class Trace:
    """Per-iteration record used by the parity tests (not part of the reference)."""
    selected: List[List[int]] = field(default_factory=list)   # atoms picked by argmax, in order, per update!
    added: List[List[int]] = field(default_factory=list)      # atoms actually appended (not already active)
    margin: List[float] = field(default_factory=list)         # (top1 - top2)/top1 of |c| at each argmax
    resnorm: List[float] = field(default_factory=list)        # ||r||_2 after each update!
    iterations: int = 0

    def order(self) -> List[int]:
        return [j for step in self.added for j in step]


# ----------------------------------------------------------------------------------------
# helpers  (`src/matchingpursuit.jl:152-193`)
# ----------------------------------------------------------------------------------------
def residual(A: np.ndarray, x: SparseVec, b: np.ndarray) -> np.ndarray:
    """`residual!` (`src/matchingpursuit.jl:158-161`): `copyto!(r,b); mul!(r, A, x, -1, 1)`.

    SparseArrays' kernel walks the stored entries in ascending index order and does
    `r[i] += A[i,j] * (x_j * -1)`; with a Float32 `A` the product is formed in Float64
    (x is always Float64) and rounded to Float32 on store.
    """
    T = A.dtype
    r = np.array(b, dtype=T, copy=True)
    for j, v in zip(x.nzind, x.nzval):
        av = np.float64(v) * -1.0
        if T == np.float64:
            r += A[:, j] * av
        else:
            r = (r.astype(np.float64) + A[:, j].astype(np.float64) * av).astype(T)
    return r


def _check_finite(*arrays: np.ndarray) -> None:
    for a in arrays:
        if not np.all(np.isfinite(a)):
            raise ValueError("non-finite input (the replacement rejects NaN/Inf at the boundary; "
                             "see SURVEY.md 8a row a14)")


def correlations(A: np.ndarray, r: np.ndarray) -> np.ndarray:
    """`mul!(P.Ar, P.A', P.r); @. P.Ar = abs(P.Ar)` (`src/matchingpursuit.jl:182-183`) -> BLAS gemv('T')."""
    return np.abs(A.T @ r)


def argmaxinner(A: np.ndarray, r: np.ndarray, trace: Optional[Trace] = None) -> int:
    """`argmaxinner!(P)` (`src/matchingpursuit.jl:181-185`): lowest index among equal maxima."""
    Ar = correlations(A, r)
    i = int(np.argmax(Ar))                       # numpy: first occurrence, same as Julia for finite input
    if trace is not None:
        top = float(Ar[i])
        if Ar.size > 1:
            second = float(np.partition(Ar, -2)[-2])
            trace.margin.append((top - second) / top if top > 0 else 0.0)
        else:
            trace.margin.append(1.0)
    return i


def argmaxinner_k(A: np.ndarray, r: np.ndarray, k: int, trace: Optional[Trace] = None) -> List[int]:
    """`argmaxinner!(P,k)` (`src/matchingpursuit.jl:189-193`): `partialsortperm(Ar, 1:k, rev=true)`.

    Indices of the k largest |c|, descending; `Base.Order.Perm` breaks ties by lower index.
    """
    Ar = correlations(A, r)
    order = np.lexsort((np.arange(Ar.size), -Ar))  # primary: value descending; secondary: index ascending
    sel = [int(j) for j in order[:k]]
    if trace is not None:
        if Ar.size > k:
            kth, nxt = float(Ar[order[k - 1]]), float(Ar[order[k]])
            trace.margin.append((kth - nxt) / kth if kth > 0 else 0.0)
        else:
            trace.margin.append(1.0)
    return sel


class _ActiveSetLS:
    """The `AiQR` field (`src/matchingpursuit.jl:50,58,102,112`) behind `addindex!` / `ldiv!!`."""

    def __init__(self, A: np.ndarray, engine: str, capacity: int):
        self.A = A
        self.engine = engine
        if engine == "givens":
            self.qr = UpdatableQR(A.dtype, A.shape[0], capacity)
        elif engine == "scipy":
            # full (M x M) Q and M x t R updated by SciPy's Givens-based column insertion: an independent,
            # published implementation of the same primitive (the call sites show the package keeps a full Q)
            self.Q = np.eye(A.shape[0], dtype=A.dtype)
            self.R = np.zeros((A.shape[0], 0), dtype=A.dtype)
        elif engine != "lapack":
            raise ValueError(f"unknown ls engine {engine!r}")

    def add_column(self, x: SparseVec, j: int) -> bool:
        """`addindex!(x, AiQR, a, i)` (`src/util.jl:118-126`)."""
        if j in x:
            return False
        x.setindex(j, np.nan)                               # util.jl:120
        pos = x.nzind.index(j)                              # util.jl:122  findfirst(==(i), x.nzind)
        if self.engine == "givens":
            self.qr.add_column(self.A[:, j], pos)           # util.jl:123
        elif self.engine == "scipy":
            from scipy.linalg import qr_insert
            self.Q, self.R = qr_insert(self.Q, self.R, self.A[:, j], pos, which="col")
        return True

    def solve(self, x: SparseVec, b: np.ndarray) -> None:
        """`ldiv!!(x.nzval, AiQR, b, r)` (`src/matchingpursuit.jl:170-176`): LS coefficients, sorted order."""
        T = self.A.dtype
        if self.engine == "givens":
            y = self.qr.solve(np.asarray(b, dtype=T))
        elif self.engine == "scipy":
            from scipy.linalg import solve_triangular
            t = self.R.shape[1]
            y = solve_triangular(self.R[:t, :t], (self.Q.T @ np.asarray(b, dtype=T))[:t])
        else:
            AS = self.A[:, np.asarray(x.nzind, dtype=np.int64)]
            y, *_ = np.linalg.lstsq(AS, np.asarray(b, dtype=T), rcond=None)
        x.nzval = [float(v) for v in np.asarray(y, dtype=np.float64)]


# ----------------------------------------------------------------------------------------
# Matching pursuit  (`src/matchingpursuit.jl:10-40`)
# ----------------------------------------------------------------------------------------
def mp(A: np.ndarray, b: np.ndarray, k: int, x: Optional[SparseVec] = None,
       trace: Optional[Trace] = None) -> SparseVec:
    """`mp(A,b,k,x=spzeros(N))` (`src/matchingpursuit.jl:34-40`): exactly k updates, no stopping rule."""
    _check_finite(A, b)
    M, N = A.shape
    x = SparseVec(N) if x is None else x
    for _ in range(k):
        r = residual(A, x, b)                                # :27
        i = argmaxinner(A, r, trace)                         # :28
        x.setindex(i, x.getindex(i) + float(np.dot(A[:, i], r)))   # :29  signed, recomputed dot
        if trace is not None:
            trace.selected.append([i])
            trace.added.append([i])
            trace.resnorm.append(float(np.linalg.norm(residual(A, x, b))))
            trace.iterations += 1
    return x


# ----------------------------------------------------------------------------------------
# Orthogonal matching pursuit  (`src/matchingpursuit.jl:44-91`)
# ----------------------------------------------------------------------------------------
def _omp_update(A, b, x: SparseVec, ls: _ActiveSetLS, trace: Optional[Trace]) -> None:
    """`update!(P::OMP, x)` (`src/matchingpursuit.jl:62-70`)."""
    if not x.nnz() < A.shape[0]:                             # :63
        if trace is not None:
            trace.selected.append([]); trace.added.append([])
        return
    r = residual(A, x, b)                                    # :64
    i = argmaxinner(A, r, trace)                             # :65  over ALL atoms, active ones included
    if trace is not None:
        trace.selected.append([i])
    if i in x:                                               # :66  iteration consumed, nothing changes
        if trace is not None:
            trace.added.append([])
        return
    ls.add_column(x, i)                                      # :67
    ls.solve(x, b)                                           # :68
    if trace is not None:
        trace.added.append([i])


def omp(A: np.ndarray, b: np.ndarray, k: Optional[int] = None, eps: Optional[float] = None,
        ls: str = "lapack", trace: Optional[Trace] = None) -> SparseVec:
    """`omp(A,b,eps,k=size(A,1))` / `omp(A,b,k)` (`src/matchingpursuit.jl:73-86`)."""
    _check_finite(A, b)
    M, N = A.shape
    eps = eps_of(A.dtype) if eps is None else eps            # :85
    k = M if k is None else k                                # :73 default
    if not eps >= 0:
        raise ValueError(f"ε = {eps} has to be non-negative")   # :74 (the reference throws a String)
    engine = _ActiveSetLS(A, ls, k)                          # :75
    x = SparseVec(N)                                         # :76
    for _ in range(k):                                       # :77
        _omp_update(A, b, x, engine, trace)                  # :78
        nr = float(np.linalg.norm(residual(A, x, b)))        # :79
        if trace is not None:
            trace.resnorm.append(nr); trace.iterations += 1
        if not nr >= eps:
            break
    return x


# ----------------------------------------------------------------------------------------
# Generalized OMP  (`src/matchingpursuit.jl:95-148`)
# ----------------------------------------------------------------------------------------
def _gomp_update(A, b, x: SparseVec, ls: _ActiveSetLS, l: int, trace: Optional[Trace]) -> None:
    """`update!(P::GOMP, x, l)` (`src/matchingpursuit.jl:116-123`)."""
    if not x.nnz() < A.shape[0]:                             # :117
        if trace is not None:
            trace.selected.append([]); trace.added.append([])
        return
    r = residual(A, x, b)                                    # :118
    idx = argmaxinner_k(A, r, l, trace)                      # :119
    added = []
    for j in idx:                                            # :120 -> util.jl:129-134, in descending-|c| order
        if ls.add_column(x, j):
            added.append(j)
    ls.solve(x, b)                                           # :121
    if trace is not None:
        trace.selected.append(list(idx)); trace.added.append(added)


def gomp(A: np.ndarray, b: np.ndarray, l: int, k: Optional[int] = None, eps: Optional[float] = None,
         ls: str = "lapack", trace: Optional[Trace] = None) -> SparseVec:
    """`gomp(A,b,l,eps,k=size(A,1))` / `gomp(A,b,l,k)` (`src/matchingpursuit.jl:126-143`)."""
    _check_finite(A, b)
    M, N = A.shape
    eps = eps_of(A.dtype) if eps is None else eps
    k = M if k is None else k
    if not eps >= 0:
        raise ValueError(f"ε = {eps} has to be non-negative")   # :127
    engine = _ActiveSetLS(A, ls, M)                          # :128  (k is NOT forwarded: capacity M)
    x = SparseVec(N)                                         # :129
    for _ in range(k // l):                                  # :130
        _gomp_update(A, b, x, engine, l, trace)              # :131
        nr = float(np.linalg.norm(residual(A, x, b)))        # :132
        if trace is not None:
            trace.resnorm.append(nr); trace.iterations += 1
        if not nr >= eps:
            break
    rem = k % l                                              # :134
    if rem > 0:                                              # :135-137  runs even after an eps-break
        _gomp_update(A, b, x, engine, rem, trace)
        if trace is not None:
            trace.resnorm.append(float(np.linalg.norm(residual(A, x, b)))); trace.iterations += 1
    return x


# ----------------------------------------------------------------------------------------
# Forward regression / OLS / OOMP / ORMP  (`src/forward.jl:1-114`)  -- SURVEY.md 8(f) rank 1
# ----------------------------------------------------------------------------------------
def findmax_first(v: np.ndarray):
    """The package's own `Base.findmax(f, x::AbstractVector)` override (`src/util.jl:173-189`): a strict-`<` scan of
    `-f(x)` starting from `(k, m) = (0, Inf)`: first index among equal maxima, NaN entries never win.
    Returns (max, index); index -1 if nothing compared below Inf (all NaN / -Inf... the reference would index x[0])."""
    k, m = -1, np.inf
    for i, vi in enumerate(v):
        g = -vi
        if g < m:
            k, m = i, g
    return -m, k


def ols_rescaling(A: np.ndarray, x: SparseVec) -> np.ndarray:
    """`ols_rescaling!(P, x)` (`src/forward.jl:97-114`): squared norm of every atom after projecting out the active set,
    `sum(abs2, A[:, j]) - sum(abs2, (Q'A)[1:nnz(x), j])` with Q the orthogonal factor of the active columns."""
    T = A.dtype
    resc = (A * A).sum(axis=0).astype(T)                                  # :105
    if x.nnz():
        Q1, _ = np.linalg.qr(A[:, np.asarray(x.nzind, dtype=np.int64)])   # P.AiQR.Q, first nnz(x) columns (:99,:106)
        QA = (Q1.T @ A).astype(T)                                         # :104
        for i in range(QA.shape[0]):                                      # :108-112
            resc = resc - QA[i, :] ** 2
    return resc


def forward_delta(A: np.ndarray, b: np.ndarray, x: SparseVec) -> np.ndarray:
    """`forward_δ!(P, x)` (`src/forward.jl:69-76`): decrease of the SQUARED residual norm for every passive atom."""
    r = residual(A, x, b)                                                 # :70
    d = (A.T @ r).astype(np.float64)                                      # :71  (P.δ² is a Float64 vector, :29)
    resc = ols_rescaling(A, x)                                            # :72
    with np.errstate(divide="ignore", invalid="ignore"):
        d = d * d / resc                                                  # :73
    if x.nnz():
        d[np.asarray(x.nzind, dtype=np.int64)] = 0.0                      # :74
    return d


def fr(A: np.ndarray, b: np.ndarray, max_eps: float = 0.0, min_delta: float = 0.0, k: Optional[int] = None,
       ls: str = "lapack", trace: Optional[Trace] = None) -> SparseVec:
    """`fr(A, b, max_ε, min_δ, k=size(A,1))` (`src/forward.jl:44-51`) == ols == oomp == ormp (:52-54);
    the keyword form `fr(A, b; max_residual=0, min_decrease=0, sparsity=size(A,2))` (:33-36) maps onto it."""
    _check_finite(A, b)
    M, N = A.shape
    k = M if k is None else k
    engine = _ActiveSetLS(A, ls, M)                                       # :27  UpdatableQR(A[:, nzind])
    x = SparseVec(N)
    for _ in range(k):                                                    # :47
        # ---- forward_step!(P, x, max_ε, min_δ)  (:56-67) ----
        if not x.nnz() < M:                                               # :57
            break
        normr = float(np.linalg.norm(residual(A, x, b)))                  # :58-59
        if not normr > max_eps:                                           # :60
            break
        d2 = forward_delta(A, b, x)                                       # :61
        max_d2, i = findmax_first(d2)                                     # :62
        if min_delta ** 2 < max_d2:                                       # :63
            engine.add_column(x, i)                                       # :65
            engine.solve(x, b)                                            # :66
            if trace is not None:
                top = np.sort(d2[np.isfinite(d2)])[::-1]
                trace.margin.append(float((top[0] - top[1]) / top[0]) if top.size > 1 and top[0] > 0 else 1.0)
                trace.selected.append([i]); trace.added.append([i])
                trace.resnorm.append(float(np.linalg.norm(residual(A, x, b)))); trace.iterations += 1
        else:
            if x.nnz():
                engine.solve(x, b)                                        # :68
            break                                                         # :69 returns false
    return x


ols = oomp = ormp = fr


# ----------------------------------------------------------------------------------------
# Oblivious selection and subspace pursuit  (`src/oblivious.jl:3-8`, `src/twostage.jl:49-123`)  -- SURVEY.md 8(f) rank 2
# ----------------------------------------------------------------------------------------
def _dense_ls(A: np.ndarray, nzind: Sequence[int], b: np.ndarray) -> List[float]:
    """`solve!(P, x)` (`src/twostage.jl:120-123`) = `ldiv!(x.nzval, qr!(A[:, nzind]), b)` / `A[:, nzind] \\ b`."""
    y, *_ = np.linalg.lstsq(A[:, np.asarray(nzind, dtype=np.int64)], np.asarray(b, dtype=A.dtype), rcond=None)
    return [float(v) for v in np.asarray(y, dtype=np.float64)]


def oblivious(A: np.ndarray, b: np.ndarray, k: int) -> SparseVec:
    """`oblivious(A, b, k)` (`src/oblivious.jl:3-8`).  The reference sizes its result `spzeros(size(b))` (length M,
    a latent bug for indices > M); the restatement uses length N like every other algorithm."""
    _check_finite(A, b)
    nzind = argmaxinner_k(A, np.asarray(b, dtype=A.dtype), k)             # :4  partialsortperm(abs.(A'b), 1:k, rev=true)
    x = SparseVec(A.shape[1])
    for j in nzind:
        x.setindex(j, np.nan)
    x.nzval = _dense_ls(A, x.nzind, b)                                    # :6
    return x


def _sp_acquisition(A, b, x: SparseVec, k: int, trace: Optional[Trace]) -> None:
    """`sp_acquisition!(P, x, k)` (`src/twostage.jl:87-92`)."""
    r = residual(A, x, b)                                                 # :88
    idx = argmaxinner_k(A, r, k, trace)                                   # :89
    for j in idx:
        x.setindex(j, np.nan)                                             # :90  (an active atom is overwritten, not duplicated)
    x.nzval = _dense_ls(A, x.nzind, b)                                    # :91
    if trace is not None:
        trace.selected.append(list(idx))


def sp(A: np.ndarray, b: np.ndarray, k: int, delta: float = 1e-12, maxiter: Optional[int] = None,
       trace: Optional[Trace] = None) -> SparseVec:
    """`sp(A, b, k, δ = 1e-12; maxiter = 16k)` (`src/twostage.jl:105-117`) with `update!(P::SP, x)` (:94-103)."""
    _check_finite(A, b)
    M, N = A.shape
    if 2 * k > M:
        raise ValueError(f"2k = {2 * k} > {M} = length(b) is invalid for Subspace Pursuit")   # :62
    maxiter = 16 * k if maxiter is None else maxiter
    x = SparseVec(N)
    _sp_acquisition(A, b, x, k, trace)                                    # :108
    resnorm = float(np.linalg.norm(residual(A, x, b)))                    # :109
    if trace is not None:
        trace.resnorm.append(resnorm)
    for _ in range(maxiter):                                              # :110
        oldnorm = resnorm
        if x.nnz() != k:
            raise ValueError(f"nnz(x) = {x.nnz()} ≠ {k} = k")             # :95
        _sp_acquisition(A, b, x, k, trace)                                # :96
        absval = np.abs(np.asarray(x.nzval))
        ndrop = x.nnz() - k
        order = np.lexsort((np.arange(absval.size), absval))
        drop = order[:ndrop]                                              # :97 partialsortperm: smallest, ties -> lower position
        if trace is not None and 0 < ndrop < absval.size:                 # how decisive the pruning was
            lo, hi = float(absval[order[ndrop - 1]]), float(absval[order[ndrop]])
            trace.margin.append((hi - lo) / hi if hi > 0 else 0.0)
        for i in sorted(drop.tolist(), reverse=True):                     # :98-100
            del x.nzind[i]; del x.nzval[i]
        x.nzval = _dense_ls(A, x.nzind, b)                                # :101
        resnorm = float(np.linalg.norm(residual(A, x, b)))                # :112
        if trace is not None:
            trace.resnorm.append(resnorm); trace.iterations += 1
            trace.added.append(list(x.nzind))
        if resnorm <= delta or oldnorm <= resnorm:                        # :113
            break
    return x


# ----------------------------------------------------------------------------------------
# Dictionary analysis  (`src/util.jl:2, 59-61, 96-117`)  -- SURVEY.md 8(f) rank 4
# ----------------------------------------------------------------------------------------
def colnorms(A: np.ndarray) -> np.ndarray:
    """`colnorms(A) = [norm(a) for a in eachcol(A)]` (`src/util.jl:2`)."""
    return np.array([np.linalg.norm(A[:, j]) for j in range(A.shape[1])], dtype=A.dtype)


def normalize(A: np.ndarray) -> np.ndarray:
    """`normalize!(A)` (`src/util.jl:59-61`)."""
    A /= colnorms(A)[None, :]
    return A


def cumbabel(A: np.ndarray, k: int) -> np.ndarray:
    """`cumbabel(A, k)` (`src/util.jl:106-117`): all Babel-function values mu_1(1..k)."""
    mu = np.zeros(k, dtype=A.dtype)                                       # :107
    for i in range(A.shape[1]):                                           # :109
        inner = np.abs(A.T @ A[:, i])                                     # :110-111
        inner[i] = 0                                                      # :112
        top = np.sort(inner)[::-1][:k]                                    # :113 partialsort!(inner, 1:k, rev=true)
        mu = np.maximum(mu, np.cumsum(top))                               # :114-115
    return mu


def babel(A: np.ndarray, k: int):
    """`babel(A, k) = cumbabel(A, k)[k]` (`src/util.jl:101`)."""
    return cumbabel(A, k)[k - 1]


def coherence(A: np.ndarray):
    """`coherence(A) = babel(A, 1)` (`src/util.jl:98`)."""
    return babel(A, 1)


# ----------------------------------------------------------------------------------------
# synthetic data  (`src/util.jl:13-33,50-55`) -- distribution restated, not Julia's RNG stream
# ----------------------------------------------------------------------------------------
def sparse_vector(rng: np.random.Generator, m: int, k: int, gaussian: bool = False) -> SparseVec:
    """`sparse_vector(m,k,gaussian)` (`src/util.jl:13-19`)."""
    if m < k:
        raise ValueError(f"m = {m} < {k} = k")
    ind = np.sort(rng.choice(m, size=k, replace=False))
    val = rng.standard_normal(k) if gaussian else rng.choice(np.array([-1.0, 1.0]), size=k)
    return SparseVec(m, [int(i) for i in ind], [float(v) for v in val])


def gaussian_dictionary(rng: np.random.Generator, n: int, m: int, dtype=np.float64) -> np.ndarray:
    """The dictionary part of `sparse_data` (`src/util.jl:21-28`): Gaussian, eps-mean-shifted, unit columns."""
    A = rng.standard_normal((n, m))
    A -= 1e-6 * A.mean(axis=0, keepdims=True)
    A /= np.sqrt((A * A).sum(axis=0, keepdims=True))
    return np.asfortranarray(A.astype(dtype))


def sparse_data(rng: np.random.Generator, n: int = 32, m: int = 64, k: int = 3, dtype=np.float64):
    """`sparse_data(n,m,k)` (`src/util.jl:21-33`): returns (A, x0, b = A*x0)."""
    A = gaussian_dictionary(rng, n, m, dtype)
    x = sparse_vector(rng, m, k)
    b = (A[:, np.asarray(x.nzind)].astype(np.float64) @ np.asarray(x.nzval)).astype(dtype)
    return A, x, b


def perturb(rng: np.random.Generator, b: np.ndarray, delta: float) -> np.ndarray:
    """`perturb(b, δ)` (`src/util.jl:50-55`): add Gaussian noise rescaled to norm exactly δ."""
    e = rng.standard_normal(b.shape)
    e *= delta / np.linalg.norm(e)
    return (b + e).astype(b.dtype)
