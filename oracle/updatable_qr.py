"""Updatable thin QR with column insertion -- TEST INFRASTRUCTURE ONLY (part of the CPU oracle).

Restates the published contract of UpdatableQRFactorizations.jl v1.0.0 (pinned at
`/root/reference/Manifest.toml:446-450`, source NOT vendored in /root/reference) as used at
`/root/reference/src/util.jl:123` (`add_column!(AiQR, a, qr_i)`) and
`/root/reference/src/matchingpursuit.jl:175` (`ldiv!(AiQR, r)`): keep a QR factorisation of
the active columns `A[:, sort(S)]` under insertion of a column at an arbitrary position and
solve the least-squares problem against it.

Algorithm (Daniel/Gragg/Kaufman/Stewart 1976, the textbook "insert a column" update the
comment at `src/util.jl:121` refers to as "update qr factorization using Givens rotations"):
    w = Q'a,  v = a - Q w  (twice, for orthogonality),  rho = ||v||,  q_new = v / rho
    R~ = [R[:, :p]  [w; rho]  R[:, p:]]  with a zero last row under the old columns
    Givens rotations on row pairs (t-1,t), (t-2,t-1), ..., (p,p+1) push the spike back to
    upper-triangular; the same rotations are applied to the columns of [Q q_new].
PARITY: unpinned against the package's bits (source unavailable); the result is pinned to
the dense LS solution, which is what `test/forward.jl:24-28` asserts.
"""
from __future__ import annotations

import numpy as np


class UpdatableQR:
    def __init__(self, dtype, n: int, capacity: int):
        self.dtype = np.dtype(dtype)
        self.n = int(n)
        self.capacity = int(capacity)
        self.t = 0
        self.Q = np.zeros((n, 0), dtype=self.dtype, order="F")
        self.R = np.zeros((0, 0), dtype=self.dtype, order="F")

    def size(self):
        return (self.n, self.t)

    def add_column(self, a: np.ndarray, pos: int | None = None) -> None:
        t = self.t
        if t >= min(self.capacity, self.n):
            raise ValueError("UpdatableQR is at capacity")
        pos = t if pos is None else int(pos)
        a = np.asarray(a, dtype=self.dtype)
        Q, R = self.Q, self.R
        w = Q.T @ a
        v = a - Q @ w
        w2 = Q.T @ v                       # second Gram-Schmidt sweep ("twice is enough")
        v = v - Q @ w2
        w = w + w2
        rho = np.linalg.norm(v)
        q = v / rho
        Rn = np.zeros((t + 1, t + 1), dtype=self.dtype, order="F")
        Rn[:t, :pos] = R[:, :pos]
        Rn[:t, pos] = w
        Rn[t, pos] = rho
        Rn[:t, pos + 1:] = R[:, pos:]
        Qn = np.empty((self.n, t + 1), dtype=self.dtype, order="F")
        Qn[:, :t] = Q
        Qn[:, t] = q
        for i in range(t, pos, -1):        # zero Rn[i, pos] against Rn[i-1, pos]
            x, y = Rn[i - 1, pos], Rn[i, pos]
            h = np.hypot(x, y)
            if h == 0:
                continue
            c, s = x / h, y / h
            G = np.array([[c, s], [-s, c]], dtype=self.dtype)
            Rn[[i - 1, i], :] = G @ Rn[[i - 1, i], :]
            Rn[i, pos] = 0
            Qn[:, [i - 1, i]] = Qn[:, [i - 1, i]] @ G.T
        self.Q, self.R, self.t = Qn, Rn, t + 1

    def solve(self, b: np.ndarray) -> np.ndarray:
        """`ldiv!(F, r)` / `F \\ b`: argmin ||A_S y - b||, coefficients in logical (column) order."""
        z = self.Q.T @ np.asarray(b, dtype=self.dtype)
        y = np.zeros(self.t, dtype=self.dtype)
        for i in range(self.t - 1, -1, -1):
            y[i] = (z[i] - self.R[i, i + 1:] @ y[i + 1:]) / self.R[i, i]
        return y
