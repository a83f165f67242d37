#!/usr/bin/env python
"""Coefficient error of the GPU least-squares scheme (implicit Q = A_S R^{-1}, CGS with DGKS re-orthogonalisation, explicit
R^{-1}) against the conditioning of the active set -- VERDICT r1 item 8.

A pair of nearly parallel atoms (a_1 = normalize(a_0 + w / c), w a unit vector orthogonal to a_0) puts cond(A_S) at about
2c.  For c = 1e2 .. 1e8 a signal planted on {a_0, a_1, six more atoms} is solved by omp / gomp on every kernel path and the
coefficients are compared with (i) the planted ones and (ii) LAPACK's least squares on the recovered support, in units of
cond(A_S) * eps -- the error level of a backward-stable QR (the reference's Givens scheme).  Prints one JSON line per case.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

EPS = float(np.finfo(np.float64).eps)


def twin_problem(rng, M, N, c, nsig):
    A = rng.standard_normal((M, N))
    A /= np.linalg.norm(A, axis=0, keepdims=True)
    w = rng.standard_normal(M)
    w -= A[:, 0] * (A[:, 0] @ w)
    w /= np.linalg.norm(w)
    a1 = A[:, 0] + w / c
    A[:, 1] = a1 / np.linalg.norm(a1)
    A = np.asfortranarray(A)
    S = np.concatenate([[0, 1], rng.choice(np.arange(2, N), size=6, replace=False)])
    coefs = rng.choice([-1.0, 1.0], size=(nsig, S.size)) * rng.uniform(0.5, 1.5, size=(nsig, S.size))
    Bm = np.asfortranarray(A[:, S] @ coefs.T)
    return A, S, coefs, Bm


def run(cs, name, A, S, coefs, Bm, env, algo):
    for key in ("CSB200_SMALL_SOLVE", "CSB200_PERSIST", "CSB200_GOMP_BLOCK"):
        os.environ.pop(key, None)
    os.environ.update(env)
    k = S.size
    with cs.Dictionary(A) as D:
        xs = cs.omp(D, Bm, 0.0, k) if algo == "omp" else cs.gomp(D, Bm, 2, 0.0, k)
    if not isinstance(xs, list):
        xs = [xs]
    order = np.argsort(S)
    cond = float(np.linalg.cond(A[:, S]))
    worst_planted, worst_ls, recovered, both, worst_any = 0.0, 0.0, 0, 0, 0.0
    for s, x in enumerate(xs):
        # whatever support came back: the coefficients must be its least-squares solution to ~cond(A_support) eps
        sup = x.nzind
        ls_any = np.linalg.lstsq(A[:, sup], Bm[:, s], rcond=None)[0]
        c_any = float(np.linalg.cond(A[:, sup]))
        both += int(0 in sup and 1 in sup)
        worst_any = max(worst_any, float(np.max(np.abs(x.nzval - ls_any)) / np.max(np.abs(ls_any))) / (c_any * EPS))
        if sup.tolist() != np.sort(S).tolist():
            continue
        recovered += 1
        want = coefs[s][order]
        worst_planted = max(worst_planted, float(np.max(np.abs(x.nzval - want)) / np.max(np.abs(want))))
        worst_ls = max(worst_ls, float(np.max(np.abs(x.nzval - ls_any)) / np.max(np.abs(ls_any))))
    return {"path": name, "algo": algo, "cond": cond, "signals": len(xs), "support_recovered": recovered,
            "both_twins_selected": both, "err_vs_planted": worst_planted, "err_vs_lapack": worst_ls,
            "err_vs_planted_over_cond_eps": worst_planted / (cond * EPS), "err_vs_lapack_over_cond_eps": worst_ls / (cond * EPS),
            "any_support_err_vs_lapack_over_its_cond_eps": worst_any}


def main():
    cs = ge.load_package()
    rng = np.random.default_rng(8)
    M, N = 256, 1024
    paths = [("batched: DMMA pass + CTA update", 64, {"CSB200_SMALL_SOLVE": "0"}),
             ("single signal: cooperative whole-solve kernel", 1, {"CSB200_PERSIST": "1"}),
             ("single signal: GEMV + cluster update", 1, {"CSB200_PERSIST": "0", "CSB200_SMALL_SOLVE": "0"}),
             ("48 signals: one-CTA whole-solve kernel", 48, {"CSB200_SMALL_SOLVE": "1"})]
    for c in (1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8):
        for name, nsig, env in paths:
            A, S, coefs, Bm = twin_problem(rng, M, N, c, nsig)
            for algo in ("omp", "gomp"):
                print(json.dumps(run(cs, name, A, S, coefs, Bm, env, algo)), flush=True)


if __name__ == "__main__":
    main()
