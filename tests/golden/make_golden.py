"""Generates tests/golden/*.npz: seeded inputs (as bytes, not seeds-in-two-languages) and the oracle's outputs.

The reference ships no golden vectors for this path and Julia is not available in this image, so the
fixtures are produced by the CPU oracle (oracle/pursuit_oracle.py), which restates the reference line by
line.  They pin (a) the oracle against silent changes and (b) the CUDA path against the oracle on the GPU
box, where /root/reference and Julia do not exist either.      Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pursuit_oracle as po  # noqa: E402


def build(name, algo, M, N, k, B, seed, dtype=np.float64, l=None, noise=0.0, eps=None, planted=None, colscale=False,
          max_eps=0.0, min_delta=0.0):
    rng = np.random.default_rng(seed)
    A = po.gaussian_dictionary(rng, M, N, dtype)
    if colscale:                     # un-normalised atoms: the OLS rescaling ||a||^2 - ||Q'a||^2 starts away from 1
        A = np.asfortranarray(A * rng.uniform(0.5, 2.0, size=(1, N)).astype(dtype))
    planted = k if planted is None else planted
    cols = []
    for _ in range(B):
        x0 = po.sparse_vector(rng, N, planted)
        b = (A[:, x0.nzind].astype(np.float64) @ np.asarray(x0.nzval)).astype(dtype)
        if noise:
            b = po.perturb(rng, b, noise)
        cols.append(b)
    Bm = np.stack(cols, axis=1)
    kk = k
    nzind = -np.ones((B, kk), dtype=np.int64)
    nzval = np.zeros((B, kk))
    order = -np.ones((B, kk), dtype=np.int64)
    nnz = np.zeros(B, dtype=np.int64)
    resn = np.zeros(B)
    margin = np.zeros(B)
    iters_out = np.zeros(B, dtype=np.int64)
    for s in range(B):
        t = po.Trace()
        if algo == "omp":
            x = po.omp(A, Bm[:, s], k, eps=eps, trace=t)
        elif algo == "gomp":
            x = po.gomp(A, Bm[:, s], l, k, eps=eps, trace=t)
        elif algo == "fr":
            x = po.fr(A, Bm[:, s], max_eps, min_delta, k, trace=t)
        elif algo == "sp":
            x = po.sp(A, Bm[:, s], k, delta=(1e-12 if eps is None else eps), trace=t)
        elif algo == "oblivious":
            x = po.oblivious(A, Bm[:, s], k)
        else:
            x = po.mp(A, Bm[:, s], k, trace=t)
        n = x.nnz()
        nnz[s] = n
        nzind[s, :n] = x.nzind
        nzval[s, :n] = x.nzval
        o = t.order() if algo not in ("sp", "oblivious") else []
        iters_out[s] = t.iterations
        if algo != "mp":
            order[s, :len(o)] = o
        resn[s] = t.resnorm[-1] if t.resnorm else float(np.linalg.norm(Bm[:, s]))
        if algo == "oblivious":
            resn[s] = float(np.linalg.norm(po.residual(A, x, Bm[:, s])))
        margin[s] = min(t.margin) if t.margin else 1.0
    meta = dict(algo=algo, M=M, N=N, k=k, B=B, seed=seed, dtype=np.dtype(dtype).name, l=l, noise=noise, eps=eps,
                planted=planted, colscale=colscale, max_eps=max_eps, min_delta=min_delta)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), A=A, B=Bm, nzind=nzind, nzval=nzval, order=order, nnz=nnz,
                        resnorm=resn, margin=margin, iters=iters_out, meta=json.dumps(meta))
    print(name, "min margin", margin.min(), "max resnorm", resn.max())


def build_fr():
    """Forward regression / OLS (src/forward.jl): reference test shape, un-normalised atoms, both stopping rules."""
    build("fr_ref_32x48_k3", "fr", 32, 48, 3, 40, seed=1250)                             # test/forward.jl:15-17
    build("fr_noisy_96x200_k8", "fr", 96, 200, 8, 12, seed=1251, noise=5e-3, colscale=True)
    build("fr_stop_eps_70x130", "fr", 70, 130, 20, 8, seed=1252, planted=6, noise=5e-3, max_eps=0.05, colscale=True)
    build("fr_stop_delta_70x130", "fr", 70, 130, 20, 8, seed=1253, planted=6, noise=5e-3, min_delta=0.05)


def build_sp():
    """Subspace pursuit / oblivious selection (src/twostage.jl, src/oblivious.jl)."""
    build("sp_ref_32x48_k3", "sp", 32, 48, 3, 40, seed=1260)                             # test/twostage.jl:42-46
    build("sp_noisy_64x256_k16", "sp", 64, 256, 16, 16, seed=1261, noise=1e-2, eps=1e-2)  # several update!s per signal
    build("sp_noisy_96x400_k24", "sp", 96, 400, 24, 8, seed=1262, noise=1e-2)
    build("sp_f32_64x160_k5", "sp", 64, 160, 5, 8, seed=1263, dtype=np.float32, noise=1e-2, eps=1e-2)
    build("oblivious_70x130_k9", "oblivious", 70, 130, 9, 8, seed=1264, noise=1e-2)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "fr":
        build_fr()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sp":
        build_sp()
        sys.exit(0)
    build("omp_c1_128x256_k8", "omp", 128, 256, 8, 4, seed=1234)                      # BASELINE config 1 (KAT-1)
    build("omp_c1_noisy", "omp", 128, 256, 8, 4, seed=1235, noise=5e-3)                # KAT-2
    build("omp_ref_32x48_k3", "omp", 32, 48, 3, 40, seed=1236)                         # reference test shape (KAT-3)
    build("gomp_ref_32x48_l2_k3", "gomp", 32, 48, 3, 40, seed=1237, l=2)               # remainder step (KAT-7)
    build("gomp_96x200_l4_k8", "gomp", 96, 200, 8, 8, seed=1238, l=4)
    # KAT-8.  Noisy on purpose: on noiseless data 30 mp iterations drive r to rounding level (1e-15), where the
    # argmax is decided by the last bits and no two implementations agree (hard parts).
    build("mp_ref_32x48_k30", "mp", 32, 48, 30, 8, seed=1239, planted=3, noise=5e-3)
    build("omp_f32_64x160_k5", "omp", 64, 160, 5, 8, seed=1240, dtype=np.float32)      # KAT-9
    build("omp_ragged_70x130_k6", "omp", 70, 130, 6, 8, seed=1241)                     # M, N not multiples of 16 / 64
