"""Seeded random sweep of shapes x algorithms x kernel paths against the CPU oracle (run with `-m gpu` on a B200).

Every case draws (M, N, k, batch, dtype, noise, un-normalised atoms or not) from a seeded generator, solves the whole
batch through the C ABI and compares a sample of signals with the oracle: selection sequence bit-exact (omp, gomp, fr),
supports as sets (sp, oblivious), coefficients / residual norms within 1e-10 (FP64) or 2e-5 (FP32 dictionaries).
Decisions the oracle itself flags as knife-edge (relative margin < 1e-8 between the winner and the runner-up) are
skipped: no two summation orders agree on those."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = {np.float64: 1e-10, np.float32: 2e-5}


def _case(po, seed, f32_ok=True):
    rng = np.random.default_rng(seed)
    M = int(rng.choice([24, 48, 70, 128, 200, 333]))
    N = int(rng.integers(M + 8, 6 * M))
    k = int(rng.integers(2, max(3, M // 8)))
    B = int(rng.choice([1, 3, 23, 24, 40, 130]))
    dtype = np.float32 if (f32_ok and rng.random() < 0.25) else np.float64
    A = po.gaussian_dictionary(rng, M, N, dtype)
    if rng.random() < 0.4:
        A = np.asfortranarray(A * rng.uniform(0.5, 2.0, size=(1, N)).astype(dtype))
    noise = float(rng.choice([0.0, 5e-3, 1e-2]))
    cols = []
    for _ in range(B):
        x0 = po.sparse_vector(rng, N, k)
        b = (A[:, x0.nzind].astype(np.float64) @ np.asarray(x0.nzval)).astype(dtype)
        if noise:
            b = po.perturb(rng, b, noise)
        cols.append(b)
    return rng, A, np.asfortranarray(np.stack(cols, axis=1)), k, noise, dtype


def _sorted(sel_row, coef_row, n):
    idx, val = sel_row[:n], coef_row[:n]
    o = np.argsort(idx, kind="stable")
    return idx[o], val[o]


def _close(a, b, rtol):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(1.0, float(np.max(np.abs(b))) if b.size else 1.0)
    return np.allclose(a, b, rtol=rtol, atol=rtol * scale)


def _sample(B):
    return sorted(set([0, B - 1, B // 2] + list(range(0, B, max(1, B // 4)))))


@pytest.mark.parametrize("seed", range(48))
def test_fuzz_omp_gomp(cs, po, seed):
    rng, A, Bm, k, noise, dtype = _case(po, 1000 + seed)
    rtol = RTOL[dtype]
    eps = 0.0 if noise else None                 # noiseless: default eps stops before r reaches rounding level
    l = int(rng.integers(2, 10))
    B = Bm.shape[1]
    with cs.Dictionary(A) as D, cs.Batch(D, B, min(k + l, A.shape[0])) as batch:
        batch.upload(Bm)
        e = float(np.finfo(dtype).eps) if eps is None else eps
        batch.omp(k, e)
        sel, coef, nnz, res, its = batch.download(k)
        kg = min(k + l - 1, A.shape[0])
        batch.gomp(l, kg, e)
        gsel, gcoef, gnnz, gres, gits = batch.download(kg)
    for s in _sample(B):
        t = po.Trace()
        ref = po.omp(A, Bm[:, s], k, eps=eps, trace=t)
        if min(t.margin) > 1e-8 * (1e4 if dtype == np.float32 else 1):
            n = int(nnz[s])
            assert sel[s, :n].tolist() == t.order(), (seed, s, "omp order")
            idx, val = _sorted(sel[s], coef[s], n)
            assert _close(val, ref.nzval, rtol), (seed, s)
            assert abs(res[s] - t.resnorm[-1]) <= rtol * max(1.0, float(np.linalg.norm(Bm[:, s]))) + 1e-6 * (dtype == np.float32)
        t = po.Trace()
        ref = po.gomp(A, Bm[:, s], l, kg, eps=eps, trace=t)
        if min(t.margin) > 1e-8 * (1e4 if dtype == np.float32 else 1) and (noise or dtype == np.float64):
            n = int(gnnz[s])
            if not noise and t.resnorm[-1] < 1e-9:
                # exact recovery reached: later picks are decided by rounding noise; compare the solution only
                assert np.allclose(np.sort(np.abs(gcoef[s, :n]))[::-1][:k], np.sort(np.abs(ref.nzval))[::-1][:k], rtol=1e-8, atol=1e-8)
                continue
            assert gsel[s, :n].tolist() == t.order(), (seed, s, "gomp order")
            idx, val = _sorted(gsel[s], gcoef[s], n)
            assert _close(val, ref.nzval, rtol), (seed, s)


@pytest.mark.parametrize("seed", range(32))
def test_fuzz_fr(cs, po, seed):
    rng, A, Bm, k, noise, dtype = _case(po, 2000 + seed, f32_ok=False)
    B = Bm.shape[1]
    max_eps = float(rng.choice([0.0, 0.02]))
    with cs.Dictionary(A) as D, cs.Batch(D, B, k) as batch:
        batch.upload(Bm)
        batch.fr(k, max_eps, 0.0)
        sel, coef, nnz, res, its = batch.download(k)
    for s in _sample(B):
        t = po.Trace()
        ref = po.fr(A, Bm[:, s], max_eps, 0.0, k, trace=t)
        if t.margin and min(t.margin) < 1e-8:
            continue
        n = int(nnz[s])
        assert n == ref.nnz(), (seed, s, n, ref.nnz())
        assert sel[s, :n].tolist() == t.order(), (seed, s, "fr order")
        idx, val = _sorted(sel[s], coef[s], n)
        assert _close(val, ref.nzval, 1e-10), (seed, s)


@pytest.mark.parametrize("seed", range(32))
def test_fuzz_sp_oblivious_babel(cs, po, seed):
    rng, A, Bm, k, noise, dtype = _case(po, 3000 + seed)
    rtol = RTOL[dtype]
    M, N = A.shape
    B = Bm.shape[1]
    k = min(k, M // 2)
    with cs.Dictionary(A) as D:
        with cs.Batch(D, B, 2 * k) as batch:
            batch.upload(Bm)
            batch.sp(k, 1e-12 if dtype == np.float64 else 1e-4)
            sel, coef, nnz, res, its = batch.download(k)
            batch.oblivious(k)
            osel, ocoef, onnz, ores, _ = batch.download(k)
        kb = int(rng.integers(1, min(N - 1, 40)))
        mu = cs.cumbabel(D, kb)
    assert np.allclose(mu, po.cumbabel(A, kb), rtol=1e-12 if dtype == np.float64 else 1e-5)
    thr = 1e-8 * (1e4 if dtype == np.float32 else 1)
    for s in _sample(B):
        t = po.Trace()
        ref = po.sp(A, Bm[:, s], k, 1e-12 if dtype == np.float64 else 1e-4, trace=t)
        # SP's loop compares successive residual norms; at rounding level (noiseless recovery) that test is a coin flip
        if min(t.margin) > thr and (noise or dtype == np.float64) and not (t.resnorm[-1] < 1e-9 and t.iterations > 1):
            idx, val = _sorted(sel[s], coef[s], int(nnz[s]))
            assert idx.tolist() == ref.nzind, (seed, s, "sp support", min(t.margin))
            assert _close(val, ref.nzval, rtol), (seed, s)
        ref = po.oblivious(A, Bm[:, s], k)
        c = np.sort(np.abs(A.astype(np.float64).T @ Bm[:, s].astype(np.float64)))[::-1]
        if k < N and (c[k - 1] - c[k]) > thr * c[0]:
            idx, val = _sorted(osel[s], ocoef[s], int(onnz[s]))
            assert idx.tolist() == ref.nzind, (seed, s, "oblivious support")
            assert _close(val, ref.nzval, rtol), (seed, s)
