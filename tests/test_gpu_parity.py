"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI, against the
CPU oracle on the same bytes.  Bars (BASELINE.json north_star): selected support sequence bit-exact,
coefficients and residual norms within 1e-10 relative in FP64; FP32 dictionaries: 2e-5 relative."""
import glob
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
RTOL64 = 1e-10
RTOL32 = 2e-5
IMPLS = {"auto": 0, "gemm": 1, "gemv": 2, "naive": 3}


def _batch(cs, D, B, kcap, impl=None):
    old = os.environ.pop("CSB200_CORR_IMPL", None)
    if impl:
        os.environ["CSB200_CORR_IMPL"] = impl
    try:
        batch = cs.Batch(D, B.shape[1], kcap)
    finally:
        os.environ.pop("CSB200_CORR_IMPL", None)
        if old:
            os.environ["CSB200_CORR_IMPL"] = old
    batch.upload(B)
    return batch


def _sorted(sel_row, coef_row, n):
    idx, val = sel_row[:n], coef_row[:n]
    o = np.argsort(idx, kind="stable")
    return idx[o], val[o]


def _close(a, b, rtol):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(1.0, float(np.max(np.abs(b))) if b.size else 1.0)
    return np.allclose(a, b, rtol=rtol, atol=rtol * scale)


# ------------------------------------------------------------------ correlation kernels
@pytest.mark.parametrize("M,N,B,s", [(128, 256, 48, 1), (70, 130, 33, 3), (1024, 700, 130, 4), (16, 64, 24, 1),
                                     (33, 1000, 257, 8)])
def test_corr_gemm_matches_numpy_topk(cs, po, M, N, B, s):
    rng = np.random.default_rng(M * 7 + N)
    A = po.gaussian_dictionary(rng, M, N)
    R = np.asfortranarray(rng.standard_normal((M, B)))
    with cs.Dictionary(A) as D, cs.Batch(D, B, 4) as batch:
        batch.upload(R)
        got = {name: batch.debug_corr_topk(s, impl) for name, impl in IMPLS.items() if name != "auto"}
    C = np.abs(A.T @ R)
    for b in range(B):
        order = np.lexsort((np.arange(N), -C[:, b]))[:s]
        gap = np.min(np.abs(np.diff(np.sort(C[:, b])[::-1][: s + 1]))) if N > s else 1.0
        for name, (idx, val) in got.items():
            assert np.allclose(val[b], C[order, b], rtol=1e-12, atol=1e-13), (name, b)
            if gap > 1e-11:
                assert idx[b].tolist() == order.tolist(), (name, b)
    # the production kernels must agree with each other exactly on the indices
    assert np.array_equal(got["gemm"][0], got["gemv"][0])
    assert np.array_equal(got["gemm"][0], got["naive"][0])


def test_corr_ties_pick_lowest_index(cs, po):
    """KAT-4: duplicate / negated duplicate atoms give bit-identical |c|: the lower index must win."""
    rng = np.random.default_rng(11)
    M, N, B = 64, 320, 40
    A = po.gaussian_dictionary(rng, M, N)
    A[:, 200] = A[:, 17]          # different atom block, different CTA tile half
    A[:, 18] = -A[:, 17]          # same block
    A[:, 300] = A[:, 129]
    R = np.asfortranarray(np.stack([A[:, 17] * (1 + 0.1 * i) for i in range(B // 2)] +
                                   [A[:, 129] * (1 + 0.1 * i) for i in range(B // 2)], axis=1))
    with cs.Dictionary(A) as D, cs.Batch(D, B, 4) as batch:
        batch.upload(R)
        for impl in (1, 2, 3):
            idx, val = batch.debug_corr_topk(3, impl)
            assert (idx[: B // 2, 0] == 17).all() and (idx[: B // 2, 1] == 18).all() and (idx[: B // 2, 2] == 200).all(), impl
            assert (idx[B // 2:, 0] == 129).all() and (idx[B // 2:, 1] == 300).all(), impl
            assert np.array_equal(val[: B // 2, 0], val[: B // 2, 2])


@pytest.mark.parametrize("M,N,B,s", [(96, 700, 50, 16), (64, 1000, 130, 33), (128, 333, 24, 64), (32, 40, 30, 64)])
def test_dense_topk_radix_select(cs, po, M, N, B, s, monkeypatch):
    """Large-s path: the DMMA pass stores |A'r| and the selection is a radix select + bitonic sort.  Same (value desc,
    index asc) order as numpy and as the per-block candidate path (CSB200_DENSE_TOPK=0); exact ties (duplicate /
    negated atoms, an all-zero residual) go to the lowest indices."""
    rng = np.random.default_rng(M + N + s)
    A = po.gaussian_dictionary(rng, M, N)
    A[:, N - 1] = A[:, 3]
    A[:, N // 2] = -A[:, 3]
    A[:, 7] = A[:, 5]
    R = np.asfortranarray(rng.standard_normal((M, B)))
    R[:, 1] = 0.0
    R[:, 2] = A[:, 3] * 2.0
    with cs.Dictionary(A) as D, cs.Batch(D, B, 4) as batch:
        batch.upload(R)
        idx, val = batch.debug_corr_topk(s, 1)
        idx_v, val_v = batch.debug_corr_topk(s, 2)            # GEMV kernel: per-block candidates
    C = np.abs(A.T @ R)
    take = min(s, N)
    assert idx[1, :take].tolist() == list(range(take)) and (val[1, :take] == 0).all()
    assert idx[2, :3].tolist() == [3, N // 2, N - 1]
    for b in range(B):
        order = np.lexsort((np.arange(N), -C[:, b]))[:take]
        assert np.allclose(val[b, :take], C[order, b], rtol=1e-12, atol=1e-13), b
        assert (idx[b, take:] == -1).all()
        srt = np.sort(C[:, b])[::-1][: take + 1]
        if b not in (1, 2) and np.min(np.abs(np.diff(srt))) > 1e-11:
            assert idx[b, :take].tolist() == order.tolist(), b
    assert np.array_equal(idx, idx_v)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("M,N,B,s", [(1024, 8192, 1, 1), (1024, 3000, 3, 1), (200, 5003, 2, 4), (4096, 333, 1, 2),
                                     (48, 70001, 1, 1)])
def test_corr_gemv_regimes_are_bit_identical(cs, po, dtype, M, N, B, s, monkeypatch):
    """The GEMV kernel has an HBM instantiation (256 threads, 4-column groups) and an L2 one (1024 threads, 2-column
    groups, grid = multiple of the SM count).  Every dot product keeps its summation order, so |c| and the selection
    must be BIT-identical between them, and equal to the naive kernel's selection."""
    rng = np.random.default_rng(M + N)
    A = po.gaussian_dictionary(rng, M, N, dtype)
    A[:, N - 1] = A[:, 2]                                     # an exact tie across the whole atom range
    R = np.asfortranarray(rng.standard_normal((M, B)).astype(dtype))
    R[:, 0] = A[:, 2] * 3
    got = {}
    with cs.Dictionary(A) as D, cs.Batch(D, B, 4) as batch:
        batch.upload(R)
        for regime in ("0", "1"):
            monkeypatch.setenv("CSB200_GEMV_L2", regime)
            got[regime] = batch.debug_corr_topk(s, 2)
        monkeypatch.delenv("CSB200_GEMV_L2")
        idx_n, val_n = batch.debug_corr_topk(s, 3)
    assert np.array_equal(got["0"][0], got["1"][0]) and np.array_equal(got["0"][1], got["1"][1])
    assert got["0"][0][0, 0] == 2 and (s < 2 or got["0"][0][0, 1] == N - 1)
    C = np.abs(A.astype(np.float64).T @ R.astype(np.float64))
    for b in range(B):
        order = np.lexsort((np.arange(N), -C[:, b]))[:s]
        assert np.allclose(got["1"][1][b], C[order, b], rtol=1e-12, atol=1e-13)
        if b > 0:
            assert got["1"][0][b].tolist() == order.tolist() == idx_n[b].tolist()


@pytest.mark.parametrize("M,N,s", [(64, 160, 2), (8192, 640, 1), (100, 77, 5)])
def test_corr_gemv_f32(cs, po, M, N, s):
    rng = np.random.default_rng(N)
    A = po.gaussian_dictionary(rng, M, N, np.float32)
    R = np.asfortranarray(rng.standard_normal((M, 3)).astype(np.float32))
    with cs.Dictionary(A) as D, cs.Batch(D, 3, 4) as batch:
        batch.upload(R)
        idx, val = batch.debug_corr_topk(s, 2)
        idx_n, val_n = batch.debug_corr_topk(s, 3)
    C = np.abs(A.astype(np.float64).T @ R.astype(np.float64))
    assert np.array_equal(idx, idx_n)
    for b in range(3):
        order = np.lexsort((np.arange(N), -C[:, b]))[:s]
        assert idx[b].tolist() == order.tolist()
        assert np.allclose(val[b], C[order, b], rtol=1e-12)     # FP64 accumulation of exact float products


# ------------------------------------------------------------------ golden fixtures
def _fixtures():
    return sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))


@pytest.mark.parametrize("path", _fixtures(), ids=lambda p: os.path.basename(p)[:-4])
@pytest.mark.parametrize("impl,update", [("gemm", "cta"), ("gemv", "cta"), ("gemm", "cluster"), ("gemv", "cluster"),
                                         ("gemv", "cluster16"), ("small", "cta")])
def test_golden(cs, path, impl, update, monkeypatch):
    # correlation kernel: DMMA GEMM / GEMV (multi-launch path) or the whole-solve small-dictionary kernel;
    # update kernel of the multi-launch path: one CTA per signal / one 8-CTA (or 16-CTA) cluster per signal
    monkeypatch.setenv("CSB200_CLUSTER", "16" if update == "cluster16" else "8")
    if update == "cluster16":
        update = "cluster"
    monkeypatch.setenv("CSB200_UPDATE_IMPL", update)
    if impl == "small":
        impl = None
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    if meta["algo"] == "fr" and (impl, update) != ("gemm", "cta"):
        pytest.skip("forward regression has one kernel combination (DMMA pass + CTA update)")
    if meta["algo"] in ("sp", "oblivious") and (update != "cta" or impl is None):
        pytest.skip("sp / oblivious have their own update kernel; correlation kernel: gemm or gemv")
    A, Bm = np.asfortranarray(z["A"]), np.asfortranarray(z["B"])
    f32 = A.dtype == np.float32
    if f32 and impl == "gemm":
        pytest.skip("the DMMA GEMM path is FP64-only")
    rtol = RTOL32 if f32 else RTOL64
    k = meta["k"]
    eps = meta["eps"] if meta["eps"] is not None else float(np.finfo(A.dtype).eps)
    with cs.Dictionary(A) as D:
        batch = _batch(cs, D, Bm, k, impl)
        if meta["algo"] == "omp":
            batch.omp(k, eps)
        elif meta["algo"] == "gomp":
            batch.gomp(meta["l"], k, eps)
        elif meta["algo"] == "fr":
            batch.fr(k, meta["max_eps"], meta["min_delta"])
        elif meta["algo"] == "sp":
            batch.close()
            batch = _batch(cs, D, Bm, 2 * k, impl)
            batch.sp(k, 1e-12 if meta["eps"] is None else meta["eps"])
        elif meta["algo"] == "oblivious":
            batch.oblivious(k)
        else:
            batch.mp(k)
        sel, coef, nnz, res, its = batch.download(k)
        batch.close()
    for s in range(Bm.shape[1]):
        n = int(z["nnz"][s])
        if meta["algo"] == "mp":
            acc = {}
            for i, c in zip(sel[s, :k].tolist(), coef[s, :k].tolist()):
                acc[i] = acc.get(i, 0.0) + c
            idx = np.array(sorted(acc))
            val = np.array([acc[i] for i in idx.tolist()])
            assert idx.tolist() == z["nzind"][s, :n].tolist(), (s, idx, z["nzind"][s, :n])
            assert _close(val, z["nzval"][s, :n], 1e-9 if not f32 else RTOL32), s
            continue
        assert int(nnz[s]) == n, (s, nnz[s], n)
        if meta["algo"] in ("sp", "oblivious"):             # no selection sequence: the support is a set per update!
            if meta["algo"] == "sp":
                assert int(its[s]) == int(z["iters"][s]) + 1, (s, its[s], z["iters"][s])   # + the initial acquisition
        else:
            assert sel[s, :n].tolist() == z["order"][s, :n].tolist(), (s, "selection sequence")
        idx, val = _sorted(sel[s], coef[s], n)
        assert idx.tolist() == z["nzind"][s, :n].tolist()
        assert _close(val, z["nzval"][s, :n], rtol), (s, val, z["nzval"][s, :n])
        assert abs(res[s] - z["resnorm"][s]) <= rtol * max(1.0, np.linalg.norm(Bm[:, s])) + 50 * np.finfo(A.dtype).eps


# ------------------------------------------------------------------ reference call surface + quirks
@pytest.fixture(params=["small_solve", "multi_launch", "persist", "cluster_solve"])
def solve_path(request, monkeypatch):
    """The ways a few-signal solve on a small dictionary can run: the one-CTA-per-signal whole-solve kernel
    (solve_small.cu), the per-iteration kernels (CSB200_SMALL_SOLVE=0), the cooperative whole-solve kernel
    (solve_persist.cu) and the cluster-resident whole-solve kernel (the default for <= 8 signals on dictionaries that fit
    an 8-CTA cluster's shared memory); each forced here so that nothing else can take the call."""
    if request.param == "persist":
        monkeypatch.delenv("CSB200_SMALL_SOLVE", raising=False)
        monkeypatch.setenv("CSB200_PERSIST", "1")
        monkeypatch.setenv("CSB200_CLUSTER_SOLVE", "0")
    elif request.param == "cluster_solve":
        monkeypatch.delenv("CSB200_SMALL_SOLVE", raising=False)
        monkeypatch.delenv("CSB200_PERSIST", raising=False)
        monkeypatch.setenv("CSB200_CLUSTER_SOLVE", "1")
    else:
        monkeypatch.setenv("CSB200_SMALL_SOLVE", "1" if request.param == "small_solve" else "0")
    return request.param


def test_call_surface_and_quirks(cs, po, solve_path):
    rng = np.random.default_rng(3)
    A, x0, b = po.sparse_data(rng, 32, 48, 3)
    with cs.Dictionary(A) as D:
        x = cs.omp(D, b, 3)
        assert x.nzind.tolist() == x0.nzind and np.allclose(x.nzval, x0.nzval, rtol=1e-8)
        assert x.nzval.dtype == np.float64 and x.n == 48
        ref = po.omp(A, b, eps=0.5)
        got = cs.omp(D, b, 0.5)                        # eps form, k = size(A,1): stops once ||r|| < 0.5
        assert got.nzind.tolist() == ref.nzind and _close(got.nzval, ref.nzval, RTOL64)
        got = cs.omp(D, b, max_residual=0.5, sparsity=2)
        ref = po.omp(A, b, 2, eps=0.5)
        assert got.nzind.tolist() == ref.nzind and _close(got.nzval, ref.nzval, RTOL64)
        xg = cs.gomp(D, b, 2, 3)
        ref = po.gomp(A, b, 2, 3)
        assert xg.nzind.tolist() == ref.nzind and _close(xg.nzval, ref.nzval, RTOL64)
        y = po.perturb(rng, b, 5e-3)                   # noisy: keeps r above rounding level for 30 iterations
        xm = cs.mp(D, y, 30)
        ref = po.mp(A, y, 30)
        assert xm.nzind.tolist() == ref.nzind and _close(xm.nzval, ref.nzval, 1e-9)
        # warm start (mp's optional x argument, matchingpursuit.jl:34)
        xw = cs.mp(D, y, 5, x=cs.mp(D, y, 25))
        assert xw.nzind.tolist() == ref.nzind and _close(xw.nzval, ref.nzval, 1e-9)
        with pytest.raises(ValueError, match="has to be non-negative"):
            cs.omp(D, b, -1.0, 3)
        with pytest.raises(ValueError):
            cs.omp(D, np.ones(31), 3)                  # DimensionMismatch
        bad = b.copy(); bad[3] = np.nan
        with pytest.raises(cs.CSB200Error) as ei:
            cs.omp(D, bad, 3)
        assert ei.value.status == -3
    Abad = A.copy(); Abad[0, 0] = np.inf
    with pytest.raises(cs.CSB200Error):
        cs.Dictionary(Abad)
    # one-shot form: a plain matrix instead of a Dictionary
    x = cs.omp(A, b, 3)
    assert x.nzind.tolist() == x0.nzind


def test_noop_iteration_zero_signal_and_remainder(cs, po, solve_path):
    A = np.asfortranarray(np.eye(6))
    with cs.Dictionary(A) as D:
        b = np.array([1.0, 1.0, 0, 0, 0, 0])
        with cs.Batch(D, 1, 3) as batch:               # KAT-5: eps = 0, third update re-picks atom 0: no-op
            batch.upload(b)
            batch.omp(3, 0.0)
            sel, coef, nnz, res, its = batch.download(3)
        assert nnz[0] == 2 and sel[0, :2].tolist() == [0, 1] and its[0] == 3 and res[0] == 0.0
        with cs.Batch(D, 1, 3) as batch:               # default eps: break after the second update
            batch.upload(b)
            batch.omp(3, float(np.finfo(float).eps))
            sel, coef, nnz, res, its = batch.download(3)
        assert nnz[0] == 2 and its[0] == 2
        x = cs.omp(D, np.zeros(6), 3)                  # KAT-6
        assert x.nzind.tolist() == [0] and x.nzval.tolist() == [0.0]
        x = cs.gomp(D, np.array([0, 3.0, 2.0, 0, 0, 0]), 2, 5)     # KAT-7: remainder after an eps-break
        ref = po.gomp(A, np.array([0, 3.0, 2.0, 0, 0, 0]), 2, 5)
        assert x.nzind.tolist() == ref.nzind == [0, 1, 2] and x.nzval.tolist() == ref.nzval
    # support can never exceed the number of rows (matchingpursuit.jl:63)
    rng = np.random.default_rng(9)
    A = po.gaussian_dictionary(rng, 6, 20)
    b = rng.standard_normal(6)
    got, ref = cs.omp(A, b, 0.0, 10), po.omp(A, b, 10, eps=0.0)
    assert got.nzind.tolist() == ref.nzind and got.nnz() <= 6


def test_dependent_atom_is_not_appended(cs, solve_path):
    """Duplicate columns: once one copy is active the other can only win on a zero residual; the update
    must leave a finite state (the reference's QR would divide by a zero diagonal here)."""
    A = np.asfortranarray(np.array([[1.0, 1.0, 0], [0, 0, 1.0], [0, 0, 0]]))
    x = cs.gomp(A, np.array([2.0, 0, 0]), 2, 0.0, 2)      # top-2 = atoms 0 and 1 (a tie): atom 1 duplicates atom 0
    assert x.nzind.tolist() == [0] and x.nzval.tolist() == [2.0]


def test_f32_dictionary_batches_take_the_dmma_path(cs, po, monkeypatch):
    """FP32 dictionaries: a batch of >= 24 signals runs on an FP64 twin of the dictionary (exact float x float products,
    DMMA path) -- against the FP32 oracle within the FP32 bound, against the per-signal GEMV path (CSB200_PROMOTE_F32=0),
    Float64 result vectors either way; fr becomes available for FP32 input through the one-shot call."""
    rng = np.random.default_rng(41)
    M, N, k, B = 96, 400, 7, 64
    A = po.gaussian_dictionary(rng, M, N, np.float32)
    X0, Bm = _planted(po, rng, A.astype(np.float64), k, B, noise=1e-2)
    Bm = np.asfortranarray(Bm.astype(np.float32))
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("CSB200_PROMOTE_F32", mode)
        with cs.Dictionary(A) as D, cs.Batch(D, B, k) as batch:
            batch.upload(Bm)
            batch.omp(k, 0.0)
            out[mode] = batch.download(k) + (batch.residual(),)
            batch.gomp(3, k, 0.0)
            out["g" + mode] = batch.download(k)
    monkeypatch.setenv("CSB200_PROMOTE_F32", "1")
    sel, coef, nnz, res, its, R = out["1"]
    assert R.dtype == np.float32 and coef.dtype == np.float64
    for s in range(B):
        t = po.Trace()
        ref = po.omp(A, Bm[:, s], k, eps=0.0, trace=t)
        if min(t.margin) < 1e-4:
            continue
        assert sel[s, :k].tolist() == t.order(), s
        assert out["0"][0][s, :k].tolist() == t.order(), s
        idx, val = _sorted(sel[s], coef[s], k)
        assert _close(val, ref.nzval, RTOL32), s
        assert _close(out["0"][1][s], coef[s], RTOL32)
        assert abs(res[s] - t.resnorm[-1]) < 1e-5
        assert np.allclose(R[:, s], Bm[:, s].astype(np.float64) - A[:, idx].astype(np.float64) @ val, atol=1e-5)
        t = po.Trace()
        ref = po.gomp(A, Bm[:, s], 3, k, eps=0.0, trace=t)
        if min(t.margin) > 1e-4:
            assert out["g1"][0][s, :k].tolist() == t.order() == out["g0"][0][s, :k].tolist(), s
    with cs.Dictionary(A) as D:
        xs = cs.fr(D, Bm[:, :5], sparsity=k)                 # fewer than 24 signals: the one-shot fr call still promotes
        for s in range(5):
            ref = po.fr(A.astype(np.float64), Bm[:, s].astype(np.float64), 0.0, 0.0, k)
            assert xs[s].nzind.tolist() == ref.nzind and _close(xs[s].nzval, ref.nzval, 1e-9)
        mu = cs.cumbabel(D, 5)
        assert np.allclose(mu, po.cumbabel(A, 5), rtol=1e-5)


def test_concurrent_host_threads(cs, po):
    """SURVEY 8b threading contract: calls on different handles run concurrently from different host threads; calls on
    one handle are serialised by its mutex.  ctypes releases the GIL for the duration of each C call."""
    import threading
    rng = np.random.default_rng(17)
    M, N, k, B = 96, 400, 6, 64
    probs = []
    for _ in range(4):
        A = po.gaussian_dictionary(rng, M, N)
        X0, Bm = _planted(po, rng, A, k, B, noise=5e-3)
        probs.append((A, Bm))
    shared = cs.Dictionary(probs[0][0])
    own = [cs.Dictionary(A) for A, _ in probs]
    results, errors = {}, []

    def work(t):
        try:
            A, Bm = probs[t]
            for rep in range(6):
                results[(t, "own", rep)] = cs.omp(own[t], Bm, 0.0, k, result="csc")
                results[(t, "shared", rep)] = cs.gomp(shared, probs[0][1], 2, 0.0, k, result="csc")
                results[(t, "fr", rep)] = cs.fr(own[t], Bm, sparsity=k, result="csc")
        except Exception as e:                                   # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for th in threads: th.start()
    for th in threads: th.join()
    assert not errors, errors
    for t in range(4):
        A, Bm = probs[t]
        for key in ("own", "fr"):
            first = results[(t, key, 0)]
            for rep in range(1, 6):
                other = results[(t, key, rep)]
                assert np.array_equal(first.indices, other.indices) and np.array_equal(first.data, other.data)
        ref = po.omp(A, Bm[:, 3], k, eps=0.0)
        col = results[(t, "own", 0)][:, 3]
        assert col.indices.tolist() == ref.nzind and np.allclose(col.data, ref.nzval, rtol=1e-10, atol=1e-12)
        sh = results[(t, "shared", 5)]
        assert np.array_equal(sh.indices, results[(0, "shared", 0)].indices) and np.array_equal(sh.data, results[(0, "shared", 0)].data)
    for D in own + [shared]:
        D.close()


def test_multi_device_handle_fans_out_bit_identically(cs, po):
    """csb200_dict_create_multi (SURVEY 8b "Threading": fan-out is library-internal): ONE csb200_omp / gomp / mp call
    on a handle with several workers -- all visible GPUs, and at least three workers (entries may repeat, which is
    how the driver's single-GPU box exercises the fan-out) -- must equal the single-device call bit for bit."""
    rng = np.random.default_rng(404)
    M, N, k, B = 96, 1024, 6, 1500
    A = po.gaussian_dictionary(rng, M, N)
    X0, Bm = _planted(po, rng, A, k, B, noise=1e-3)
    ndev = cs.device_count()
    devices = list(range(ndev)) if ndev >= 3 else [i % ndev for i in range(3)]
    with cs.Dictionary(A) as D1, cs.Dictionary(A, devices=devices) as Dn:
        assert Dn.devices == devices
        for call in (lambda D: cs.omp(D, Bm, k), lambda D: cs.gomp(D, Bm, 2, k), lambda D: cs.mp(D, Bm, 2 * k),
                     lambda D: cs.fr(D, Bm, 0.0, 0.0, k), lambda D: cs.sp(D, Bm, k), lambda D: cs.oblivious(D, Bm, k)):
            one, many = call(D1), call(Dn)
            assert len(one) == len(many) == B
            for s in range(B):
                assert np.array_equal(one[s].nzind, many[s].nzind), s
                assert np.array_equal(one[s].nzval, many[s].nzval), s
        x = cs.omp(Dn, Bm[:, :10], k)                        # below the fan-out threshold: one worker
        assert all(np.array_equal(x[s].nzind, cs.omp(D1, Bm[:, s], k).nzind) for s in range(10))
    bad = np.array(Bm, copy=True)
    bad[3, B - 1] = np.nan                                   # the error of one worker reaches the caller
    with cs.Dictionary(A, devices=devices) as Dn:
        with pytest.raises(cs.CSB200Error) as ei:
            cs.omp(Dn, bad, k)
        assert ei.value.status == -3


@pytest.mark.parametrize("parts,ctas_per_sm", [(2, 0), (3, 0), (3, 1)])
def test_two_half_overlap_is_bit_identical(cs, po, monkeypatch, parts, ctas_per_sm):
    """Large omp batches run as two halves (or CSB200_SPLIT_PARTS parts) on a high- and a low-priority stream so that the
    update of one part overlaps the correlation pass of the next (api.cu run_omp_split): same kernels on the same data,
    so every output must equal the plain loop (CSB200_SPLIT=0) bit for bit -- uneven parts, ragged last tile, eps-breaks
    included -- and the oracle on a sample.  Also with the update as a fixed grid of one CTA per SM walking the signals
    (CSB200_SPLIT_CTAS_PER_SM=1)."""
    monkeypatch.setenv("CSB200_SPLIT_PARTS", str(parts))
    monkeypatch.setenv("CSB200_SPLIT_CTAS_PER_SM", str(ctas_per_sm))
    monkeypatch.setenv("CSB200_SCREEN", "0")                 # this test is about the FP64 DMMA path (the default would screen)
    rng = np.random.default_rng(515)
    M, N, k, B = 100, 300, 5, 8192 + 77
    A = po.gaussian_dictionary(rng, M, N)
    idx = np.stack([rng.choice(N, size=k, replace=False) for _ in range(B)])
    sign = rng.choice(np.array([-1.0, 1.0]), size=(B, k))
    Bm = np.asfortranarray(np.einsum("msk,sk->ms", A[:, idx], sign))
    Bm[:, 5::7] = A[:, idx[5::7, 0]] * 2.0                   # 1-sparse signals: eps-break after the first update!
    out = {}
    with cs.Dictionary(A) as D:
        for mode in ("0", "1"):
            monkeypatch.setenv("CSB200_SPLIT", mode)
            with cs.Batch(D, B, k) as batch:
                batch.upload(Bm)
                batch.profile(True)
                batch.omp(k, 1e-9)
                ms, launches, other = batch.corr_time()
                batch.profile(False)
                assert launches == (parts * k if mode == "1" else k) and ms > 0
                out[mode] = batch.download(k) + (batch.residual(),)
    for a, b in zip(out["0"], out["1"]):
        assert np.array_equal(a, b)
    sel, coef, nnz, res, its, R = out["1"]
    assert (its[5::7] == 1).all() and (nnz[5::7] == 1).all()
    for s in (0, 5, 4100, 4200, B - 1):
        t = po.Trace()
        ref = po.omp(A, Bm[:, s], k, eps=1e-9, trace=t)
        n = int(nnz[s])
        assert sel[s, :n].tolist() == t.order() and int(its[s]) == t.iterations
        i2, v2 = _sorted(sel[s], coef[s], n)
        assert i2.tolist() == ref.nzind and _close(v2, ref.nzval, RTOL64)


def test_pipelined_one_shot_matches_single_upload(cs, po, monkeypatch):
    """Host batches of >= 32 768 signals are cut into whole-wave chunks whose uploads / downloads overlap the solves
    (csb200_omp / _gomp / _fr one-shot calls).  Results must be bit-identical to the single-upload path, ragged last
    chunk included, and a NaN anywhere in the batch must still be reported."""
    rng = np.random.default_rng(5)
    M, N, k, B = 64, 300, 4, 40001
    A = po.gaussian_dictionary(rng, M, N)
    idx = rng.integers(0, N, size=(B, k))
    Bm = np.asfortranarray((A[:, idx] * rng.choice([-1.0, 1.0], size=(1, B, k))).sum(axis=2) + 1e-3 * rng.standard_normal((M, B)))
    with cs.Dictionary(A) as D:
        out = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("CSB200_PIPELINE", mode)
            out[mode] = (cs.omp(D, Bm, 0.0, k, result="csc"), cs.gomp(D, Bm, 2, 0.0, k, result="csc"),
                         cs.fr(D, Bm, 0.0, 0.0, k, result="csc"))
        for a, b in zip(out["1"], out["0"]):
            assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
            assert np.array_equal(a.data, b.data)
        for s in (0, 18943, 18944, 40000):
            ref = po.omp(A, Bm[:, s], k, eps=0.0)
            col = out["1"][0][:, s]
            assert col.indices.tolist() == ref.nzind and np.allclose(col.data, ref.nzval, rtol=1e-10, atol=1e-12)
        monkeypatch.setenv("CSB200_PIPELINE", "1")
        Bm[5, 30000] = np.nan
        with pytest.raises(cs.CSB200Error) as ei:
            cs.omp(D, Bm, 0.0, k)
        assert ei.value.status == -3


# ------------------------------------------------------------------ forward regression / OLS
def test_fr_call_surface(cs, po):
    """`fr` / `ols` / `oomp` / `ormp` (src/forward.jl:33-54) through the one-shot C entry point, against the oracle:
    positional and keyword forms, both stopping rules, zero signal, single signal and a small batch."""
    rng = np.random.default_rng(21)
    A, x0, b = po.sparse_data(rng, 40, 90, 4)
    A = np.asfortranarray(A * rng.uniform(0.5, 2.0, size=(1, 90)))
    b = A[:, x0.nzind] @ np.asarray(x0.nzval)
    y = po.perturb(rng, b, 1e-2)
    with cs.Dictionary(A) as D:
        for args, kw, oargs in [((), dict(sparsity=4), (0.0, 0.0, 4)), ((0.05, 0.0), {}, (0.05, 0.0, None)),
                                ((0.0, 0.05), {}, (0.0, 0.05, None)), ((0.0, 0.0, 7), {}, (0.0, 0.0, 7)),
                                ((), dict(max_residual=0.05), (0.05, 0.0, 90))]:
            got = cs.fr(D, y, *args, **kw)
            ref = po.fr(A, y, *oargs)
            assert got.nzind.tolist() == ref.nzind, (args, kw)
            assert _close(got.nzval, ref.nzval, RTOL64), (args, kw)
        assert cs.ols is cs.fr and cs.oomp is cs.fr and cs.ormp is cs.fr
        assert cs.fr(D, np.zeros(40)).nnz() == 0
        Bm = np.asfortranarray(np.stack([y, b, 2 * y - b], axis=1))
        out = cs.fr(D, Bm, sparsity=4)
        for s in range(3):
            ref = po.fr(A, Bm[:, s], 0.0, 0.0, 4)
            assert out[s].nzind.tolist() == ref.nzind and _close(out[s].nzval, ref.nzval, RTOL64)
    got = cs.fr(A.astype(np.float32), y.astype(np.float32), sparsity=4)      # FP32 input: solved on the FP64 twin
    ref = po.fr(A.astype(np.float32).astype(np.float64), y.astype(np.float32).astype(np.float64), 0.0, 0.0, 4)
    assert got.nzind.tolist() == ref.nzind and _close(got.nzval, ref.nzval, 1e-9)


def test_fr_midsize_batch_vs_oracle(cs, po):
    """Ragged mid-size batch, un-normalised atoms, noisy signals: selection sequence bit-exact on a sample,
    coefficients / residual norms within 1e-10, planted support recovered for every signal."""
    rng = np.random.default_rng(77)
    M, N, k, B = 200, 1111, 12, 150
    A = po.gaussian_dictionary(rng, M, N)
    A = np.asfortranarray(A * rng.uniform(0.5, 2.0, size=(1, N)))
    X0, Bm = _planted(po, rng, A, k, B, noise=5e-3)
    with cs.Dictionary(A) as D, cs.Batch(D, B, k) as batch:
        batch.upload(Bm)
        batch.fr(k)
        sel, coef, nnz, res, its = batch.download(k)
        batch.fr(k)                                                     # bit-identical re-solve on the same batch
        sel2, coef2, *_ = batch.download(k)
    assert np.array_equal(sel, sel2) and np.array_equal(coef, coef2)
    for s in range(B):
        assert sorted(sel[s, :k].tolist()) == X0[s].nzind, s
    for s in range(0, B, 10):
        t = po.Trace()
        ref = po.fr(A, Bm[:, s], 0.0, 0.0, k, trace=t)
        assert sel[s, :k].tolist() == t.order(), (s, min(t.margin))
        idx, val = _sorted(sel[s], coef[s], k)
        assert _close(val, ref.nzval, RTOL64) and abs(res[s] - t.resnorm[-1]) < 1e-10


# ------------------------------------------------------------------ subspace pursuit / oblivious
def test_sp_and_oblivious_call_surface(cs, po):
    rng = np.random.default_rng(31)
    A, x0, b = po.sparse_data(rng, 64, 256, 16)
    y = po.perturb(rng, b, 1e-2)
    with cs.Dictionary(A) as D:
        for args in [(16,), (16, 1e-2), (16, 1e-12, 1), (16, 1e-12, 0), (5,)]:
            got = cs.sp(D, y, *args)
            ref = po.sp(A, y, *args)
            assert got.nzind.tolist() == ref.nzind, args
            assert _close(got.nzval, ref.nzval, RTOL64), args
        got, ref = cs.oblivious(D, y, 7), po.oblivious(A, y, 7)
        assert got.nzind.tolist() == ref.nzind and _close(got.nzval, ref.nzval, RTOL64)
        with pytest.raises(ValueError, match="invalid for Subspace Pursuit"):
            cs.sp(D, y, 33)
        Bm = np.asfortranarray(np.stack([y, b, 0.5 * (y + b)], axis=1))
        out = cs.sp(D, Bm, 16)
        for s in range(3):
            ref = po.sp(A, Bm[:, s], 16)
            assert out[s].nzind.tolist() == ref.nzind and _close(out[s].nzval, ref.nzval, 1e-9)


@pytest.mark.parametrize("k", [128, 256])
def test_few_signal_topk_with_clustered_correlations(cs, po, k):
    """One signal on a long dictionary takes the GEMV correlation pass, whose candidate blocks are CTA ranges of up to
    2048 atoms -- not 64-atom blocks.  All of the global top-k sit in ONE range here (a cluster of adjacent atoms that
    correlate with b): every one of them must come back (`partialsortperm(abs.(A'b), 1:k, rev=true)`,
    src/oblivious.jl:4).  [Round-1 code clamped the per-range candidate count to 64 and silently lost the rest.]"""
    rng = np.random.default_rng(4242 + k)
    M, N = 288, 262144
    A = rng.standard_normal((M, N))
    u = rng.standard_normal(M)
    lo = 70000
    A[:, lo:lo + k + 40] = u[:, None] + 0.6 * A[:, lo:lo + k + 40]            # k + 40 adjacent atoms correlate with u
    A /= np.sqrt((A * A).sum(axis=0, keepdims=True))
    A = np.asfortranarray(A)
    b = u / np.linalg.norm(u)
    got, ref = cs.oblivious(A, b, k), po.oblivious(A, b, k)
    assert lo <= min(ref.nzind) and max(ref.nzind) < lo + k + 40               # the premise: one cluster holds them all
    assert got.nzind.tolist() == ref.nzind
    assert _close(got.nzval, ref.nzval, 1e-8)
    if 2 * k <= M:
        got, ref = cs.sp(A, b, k, 1e-12, 2), po.sp(A, b, k, 1e-12, 2)
        assert got.nzind.tolist() == ref.nzind and _close(got.nzval, ref.nzval, 1e-8)


@pytest.mark.parametrize("nsig", [2, 30])
def test_sp_oblivious_cumbabel_beyond_256_atoms(cs, po, nsig):
    """`sp(A, b, k)`, `oblivious(A, b, k)` and `cumbabel(A, k)` take any k in the reference (src/twostage.jl:105,
    src/oblivious.jl:4, src/util.jl:106); round 1 stopped at 256 / 255.  k = 300 on two signals (GEMV pass, k candidates per
    CTA range) and on a batch (dense |A'r| + radix select), against the oracle."""
    rng = np.random.default_rng(6000 + nsig)
    M, N, k = 640, 2000, 300
    A = po.gaussian_dictionary(rng, M, N)
    X0, Bm = _planted(po, rng, A, k, nsig, noise=1e-2)
    with cs.Dictionary(A) as D:
        xs = cs.sp(D, Bm, k, 1e-12, 3)
        xo = cs.oblivious(D, Bm, k)
        mu = cs.cumbabel(D, k) if nsig == 2 else None
    for s in range(min(nsig, 2)):
        t = po.Trace()
        ref = po.sp(A, Bm[:, s], k, 1e-12, 3, trace=t)
        if min(t.margin) > 1e-8:
            assert xs[s].nzind.tolist() == ref.nzind, (s, min(t.margin))
            assert _close(xs[s].nzval, ref.nzval, 1e-8)
        ref = po.oblivious(A, Bm[:, s], k)
        assert xo[s].nzind.tolist() == ref.nzind and _close(xo[s].nzval, ref.nzval, 1e-8)
    if mu is not None:
        assert np.allclose(mu, po.cumbabel(A, k), rtol=1e-12)


@pytest.mark.parametrize("gram", ["0", "1"])
def test_sp_midsize_batch_vs_oracle(cs, po, gram, monkeypatch):
    monkeypatch.setenv("CSB200_GRAM", gram)
    _sp_midsize(cs, po)


def _sp_midsize(cs, po):
    """Mid-size batch (DMMA correlation path, block appends, k > 64-atom candidate blocks exercised via N ragged):
    per-signal iteration counts, supports and coefficients against the oracle; bit-identical re-solve."""
    rng = np.random.default_rng(177)
    M, N, k, B = 160, 1000, 30, 96
    A = po.gaussian_dictionary(rng, M, N)
    X0, Bm = _planted(po, rng, A, k, B, noise=1e-2)
    with cs.Dictionary(A) as D, cs.Batch(D, B, 2 * k) as batch:
        batch.upload(Bm)
        batch.sp(k)
        sel, coef, nnz, res, its = batch.download(k)
        batch.sp(k)
        sel2, coef2, *_ = batch.download(k)
        batch.oblivious(k)
        osel, ocoef, onnz, ores, _ = batch.download(k)
    assert np.array_equal(sel, sel2) and np.array_equal(coef, coef2)
    assert (nnz == k).all() and len(set(its.tolist())) > 1            # signals stop after different numbers of update!s
    for s in range(0, B, 6):
        t = po.Trace()
        ref = po.sp(A, Bm[:, s], k, trace=t)
        if min(t.margin) < 1e-7:
            continue
        idx, val = _sorted(sel[s], coef[s], k)
        assert idx.tolist() == ref.nzind, (s, min(t.margin))
        assert _close(val, ref.nzval, 1e-9) and abs(res[s] - t.resnorm[-1]) < 1e-9
        assert int(its[s]) == t.iterations + 1
        ref = po.oblivious(A, Bm[:, s], k)
        idx, val = _sorted(osel[s], ocoef[s], k)
        assert idx.tolist() == ref.nzind and _close(val, ref.nzval, 1e-9)


# ------------------------------------------------------------------ dictionary analysis
@pytest.mark.parametrize("M,N,k,dtype", [(64, 128, 16, np.float64), (70, 333, 40, np.float64), (32, 20, 19, np.float64),
                                         (64, 160, 8, np.float32), (100, 1500, 70, np.float64)])
def test_cumbabel_coherence_colnorms(cs, po, M, N, k, dtype):
    """`cumbabel` / `babel` / `coherence` / `colnorms` / `normalize!` against the oracle (test/util.jl:7-20 properties)."""
    rng = np.random.default_rng(N + k)
    A = po.gaussian_dictionary(rng, M, N, dtype)
    if N == 333:
        A = np.asfortranarray(A * rng.uniform(0.5, 2.0, size=(1, N)))      # un-normalised: self product is not the maximum
    rt = 1e-12 if dtype == np.float64 else 1e-5
    with cs.Dictionary(A) as D:
        mu1 = cs.cumbabel(D, k)
        assert mu1.dtype == dtype
        assert np.allclose(mu1, po.cumbabel(A, k), rtol=rt)
        mu = cs.coherence(D)
        assert np.isclose(mu, po.coherence(A), rtol=rt) and np.isclose(cs.babel(D, 1), mu)
        assert np.isclose(cs.babel(D, k), mu1[k - 1], rtol=rt)
        assert all(mu1[i] <= (i + 1) * mu * (1 + 1e-6) + 1e-12 for i in range(k))
        assert np.allclose(cs.colnorms(D), po.colnorms(A), rtol=1e-6 if dtype == np.float32 else 1e-14)
    B = np.asfortranarray(A * 3.0)
    assert np.allclose(po.colnorms(cs.normalize(B)), 1.0, rtol=1e-6)


# ------------------------------------------------------------------ mid-size parity and properties
def _planted(po, rng, A, k, B, noise=0.0):
    N = A.shape[1]
    X0, cols = [], []
    for _ in range(B):
        x0 = po.sparse_vector(rng, N, k)
        b = A[:, x0.nzind] @ np.asarray(x0.nzval)
        if noise:
            b = po.perturb(rng, b, noise)
        X0.append(x0); cols.append(b)
    return X0, np.asfortranarray(np.stack(cols, axis=1))


def test_c2_shape_reduced_batch_vs_oracle(cs, po):
    """BASELINE config 2 shape (1024 x 8192, k = 32) on 192 signals: first 12 against the oracle, all
    against the planted support (noiseless, k = planted sparsity: every decision has a wide margin)."""
    rng = np.random.default_rng(1234)
    M, N, k, B = 1024, 8192, 32, 192
    A = po.gaussian_dictionary(rng, M, N)
    X0, Bm = _planted(po, rng, A, k, B)
    with cs.Dictionary(A) as D, cs.Batch(D, B, k) as batch:
        batch.upload(Bm)
        batch.omp(k, float(np.finfo(float).eps))
        sel, coef, nnz, res, its = batch.download(k)
    for s in range(B):
        idx, val = _sorted(sel[s], coef[s], int(nnz[s]))
        assert idx.tolist() == X0[s].nzind, s
        assert np.allclose(val, X0[s].nzval, rtol=1e-10, atol=1e-10), s
        assert res[s] < 1e-12
    for s in range(12):
        t = po.Trace()
        ref = po.omp(A, Bm[:, s], k, trace=t)
        assert sel[s, :k].tolist() == t.order(), (s, "selection sequence", min(t.margin))
        idx, val = _sorted(sel[s], coef[s], k)
        assert _close(val, ref.nzval, RTOL64)


@pytest.mark.parametrize("gram", ["0", "1"])
def test_gram_matrix_path_matches_oracle(cs, po, gram, monkeypatch):
    """The batched update can take A_S'a_j from a cached Gram matrix (CSB200_GRAM=1 forces it, =0 forbids it):
    same selection sequence, same coefficients, ragged shape, omp and gomp."""
    monkeypatch.setenv("CSB200_GRAM", gram)
    monkeypatch.setenv("CSB200_SMALL_SOLVE", "0")
    rng = np.random.default_rng(99)
    M, N, k, B = 200, 1111, 14, 70
    A = po.gaussian_dictionary(rng, M, N)
    X0, Bm = _planted(po, rng, A, k, B, noise=5e-3)
    with cs.Dictionary(A) as D, cs.Batch(D, B, k) as batch:
        batch.upload(Bm)
        batch.omp(k, 0.0)
        sel, coef, nnz, res, its = batch.download(k)
        batch.gomp(3, k, 0.0)
        gsel, gcoef, gnnz, gres, gits = batch.download(k)
    for s in range(0, B, 5):
        t = po.Trace()
        ref = po.omp(A, Bm[:, s], k, eps=0.0, trace=t)
        assert sel[s, :k].tolist() == t.order(), s
        idx, val = _sorted(sel[s], coef[s], k)
        assert _close(val, ref.nzval, RTOL64) and abs(res[s] - t.resnorm[-1]) < 1e-10
        t = po.Trace()
        ref = po.gomp(A, Bm[:, s], 3, k, eps=0.0, trace=t)
        assert gsel[s, :int(gnnz[s])].tolist() == t.order(), s
        idx, val = _sorted(gsel[s], gcoef[s], int(gnnz[s]))
        assert idx.tolist() == ref.nzind and _close(val, ref.nzval, RTOL64)


@pytest.mark.parametrize("block", ["0", "1"])
def test_gomp_block_append_and_near_dependent_atoms(cs, po, block, monkeypatch):
    """gomp appends its l atoms as one block (two gather sweeps per update instead of 2 l); a nearly dependent
    atom inside a block must fall back to the explicit, re-orthogonalised append.  CSB200_GOMP_BLOCK=0 forces the
    one-by-one path: both must agree with the oracle."""
    monkeypatch.setenv("CSB200_GOMP_BLOCK", block)
    monkeypatch.setenv("CSB200_SMALL_SOLVE", "0")
    rng = np.random.default_rng(123)
    M, N, k, l, B = 300, 2500, 20, 5, 60
    A = po.gaussian_dictionary(rng, M, N)
    twin = A[:, 100] + 2e-3 * rng.standard_normal(M)              # atom 101 ~ atom 100: cond(A_S) ~ 1e3
    A[:, 101] = twin / np.linalg.norm(twin)
    X0, Bm = _planted(po, rng, A, k, B, noise=5e-3)
    Bm[:, 0] = 3.0 * A[:, 100] + 2.9 * A[:, 101] + 0.5 * A[:, 7]    # both twins in the first top-l
    with cs.Dictionary(A) as D, cs.Batch(D, B, k) as batch:
        batch.upload(Bm)
        batch.gomp(l, k, 0.0)
        sel, coef, nnz, res, its = batch.download(k)
        fl = batch.flags()
    assert fl[0] & 16                                  # the twin support is flagged ill-conditioned (and was refined)
    for s in range(0, B, 3):
        t = po.Trace()
        ref = po.gomp(A, Bm[:, s], l, k, eps=0.0, trace=t)
        n = int(nnz[s])
        if s == 0:
            # noise-free 3-atom signal: exactly recovered after two update!s; from then on the residual is rounding noise
            # (1e-15) and WHICH atoms the remaining update!s add is decided by that noise (SURVEY 7 "bit-exact support vs a
            # different summation order").  Pinned: the well-posed prefix of the selection sequence, the coefficients of
            # those atoms -- cond(A_S) ~ 1e3 here; after the iterative refinement they meet the 1e-10 bar (round 1 needed
            # 1e-6: x = R^{-1} z through the stored inverse alone is cond^2 eps accurate) -- and that every later atom
            # carries a rounding-level coefficient.
            assert len(t.order()) >= 10 and sel[s, :10].tolist() == t.order()[:10]
            lut = dict(zip(sel[s, :n].tolist(), coef[s, :n].tolist()))
            want = dict(zip(ref.nzind, ref.nzval))
            scale = max(abs(v) for v in want.values())
            assert all(abs(lut[j] - want[j]) <= RTOL64 * scale for j in (100, 101, 7))
            assert all(abs(v) < 1e-9 * scale for j, v in lut.items() if j not in (100, 101, 7))
            assert res[s] < 1e-12 and t.resnorm[-1] < 1e-12
            continue
        assert sel[s, :n].tolist() == t.order(), (s, sel[s, :n], t.order())
        idx, val = _sorted(sel[s], coef[s], n)
        assert idx.tolist() == ref.nzind
        assert _close(val, ref.nzval, RTOL64), (s, val, ref.nzval)
        assert abs(res[s] - t.resnorm[-1]) < 1e-9


@pytest.mark.parametrize("nsig", [40, 2])
def test_gomp_many_atoms_per_update(cs, po, nsig, monkeypatch):
    """`gomp(A, b, l, k)` takes any l (src/matchingpursuit.jl:189-193: `partialsortperm(P.Ar, 1:k, rev=true)`); round 1
    stopped at l = 64.  l = 128 on a batch (dense |A'r| + radix select in the block-append kernel) and on two signals
    (GEMV pass with l candidates per CTA range + cluster update) against the oracle; l = 257 is refused cleanly."""
    monkeypatch.setenv("CSB200_SMALL_SOLVE", "0")
    rng = np.random.default_rng(900 + nsig)
    M, N, l, k = 384, 3000, 128, 256
    A = po.gaussian_dictionary(rng, M, N)
    X0, Bm = _planted(po, rng, A, 60, nsig, noise=1e-2)
    with cs.Dictionary(A) as D, cs.Batch(D, nsig, k) as batch:
        batch.upload(Bm)
        batch.gomp(l, k, 0.0)
        sel, coef, nnz, res, its = batch.download(k)
        with pytest.raises(cs.CSB200Error) as ei:
            batch.gomp(257, 257, 0.0)
        assert ei.value.status == -7
    for s in range(min(nsig, 4)):
        t = po.Trace()
        ref = po.gomp(A, Bm[:, s], l, k, eps=0.0, trace=t)
        n = int(nnz[s])
        assert n == ref.nnz() == k and int(its[s]) == t.iterations
        assert sel[s, :n].tolist() == t.order(), (s, "selection sequence", min(t.margin))
        idx, val = _sorted(sel[s], coef[s], n)
        assert idx.tolist() == ref.nzind and _close(val, ref.nzval, 1e-9)
        assert abs(res[s] - t.resnorm[-1]) < 1e-9


def test_gomp_and_mp_midsize_vs_oracle(cs, po):
    rng = np.random.default_rng(77)
    M, N, k, B = 256, 2048, 16, 64
    A = po.gaussian_dictionary(rng, M, N)
    X0, Bm = _planted(po, rng, A, k, B, noise=5e-3)
    with cs.Dictionary(A) as D:
        xs = cs.gomp(D, Bm, 4, k)
        xm = cs.mp(D, Bm[:, :32], 40)
    for s in range(B):
        ref = po.gomp(A, Bm[:, s], 4, k)
        assert xs[s].nzind.tolist() == ref.nzind, s
        assert _close(xs[s].nzval, ref.nzval, RTOL64), s
    for s in range(8):
        ref = po.mp(A, Bm[:, s], 40)
        assert xm[s].nzind.tolist() == ref.nzind, s
        assert _close(xm[s].nzval, ref.nzval, 1e-9), s


def test_residual_orthogonality_and_idempotence(cs, po):
    """Size-independent properties: r is orthogonal to the active atoms; re-solving the same batch gives
    bit-identical output (deterministic kernels); signals are independent of their batch neighbours."""
    rng = np.random.default_rng(5)
    M, N, k, B = 512, 4096, 24, 300
    A = po.gaussian_dictionary(rng, M, N)
    X0, Bm = _planted(po, rng, A, k, B, noise=1e-2)
    with cs.Dictionary(A) as D, cs.Batch(D, B, k) as batch:
        batch.upload(Bm)
        batch.omp(k, 0.0)
        out1 = batch.download(k)
        R = batch.residual()
        batch.omp(k, 0.0)
        out2 = batch.download(k)
        with cs.Batch(D, 40, k) as small:
            small.upload(Bm[:, 100:140])
            small.omp(k, 0.0)
            out3 = small.download(k)
    for a, b in zip(out1, out2):
        assert np.array_equal(a, b)
    assert np.array_equal(out1[0][100:140], out3[0])
    assert np.array_equal(out1[1][100:140], out3[1])
    for s in range(0, B, 17):
        S = out1[0][s, :k]
        assert np.max(np.abs(A[:, S].T @ R[:, s])) < 1e-12
        assert abs(np.linalg.norm(R[:, s]) - out1[3][s]) < 1e-12


def test_f32_single_signal_omp(cs, po):
    rng = np.random.default_rng(8)
    M, N, k = 2048, 16384, 24
    A = po.gaussian_dictionary(rng, M, N, np.float32)
    x0 = po.sparse_vector(rng, N, k)
    b = (A[:, x0.nzind].astype(np.float64) @ np.asarray(x0.nzval)).astype(np.float32)
    x = cs.omp(A, b, k)
    assert x.nzind.tolist() == x0.nzind
    assert np.allclose(x.nzval, x0.nzval, atol=RTOL32)
    assert x.nzval.dtype == np.float64


@pytest.mark.slow
def test_c2_full_size_properties(cs, po):
    """BASELINE config 2 at full size (65 536 signals): planted supports recovered, coefficients +-1."""
    rng = np.random.default_rng(1234)
    M, N, k, B = 1024, 8192, 32, 65536
    A = po.gaussian_dictionary(rng, M, N)
    idx = rng.integers(0, N, size=(B, k))
    while True:                                            # resample rows that drew an atom twice
        srt = np.sort(idx, axis=1)
        bad = np.nonzero((np.diff(srt, axis=1) == 0).any(axis=1))[0]
        if bad.size == 0:
            break
        idx[bad] = rng.integers(0, N, size=(bad.size, k))
    idx = np.sort(idx, axis=1)
    sign = rng.choice(np.array([-1.0, 1.0]), size=(B, k))
    Bm = np.empty((M, B), order="F")
    for s0 in range(0, B, 512):
        blk = slice(s0, min(B, s0 + 512))
        Bm[:, blk] = np.einsum("mbk,bk->mb", A[:, idx[blk]], sign[blk])
    with cs.Dictionary(A) as D, cs.Batch(D, B, k) as batch:
        batch.upload(Bm)
        batch.omp(k, float(np.finfo(float).eps))
        sel, coef, nnz, res, its = batch.download(k)
    assert (nnz == k).all()
    order = np.argsort(sel, axis=1)
    assert np.array_equal(np.take_along_axis(sel, order, axis=1), idx)
    assert np.allclose(np.take_along_axis(coef, order, axis=1), sign, rtol=1e-10, atol=1e-10)
    assert res.max() < 1e-12


@pytest.mark.slow
def test_widened_rows_full_c2_shape_properties(cs, po):
    """fr, sp and oblivious at the config-2 dictionary shape on 16 384 signals (the sizes the oracle cannot reach):
    size-independent properties -- planted supports recovered with coefficients +-1 (fr, sp), the residual orthogonal
    to the active atoms, oblivious = the k largest |A'b|, and the one-shot (pipelined) call agreeing with the batch API."""
    rng = np.random.default_rng(4321)
    M, N, k, B = 1024, 8192, 24, 16384
    A = po.gaussian_dictionary(rng, M, N)
    idx = np.sort(np.stack([rng.choice(N, size=k, replace=False) for _ in range(B)]), axis=1)
    sign = rng.choice(np.array([-1.0, 1.0]), size=(B, k))
    Bm = np.empty((M, B), order="F")
    for s0 in range(0, B, 512):
        blk = slice(s0, min(B, s0 + 512))
        Bm[:, blk] = np.einsum("mbk,bk->mb", A[:, idx[blk]], sign[blk])
    with cs.Dictionary(A) as D, cs.Batch(D, B, 2 * k) as batch:
        batch.upload(Bm)
        for algo in ("fr", "sp"):
            getattr(batch, algo)(k)
            sel, coef, nnz, res, its = batch.download(k)
            assert (nnz == k).all(), algo
            order = np.argsort(sel, axis=1)
            assert np.array_equal(np.take_along_axis(sel, order, axis=1), idx), algo
            assert np.allclose(np.take_along_axis(coef, order, axis=1), sign, rtol=1e-9, atol=1e-9), algo
            assert res.max() < 1e-11, algo
            R = batch.residual()
            for s in range(0, B, 1024):
                assert np.abs(A[:, idx[s]].T @ R[:, s]).max() < 1e-12
        batch.oblivious(k)
        osel, ocoef, onnz, ores, _ = batch.download(k)
        C = np.abs(A.T @ Bm[:, :64])
        for s in range(64):
            assert sorted(osel[s].tolist()) == sorted(np.argsort(-C[:, s], kind="stable")[:k].tolist())
            c, *_ = np.linalg.lstsq(A[:, osel[s]], Bm[:, s], rcond=None)
            assert np.allclose(ocoef[s], c, rtol=1e-9, atol=1e-11)
        X = cs.fr(D, Bm, sparsity=k, result="csc")
        assert np.array_equal(X.indices.reshape(B, k), idx) and np.allclose(X.data.reshape(B, k), sign, rtol=1e-9)


# ------------------------------------------------------------------ CUDA-graph replay of few-signal solves
@pytest.mark.parametrize("algo", ["omp", "gomp", "mp"])
@pytest.mark.parametrize("dtype,nsig", [(np.float64, 1), (np.float32, 3), (np.float64, 23)])
def test_few_signal_solves_replay_a_cuda_graph(cs, po, algo, dtype, nsig, monkeypatch):
    """A solve with < 24 signals is captured into a CUDA graph the second time it runs with the same (algorithm, k, l,
    eps, signal count) on a batch and replayed afterwards.  Every run -- direct, captured, replayed, and direct again
    after the key changes -- must reproduce the oracle on fresh signals."""
    for hook in ("CSB200_UPDATE_IMPL", "CSB200_CLUSTER", "CSB200_GEMV_L2", "CSB200_GRAM", "CSB200_GRAPH", "CSB200_CORR_IMPL"):
        monkeypatch.delenv(hook, raising=False)
    monkeypatch.setenv("CSB200_PERSIST", "0")         # <= 8 signals would otherwise take the cooperative whole-solve kernel
    rng = np.random.default_rng(77 + nsig)
    M, N, k, l = 96, 4500, 6, 2                       # N > 4096: not the whole-solve small-dictionary kernel
    A = po.gaussian_dictionary(rng, M, N, dtype)
    rtol = RTOL32 if dtype == np.float32 else RTOL64
    eps = float(np.finfo(dtype).eps)

    def run_and_check(batch, kk):
        X0, Bm = _planted(po, rng, A.astype(np.float64), kk, nsig, noise=1e-3)
        Bm = np.asfortranarray(Bm.astype(dtype))
        batch.upload(Bm)
        if algo == "omp":
            batch.omp(kk, eps)
        elif algo == "gomp":
            batch.gomp(l, kk, eps)
        else:
            batch.mp(kk)
        sel, coef, nnz, res, its = batch.download(kk)
        for s in range(min(nsig, 4)):
            t = po.Trace()
            if algo == "omp":
                ref = po.omp(A, Bm[:, s], kk, trace=t)
            elif algo == "gomp":
                ref = po.gomp(A, Bm[:, s], l, kk, trace=t)
            else:
                ref = po.mp(A, Bm[:, s], kk, trace=t)
            if algo == "mp":
                acc = {}
                for i, c in zip(sel[s, :kk].tolist(), coef[s, :kk].tolist()):
                    acc[i] = acc.get(i, 0.0) + c
                assert sorted(acc) == ref.nzind
                assert _close(np.array([acc[i] for i in sorted(acc)]), ref.nzval, max(rtol, 1e-9))
            else:
                n = int(nnz[s])
                assert sel[s, :n].tolist() == t.order(), (s, "selection sequence")
                idx, val = _sorted(sel[s], coef[s], n)
                assert idx.tolist() == ref.nzind and _close(val, ref.nzval, rtol)

    with cs.Dictionary(A) as D, cs.Batch(D, nsig, k) as batch:
        for rep in range(4):
            run_and_check(batch, k)
        assert batch.graph_replays() == 3             # rep 0 direct, rep 1 captured + launched, reps 2-3 replayed
        run_and_check(batch, k - 2)                   # another key: direct again
        assert batch.graph_replays() == 3
        run_and_check(batch, k - 2)                   # ... captured
        run_and_check(batch, k)                       # the first key is seen anew (one graph is kept per batch)
        assert batch.graph_replays() == 4
    monkeypatch.setenv("CSB200_UPDATE_IMPL", "cluster")   # a test hook in the environment switches replay off
    with cs.Dictionary(A) as D, cs.Batch(D, nsig, k) as batch:
        for rep in range(3):
            run_and_check(batch, k)
        assert batch.graph_replays() == 0


# ------------------------------------------------------------------ cooperative whole-solve kernel (solve_persist.cu)
@pytest.mark.parametrize("algo", ["omp", "mp"])
@pytest.mark.parametrize("dtype,M,N,k,nsig", [(np.float64, 1024, 8192, 32, 1), (np.float32, 1024, 8192, 32, 1),
                                              (np.float64, 1024, 8192, 16, 5), (np.float64, 1024, 8192, 12, 8),
                                              (np.float32, 520, 3001, 9, 3), (np.float64, 70, 130, 6, 2),
                                              (np.float64, 128, 256, 8, 1), (np.float64, 4096, 1500, 20, 1)])
def test_persistent_whole_solve_matches_multi_launch_and_oracle(cs, po, algo, dtype, M, N, k, nsig, monkeypatch):
    """solve_persist.cu (1..8 signals, one cooperative launch per solve: workers with a dictionary slice resident in
    shared memory + one updater CTA per signal) against (a) the per-iteration kernels -- identical selection sequence,
    values to rounding -- and (b) the oracle at the usual bar.  Shapes: the config-2 dictionary (L2 regime, columns
    partly streamed), ragged sizes, config 1, a long-atom case; noisy signals keep every arg-max well posed."""
    rng = np.random.default_rng(1000 + M + nsig)
    A = po.gaussian_dictionary(rng, M, N, dtype)
    X0, Bm = _planted(po, rng, A.astype(np.float64), k, nsig, noise=1e-3)
    Bm = np.asfortranarray(Bm.astype(dtype))
    eps = float(np.finfo(dtype).eps)
    rtol = RTOL32 if dtype == np.float32 else RTOL64
    iters = k if algo == "omp" else 2 * k
    out = {}
    with cs.Dictionary(A) as D:
        for path, env in (("persist", {"CSB200_PERSIST": "1", "CSB200_CLUSTER_SOLVE": "0"}),
                          ("multi", {"CSB200_PERSIST": "0", "CSB200_SMALL_SOLVE": "0", "CSB200_CLUSTER_SOLVE": "0"})):
            for key in ("CSB200_PERSIST", "CSB200_SMALL_SOLVE", "CSB200_CLUSTER_SOLVE"):
                monkeypatch.delenv(key, raising=False)
            for key, val in env.items():
                monkeypatch.setenv(key, val)
            with cs.Batch(D, nsig, iters) as batch:
                for rep in range(2):                                   # the second solve reuses scratch and sync words
                    batch.upload(Bm)
                    batch.omp(iters, eps) if algo == "omp" else batch.mp(iters)
                    out[path, rep] = batch.download(iters) + (batch.residual(),)
        assert all(np.array_equal(out["persist", 0][i], out["persist", 1][i]) for i in range(6))    # deterministic
        sel, coef, nnz, res, its, R = out["persist", 0]
        msel, mcoef, mnnz, mres, mits, mR = out["multi", 0]
        assert np.array_equal(sel, msel) and np.array_equal(nnz, mnnz) and np.array_equal(its, mits)
        scale = np.abs(mcoef).max()
        assert np.max(np.abs(coef - mcoef)) <= (1e-5 if dtype == np.float32 else 1e-12) * scale
        assert np.allclose(res, mres, rtol=1e-4 if dtype == np.float32 else 1e-9, atol=1e-12)
        assert np.allclose(R, mR, rtol=0, atol=(1e-5 if dtype == np.float32 else 1e-12) * np.abs(Bm).max())
    for s in range(min(nsig, 3)):
        t = po.Trace()
        if algo == "omp":
            ref = po.omp(A, Bm[:, s], iters, trace=t)
            n = int(nnz[s])
            assert sel[s, :n].tolist() == t.order(), (s, "selection sequence", min(t.margin))
            idx, val = _sorted(sel[s], coef[s], n)
            assert idx.tolist() == ref.nzind and _close(val, ref.nzval, rtol)
            assert int(its[s]) == t.iterations and abs(res[s] - t.resnorm[-1]) <= rtol * np.linalg.norm(Bm[:, s])
        else:
            ref = po.mp(A, Bm[:, s], iters, trace=t)
            assert sel[s, :iters].tolist() == t.order(), (s, "selection sequence", min(t.margin))
            acc = {}
            for i, c in zip(sel[s, :iters].tolist(), coef[s, :iters].tolist()):
                acc[i] = acc.get(i, 0.0) + c
            assert sorted(acc) == ref.nzind
            assert _close(np.array([acc[i] for i in sorted(acc)]), ref.nzval, max(rtol, 1e-9))


@pytest.mark.parametrize("algo", ["omp", "gomp", "mp"])
@pytest.mark.parametrize("dtype,M,N,k,nsig", [(np.float64, 128, 256, 8, 1), (np.float64, 70, 130, 6, 3), (np.float32, 64, 160, 5, 8),
                                              (np.float64, 96, 1500, 12, 2), (np.float32, 256, 600, 20, 1)])
def test_cluster_resident_whole_solve_matches_multi_launch_and_oracle(cs, po, algo, dtype, M, N, k, nsig, monkeypatch):
    """solve_small.cu cluster_solve_kernel (<= 8 signals; the dictionary lives in the shared memory of an 8-CTA cluster
    per signal, candidates cross by distributed shared memory, every CTA runs the same update) against the
    per-iteration kernels -- identical selection sequence, values to rounding -- and the oracle.  Includes config 1."""
    rng = np.random.default_rng(3000 + M + nsig)
    A = po.gaussian_dictionary(rng, M, N, dtype)
    X0, Bm = _planted(po, rng, A.astype(np.float64), k, nsig, noise=1e-3)
    Bm = np.asfortranarray(Bm.astype(dtype))
    eps = float(np.finfo(dtype).eps)
    rtol = RTOL32 if dtype == np.float32 else RTOL64
    iters, l = (k, 3) if algo != "mp" else (2 * k, 1)
    out = {}
    with cs.Dictionary(A) as D:
        for path, env in (("cluster", {"CSB200_CLUSTER_SOLVE": "1"}),
                          ("multi", {"CSB200_CLUSTER_SOLVE": "0", "CSB200_PERSIST": "0", "CSB200_SMALL_SOLVE": "0"})):
            for key in ("CSB200_CLUSTER_SOLVE", "CSB200_PERSIST", "CSB200_SMALL_SOLVE"):
                monkeypatch.delenv(key, raising=False)
            for key, val in env.items():
                monkeypatch.setenv(key, val)
            with cs.Batch(D, nsig, iters) as batch:
                for rep in range(2):
                    batch.upload(Bm)
                    {"omp": lambda: batch.omp(iters, eps), "gomp": lambda: batch.gomp(l, iters, eps), "mp": lambda: batch.mp(iters)}[algo]()
                    out[path, rep] = batch.download(iters) + (batch.residual(),)
    assert all(np.array_equal(out["cluster", 0][i], out["cluster", 1][i]) for i in range(6))
    sel, coef, nnz, res, its, R = out["cluster", 0]
    msel, mcoef, mnnz, mres, mits, mR = out["multi", 0]
    assert np.array_equal(sel, msel) and np.array_equal(nnz, mnnz) and np.array_equal(its, mits)
    scale = np.abs(mcoef).max()
    assert np.max(np.abs(coef - mcoef)) <= (1e-5 if dtype == np.float32 else 1e-12) * scale
    assert np.allclose(R, mR, rtol=0, atol=(1e-5 if dtype == np.float32 else 1e-12) * np.abs(Bm).max())
    for s in range(min(nsig, 3)):
        t = po.Trace()
        if algo == "mp":
            ref = po.mp(A, Bm[:, s], iters, trace=t)
            assert sel[s, :iters].tolist() == t.order(), (s, "selection sequence", min(t.margin))
            acc = {}
            for i, c in zip(sel[s, :iters].tolist(), coef[s, :iters].tolist()):
                acc[i] = acc.get(i, 0.0) + c
            assert sorted(acc) == ref.nzind
            assert _close(np.array([acc[i] for i in sorted(acc)]), ref.nzval, max(rtol, 1e-9))
        else:
            ref = po.omp(A, Bm[:, s], iters, trace=t) if algo == "omp" else po.gomp(A, Bm[:, s], l, iters, trace=t)
            n = int(nnz[s])
            assert sel[s, :n].tolist() == t.order(), (s, "selection sequence", min(t.margin))
            idx, val = _sorted(sel[s], coef[s], n)
            assert idx.tolist() == ref.nzind and _close(val, ref.nzval, rtol)
            assert int(its[s]) == t.iterations


def test_persistent_whole_solve_eps_break_mixed_signals(cs, po, monkeypatch):
    """Signals of one call stop at different update!s (eps-break): a stopped signal's updater leaves, the workers keep
    serving the others; iteration counts and supports per signal must match the oracle."""
    monkeypatch.setenv("CSB200_PERSIST", "1")
    monkeypatch.setenv("CSB200_CLUSTER_SOLVE", "0")
    rng = np.random.default_rng(2024)
    M, N = 200, 900
    A = po.gaussian_dictionary(rng, M, N)
    cols = []
    for kk in (2, 9, 5, 1):                                            # planted sparsities: the eps-break comes at kk
        x0 = po.sparse_vector(rng, N, kk)
        cols.append(A[:, x0.nzind] @ np.asarray(x0.nzval))
    Bm = np.asfortranarray(np.stack(cols, axis=1))
    with cs.Dictionary(A) as D, cs.Batch(D, 4, 12) as batch:
        batch.upload(Bm)
        batch.omp(12, 1e-9)
        sel, coef, nnz, res, its = batch.download(12)
    for s in range(4):
        t = po.Trace()
        ref = po.omp(A, Bm[:, s], 12, eps=1e-9, trace=t)
        n = int(nnz[s])
        assert int(its[s]) == t.iterations and sel[s, :n].tolist() == t.order(), s
        idx, val = _sorted(sel[s], coef[s], n)
        assert idx.tolist() == ref.nzind and _close(val, ref.nzval, RTOL64)


# ------------------------------------------------------------------ column-sharded mode
@pytest.mark.parametrize("exchange", ["peer-memory", "nccl"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sharded_single_rank_matches_oracle(cs, po, dtype, exchange, monkeypatch):
    """The whole sharded code path (GEMV -> exchange [mailbox kernel | local best + NCCL all-gather + global pick] ->
    cached-atom update) with a communicator of one rank: must equal the oracle and the unsharded GPU path."""
    if exchange == "nccl":
        monkeypatch.setenv("CSB200_SHARD_EXCHANGE", "nccl")
    else:
        monkeypatch.delenv("CSB200_SHARD_EXCHANGE", raising=False)
    rng = np.random.default_rng(21)
    M, N, k = 256, 3000, 12
    A = po.gaussian_dictionary(rng, M, N, dtype)
    x0 = po.sparse_vector(rng, N, k)
    b = (A[:, x0.nzind].astype(np.float64) @ np.asarray(x0.nzval)).astype(dtype)
    b = po.perturb(rng, b, 5e-3)
    comm = cs.ShardComm(cs.ShardComm.unique_id(), 0, 1, 0)
    try:
        with cs.Dictionary(A, n_offset=0, n_total=N) as shard:
            x, info = cs.omp_sharded(shard, comm, b, k)
            x1 = cs.omp(shard, b, k)
    finally:
        comm.close()
    assert info["exchange"] == exchange
    t = po.Trace()
    ref = po.omp(A, b, k, trace=t)
    assert info["order"].tolist() == t.order()
    assert x.nzind.tolist() == ref.nzind == x1.nzind.tolist()
    rtol = RTOL32 if dtype == np.float32 else RTOL64
    assert _close(x.nzval, ref.nzval, rtol) and _close(x.nzval, x1.nzval, 1e-12)
    assert info["iters"] == k and abs(info["resnorm"] - t.resnorm[-1]) < 1e-5


def test_sharded_multi_gpu(cs):
    """Two or more ranks over NCCL (runs only where the box has >= 2 GPUs; see tests/run_sharded_gpu.py)."""
    import subprocess
    import sys
    n = cs.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(n, 8)
    script = os.path.join(os.path.dirname(__file__), "run_sharded_gpu.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", script], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def _c4_twin(torch, dev, N=131072, M=8192, k=128, seed=100):
    """SURVEY 8(d)'s scaled-down twin of BASELINE config 4: 8192 x 131072 FP32 (4 GiB), generated on the device by a
    seeded generator and read back, so that the oracle sees the very bytes the GPU path does.  Planted: k atoms, +1."""
    g = torch.Generator(device=dev).manual_seed(seed)
    A_t = torch.empty(N, M, dtype=torch.float32, device=dev)
    for n0 in range(0, N, 16384):
        blk = torch.randn(16384, M, dtype=torch.float32, device=dev, generator=g)
        blk /= blk.norm(dim=1, keepdim=True)
        A_t[n0:n0 + 16384] = blk
    idx = torch.randperm(N, device=dev, generator=g)[:k]
    b = A_t[idx].to(torch.float64).sum(dim=0).to(torch.float32).cpu().numpy()
    A = A_t.cpu().numpy().T                                   # (M, N) Fortran-ordered view
    del A_t
    torch.cuda.empty_cache()
    return A, b, sorted(idx.cpu().tolist())


@pytest.mark.slow
def test_c4_scaled_down_twin_vs_oracle(cs, po):
    """BASELINE config 4's scaled-down twin: single-signal omp, 8192 x 131072 FP32, k = 128, on one GPU
    through the column-sharded entry point (world size 1, HBM-regime GEMV + cluster update) and through plain `omp`,
    against the FP32 oracle on the same bytes: selection sequence exact, coefficients within 2e-5."""
    import torch
    k = 128
    A, b, planted = _c4_twin(torch, torch.device("cuda", 0), k=k)
    comm = cs.ShardComm(cs.ShardComm.unique_id(), 0, 1, 0)
    try:
        with cs.Dictionary(A) as D:
            x, info = cs.omp_sharded(D, comm, b, k)
            x1 = cs.omp(D, b, k)
    finally:
        comm.close()
    t = po.Trace()
    ref = po.omp(A, b, k, trace=t)
    assert info["order"].tolist() == t.order(), ("selection sequence", min(t.margin))
    assert x.nzind.tolist() == ref.nzind == planted == x1.nzind.tolist()
    assert _close(x.nzval, ref.nzval, RTOL32) and _close(x1.nzval, ref.nzval, RTOL32)
    assert np.array_equal(x.nzval, x1.nzval)                 # sharded entry point == unsharded GPU path, bit for bit
    assert info["iters"] == k and abs(info["resnorm"] - t.resnorm[-1]) < 1e-4


# ------------------------------------------------------------------ BASELINE configs 3 and 5 at their full dictionary shapes
@pytest.mark.slow
def test_c3_gomp_full_dictionary_shape(cs, po):
    """BASELINE config 3: gomp (l = 4) on a 2048 x 32768 FP64 dictionary, k = 64; 256 signals on the GPU
    (DMMA GEMM path), the first 3 against the oracle, all against the planted support."""
    rng = np.random.default_rng(33)
    M, N, k, l, B = 2048, 32768, 64, 4, 256
    A = po.gaussian_dictionary(rng, M, N)
    X0, Bm = _planted(po, rng, A, k, B)
    with cs.Dictionary(A) as D, cs.Batch(D, B, k) as batch:
        batch.upload(Bm)
        batch.gomp(l, k, float(np.finfo(float).eps))
        sel, coef, nnz, res, its = batch.download(k)
    for s in range(B):
        idx, val = _sorted(sel[s], coef[s], int(nnz[s]))
        assert idx.tolist() == X0[s].nzind, s
        assert np.allclose(val, X0[s].nzval, rtol=1e-10, atol=1e-10), s
    for s in range(3):
        t = po.Trace()
        ref = po.gomp(A, Bm[:, s], l, k, trace=t)
        assert sel[s, :k].tolist() == t.order(), (s, "selection sequence", min(t.margin))
        idx, val = _sorted(sel[s], coef[s], k)
        assert idx.tolist() == ref.nzind and _close(val, ref.nzval, RTOL64)
        assert int(its[s]) == t.iterations


@pytest.mark.slow
def test_c5_mp_full_dictionary_shape(cs, po):
    """BASELINE config 5: mp on a 4096 x 65536 FP64 dictionary (2 GiB); 64 signals, 24 iterations (the config's 200
    are a matter of time, not of code path), 2 signals against the oracle; residual norms decrease monotonically."""
    rng = np.random.default_rng(55)
    M, N, iters, B = 4096, 65536, 24, 64
    A = po.gaussian_dictionary(rng, M, N)
    X0, Bm = _planted(po, rng, A, 16, B, noise=1e-2)
    with cs.Dictionary(A) as D, cs.Batch(D, B, iters) as batch:
        batch.upload(Bm)
        batch.mp(iters)
        sel, coef, nnz, res, its = batch.download(iters)
        R = batch.residual()
    assert (nnz == iters).all()
    for s in range(2):
        t = po.Trace()
        ref = po.mp(A, Bm[:, s], iters, trace=t)
        assert sel[s, :iters].tolist() == t.order(), (s, min(t.margin))
        acc = {}
        for i, c in zip(sel[s].tolist(), coef[s].tolist()):
            acc[i] = acc.get(i, 0.0) + c
        assert sorted(acc) == ref.nzind
        assert _close([acc[i] for i in sorted(acc)], ref.nzval, 1e-9)
        assert abs(res[s] - t.resnorm[-1]) < 1e-10
    for s in range(B):
        x = np.zeros(N)
        np.add.at(x, sel[s], coef[s])
        assert np.linalg.norm(Bm[:, s] - A @ x - R[:, s]) < 1e-11          # r really is b - A x
        assert res[s] < np.linalg.norm(Bm[:, s])
