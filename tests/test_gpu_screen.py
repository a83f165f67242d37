"""GPU tests (`-m gpu`, B200) of the screening path of large batched omp solves: the tcgen05 screening pass
(corr_screen_tf32.cu; operands in TF32 or in FP16 scaled by powers of two) must respect its proven error bound, its candidate
lists must contain the FP64 arg-max, and the whole solve (screening + exact FP64 re-evaluation, update.cu screen_select /
omp_append_warp_kernel / omp_residual_slice_kernel) must select the same supports as the FP64 DMMA path and the CPU oracle --
bit-exact selection order, coefficients within 1e-10 -- in every combination of pass format and update variant
(/root/reference/src/matchingpursuit.jl:62-70, 181-185)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL64 = 1e-10


def _kappa(M):
    """screen_kappa(M) of csrc/common.cuh."""
    return 1.05 * (2.0 ** -10 + 2.0 ** -22 + M * 2.0 ** -22)


def _kappa_f16(M):
    """Relative bound csb200_debug_screen_pass reports for the FP16 pass: screen_kappa_f16(rows padded to 64) plus the
    additive term of a residual converted at its own norm (< 2e-3 of the relative one)."""
    ld16 = -(-M // 64) * 64
    return (_kappa(ld16) + 1.05 * np.sqrt(ld16) * 2.0 ** -25) * (1.0 + 2e-3)


# every variant of the screened omp loop must give the bits of the FP64 DMMA path: operand format of the pass (TF32 / scaled
# FP16) x update (warp-per-signal append + row-sliced residual sweep / the CTA-per-signal kernel / CTA kernel + deferred sweep)
VARIANTS = {
    "tf32-warp": {"CSB200_SCREEN_F16": "0", "CSB200_UPD_WARP": "1", "CSB200_UPD_DEFER": "2"},
    "f16-warp": {"CSB200_SCREEN_F16": "1", "CSB200_UPD_WARP": "1", "CSB200_UPD_DEFER": "2"},
    "f16-warp-1slice": {"CSB200_SCREEN_F16": "1", "CSB200_UPD_WARP": "1", "CSB200_UPD_DEFER": "1"},
    "f16-cta": {"CSB200_SCREEN_F16": "1", "CSB200_UPD_WARP": "0", "CSB200_UPD_DEFER": "0"},
    "tf32-cta-deferred": {"CSB200_SCREEN_F16": "0", "CSB200_UPD_WARP": "0", "CSB200_UPD_DEFER": "4"},
    "tf32-cta-ring": {"CSB200_SCREEN_F16": "0", "CSB200_UPD_WARP": "0", "CSB200_UPD_DEFER": "0", "CSB200_UPD_RING": "4",
                      "CSB200_UPD_HINTS": "3"},
}


def _planted(po, rng, A, B, k, noise=0.0):
    M, N = A.shape
    idx = np.stack([rng.choice(N, size=k, replace=False) for _ in range(B)])
    sign = rng.choice(np.array([-1.0, 1.0]), size=(B, k))
    Bm = np.einsum("msk,sk->ms", A[:, idx], sign)
    if noise:
        Bm = Bm + noise * rng.standard_normal(Bm.shape)
    return np.asfortranarray(Bm), idx


@pytest.mark.parametrize("f16", [0, 1])
@pytest.mark.parametrize("M,N,B", [(256, 2048, 300), (100, 300, 130), (1024, 8192, 257), (96, 4096 + 40, 128)])
def test_screening_pass_respects_its_bound_and_lists_the_argmax(cs, po, monkeypatch, M, N, B, f16):
    monkeypatch.setenv("CSB200_SCREEN_F16", str(f16))
    rng = np.random.default_rng(M + N + B)
    A = po.gaussian_dictionary(rng, M, N)
    Bm, _ = _planted(po, rng, A, B, 6, noise=0.01)
    Bm[:, 3] *= 1e-6                                             # the bound scales with ||r||
    Bm[:, 4] *= 1e5
    with cs.Dictionary(A) as D, cs.Batch(D, B, 8) as batch:
        batch.upload(Bm)
        val, idx, bound = batch.debug_screen_pass()
    C = np.abs(A.T @ Bm)                                         # FP64 reference correlations
    nrm = np.linalg.norm(Bm, axis=0)
    nc = val.shape[1]
    chunks = nc // 8
    assert bound == pytest.approx((_kappa_f16(M) if f16 else _kappa(M)) * np.max(np.linalg.norm(A, axis=0)))
    worst = 0.0
    for s in range(B):
        E = bound * nrm[s]
        ok = idx[s] >= 0
        assert ok.any()
        assert len(set(idx[s][ok].tolist())) == ok.sum()                         # no atom listed twice
        err = np.abs(val[s][ok] - C[idx[s][ok], s])
        worst = max(worst, float(err.max() / E))
        assert (err <= E).all(), (s, err.max(), E)
        for c in range(chunks):                                                  # lists are sorted, descending
            v = val[s, c * 8:(c + 1) * 8][idx[s, c * 8:(c + 1) * 8] >= 0]
            assert (np.diff(v) <= 0).all()
        j = int(np.argmax(C[:, s]))
        vmax = val[s][ok].max()
        # the FP64 arg-max is listed, or its whole chunk list lies inside the window (=> exact scan in the solve)
        listed = j in idx[s][ok].tolist()
        tails_in_window = any(idx[s, c * 8 + 7] >= 0 and val[s, c * 8 + 7] >= vmax - 2 * E for c in range(chunks))
        assert listed or tails_in_window, s
        # nothing outside the lists can beat the listed maximum by more than the bound allows
        mask = np.ones(N, bool)
        mask[idx[s][ok]] = False
        if mask.any():
            assert C[mask, s].max() <= max(val[s, c * 8:(c + 1) * 8].min() for c in range(chunks)) + E + 1e-300
    print(f"screening {M}x{N} ({'fp16' if f16 else 'tf32'}): worst |c~ - c| / bound = {worst:.3f}")
    assert worst < 0.5                                           # the Cauchy-Schwarz bound is far from tight on Gaussian data


@pytest.mark.parametrize("f16", [0, 1])
def test_screening_bound_holds_without_cancellation(cs, po, monkeypatch, f16):
    """All-positive operands: every product a_i r_i has the same sign, so sum |a_i r_i| = |c| (the Cauchy-Schwarz step of the
    bound is tight when r is parallel to an atom) and an accumulator that truncates would lose up to one ulp per add, all in
    the same direction.  The measured error must still be inside the bound -- this is the case that pins the accumulation
    term M * 2^-22 of screen_kappa(M)."""
    monkeypatch.setenv("CSB200_SCREEN_F16", str(f16))
    rng = np.random.default_rng(2024)
    M, N, B = 4096, 512, 128
    A = np.abs(rng.standard_normal((M, N))) + 0.05
    A /= np.linalg.norm(A, axis=0)
    A = np.asfortranarray(A)
    Bm = np.asfortranarray(A[:, rng.integers(0, N, size=B)] * rng.uniform(0.5, 2.0, size=B))   # r parallel to an atom
    Bm[:, B // 2:] = np.abs(rng.standard_normal((M, B - B // 2))) + 0.05
    with cs.Dictionary(A) as D, cs.Batch(D, B, 4) as batch:
        batch.upload(Bm)
        val, idx, bound = batch.debug_screen_pass()
    C = np.abs(A.T @ Bm)
    nrm = np.linalg.norm(Bm, axis=0)
    worst = 0.0
    for s in range(B):
        ok = idx[s] >= 0
        err = np.abs(val[s][ok] - C[idx[s][ok], s]) / (bound * nrm[s])
        worst = max(worst, float(err.max()))
        assert int(np.argmax(C[:, s])) in idx[s][ok].tolist() or (val[s][ok].min() >= val[s][ok].max() - 2 * bound * nrm[s])
    print(f"all-positive operands, M = {M}: worst |c~ - c| / bound = {worst:.3f}")
    assert worst < 1.0


def test_screened_mp_equals_the_dmma_path_and_the_oracle(cs, po, monkeypatch):
    """Plain mp (src/matchingpursuit.jl:26-40) through the screening pass: same atom sequence and coefficient increments as
    the FP64 DMMA path bit for bit, and the oracle's on a sample."""
    rng = np.random.default_rng(31)
    M, N, iters, B = 128, 1024, 12, 4096 + 5
    A = po.gaussian_dictionary(rng, M, N)
    Bm, _ = _planted(po, rng, A, B, 6, noise=1e-2)
    out = {}
    with cs.Dictionary(A) as D:
        for mode in ("0", "1"):
            monkeypatch.setenv("CSB200_SCREEN", mode)
            with cs.Batch(D, B, iters) as batch:
                batch.upload(Bm)
                batch.mp(iters)
                st = batch.screen_stats(reset=True)
                assert st["path_id"] == (3 if mode == "1" else 1), st
                out[mode] = batch.download(iters) + (batch.residual(),)
    for a, b in zip(out["0"], out["1"]):
        assert np.array_equal(a, b)
    sel, coef, nnz, res, its, R = out["1"]
    for s in (0, 17, B - 1):
        t = po.Trace()
        ref = po.mp(A, Bm[:, s], iters, trace=t)
        assert sel[s, :iters].tolist() == t.order()
        x = np.zeros(N)
        np.add.at(x, sel[s, :iters], coef[s, :iters])
        dense = np.zeros(N); dense[ref.nzind] = ref.nzval
        assert np.allclose(x, dense, rtol=RTOL64, atol=RTOL64)


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("M,N,k,B,noise", [(256, 2048, 8, 4096 + 37, 0.0), (100, 300, 5, 4096 + 130, 5e-3),
                                            (1024, 8192, 32, 4096, 0.0),
                                            (1500, 3000, 6, 4096 + 3, 1e-3),      # rows not a multiple of 256: ragged last row slot of the sweep
                                            (256, 2048, 40, 4096 + 11, 5e-3)])    # support capacity > 32: no warp-per-signal append
def test_screened_omp_equals_the_dmma_path_and_the_oracle(cs, po, monkeypatch, M, N, k, B, noise, variant):
    for key, value in VARIANTS[variant].items():
        monkeypatch.setenv(key, value)
    rng = np.random.default_rng(M * 3 + N + k)
    A = po.gaussian_dictionary(rng, M, N)
    Bm, planted = _planted(po, rng, A, B, k, noise)
    Bm[:, 5::97] = A[:, planted[5::97, 0]] * 2.0                 # 1-sparse signals: eps-break after the first update!
    out = {}
    with cs.Dictionary(A) as D:
        for mode in ("0", "1"):
            monkeypatch.setenv("CSB200_SCREEN", mode)
            with cs.Batch(D, B, k) as batch:
                batch.upload(Bm)
                batch.omp(k, 1e-9)
                st = batch.screen_stats(reset=True)
                assert st["path_id"] == ((4 if VARIANTS[variant]["CSB200_SCREEN_F16"] == "1" else 3) if mode == "1" else 2 if B >= 8192 else 1), st
                if mode == "1":
                    assert st["signal_updates"] > 0 and st["exact_scans"] < 0.01 * st["signal_updates"], st
                    print("screen stats", st)
                out[mode] = batch.download(k) + (batch.residual(),)
    sel0, coef0, nnz0, res0, its0, R0 = out["0"]
    sel1, coef1, nnz1, res1, its1, R1 = out["1"]
    assert np.array_equal(nnz0, nnz1) and np.array_equal(its0, its1)
    assert np.array_equal(sel0, sel1)                            # same supports in the same order
    assert np.array_equal(coef0, coef1) and np.array_equal(R0, R1)   # same update kernel on the same atoms: same bits
    if noise == 0.0:
        ok = np.ones(B, bool); ok[5::97] = False
        assert all(set(sel1[s, :k].tolist()) == set(planted[s].tolist()) for s in np.flatnonzero(ok)[:512])
    for s in (0, 5, 102, B - 1):
        t = po.Trace()
        ref = po.omp(A, Bm[:, s], k, eps=1e-9, trace=t)
        n = int(nnz1[s])
        assert sel1[s, :n].tolist() == t.order() and int(its1[s]) == t.iterations
        o = np.argsort(sel1[s, :n], kind="stable")
        assert sel1[s, :n][o].tolist() == ref.nzind
        assert np.allclose(coef1[s, :n][o], ref.nzval, rtol=RTOL64, atol=RTOL64 * max(1.0, np.max(np.abs(ref.nzval))))


@pytest.mark.parametrize("parts", [2, 3])
def test_overlapped_screening_schedule_is_bit_identical(cs, po, monkeypatch, parts):
    """Batches of >= 8192 signals run the screening pass of one part of the batch over the update of another (two streams,
    3-stage pass co-resident with update CTAs): same kernels on the same data, so every output equals the serial schedule's."""
    monkeypatch.setenv("CSB200_SCREEN", "1")
    rng = np.random.default_rng(99 + parts)
    M, N, k, B = 100, 300, 5, 8192 + 77
    A = po.gaussian_dictionary(rng, M, N)
    Bm, planted = _planted(po, rng, A, B, k, noise=1e-3)
    Bm[:, 5::7] = A[:, planted[5::7, 0]] * 2.0
    out = {}
    with cs.Dictionary(A) as D:
        for p in (1, parts):
            monkeypatch.setenv("CSB200_SCREEN_PARTS", str(p))
            with cs.Batch(D, B, k) as batch:
                batch.upload(Bm)
                batch.profile(True)
                batch.omp(k, 1e-9)
                ms, launches, other = batch.corr_time()
                batch.profile(False)
                assert launches == p * k and ms > 0 and batch.screen_stats()["path_id"] in (3, 4)
                out[p] = batch.download(k) + (batch.residual(),)
    for a, b in zip(out[1], out[parts]):
        assert np.array_equal(a, b)
    sel, coef, nnz, res, its, R = out[parts]
    assert (its[5::7] == 1).all() and (nnz[5::7] == 1).all()
    for s in (0, 5, 4100, B - 1):
        t = po.Trace()
        po.omp(A, Bm[:, s], k, eps=1e-9, trace=t)
        assert sel[s, :int(nnz[s])].tolist() == t.order() and int(its[s]) == t.iterations


@pytest.mark.parametrize("variant", ["tf32-warp", "f16-warp", "f16-cta"])
def test_screened_omp_ties_zero_signals_and_out_of_range_norms(cs, po, monkeypatch, variant):
    """Duplicate atoms (bit-identical |c|: the lower index must win, KAT-4), an all-zero signal (arg-max of zeros is atom 0,
    which is appended with coefficient 0: KAT-6), signals whose norm is outside the range the FP32 operands cover (exact
    scan) -- all through the screening path."""
    monkeypatch.setenv("CSB200_SCREEN", "1")
    for key, value in VARIANTS[variant].items():
        monkeypatch.setenv(key, value)
    rng = np.random.default_rng(77)
    M, N, k, B = 64, 512, 3, 600                                # > 512 signals: not the small-dictionary whole-solve kernel
    A = po.gaussian_dictionary(rng, M, N)
    A[:, 400] = A[:, 17]
    A[:, 18] = -A[:, 17]
    Bm, planted = _planted(po, rng, A, B, k)
    Bm[:, 0] = 2.0 * A[:, 17] + 0.5 * A[:, 300]
    Bm[:, 1] = 0.0
    Bm[:, 2] *= 1e-25
    Bm[:, 3] *= 1e25
    with cs.Dictionary(A) as D, cs.Batch(D, B, k) as batch:
        batch.upload(Bm)
        batch.omp(k, 0.0)
        st = batch.screen_stats()
        sel, coef, nnz, res, its = batch.download(k)
    assert st["path_id"] == (4 if VARIANTS[variant]["CSB200_SCREEN_F16"] == "1" else 3) and st["exact_scans"] >= 3
    assert sel[0, 0] == 17 and sel[0, 1] == 300
    for s in (0, 1, 2, 3, 4, 100, B - 1):
        t = po.Trace()
        ref = po.omp(A, Bm[:, s], k, eps=0.0, trace=t)
        n = int(nnz[s])
        if s in (0, 1):          # once the residual is at rounding level the arg-max is ill-posed: compare the well-posed prefix
            n = min(n, 2 if s == 0 else 1)
        assert sel[s, :n].tolist() == t.order()[:n], s


@pytest.mark.parametrize("variant", ["tf32-warp", "f16-warp"])
def test_screened_omp_on_unnormalised_atoms(cs, po, monkeypatch, variant):
    """The reference does not require unit-norm atoms (`argmaxinner!` compares raw |<a_j, r>|, src/matchingpursuit.jl:181-185).
    Column norms spread over 0.05 .. 20: the bound scales with the LARGEST norm, so the windows of signals built from short atoms
    hold many candidates -- the decision must still be the FP64 one, bit for bit, and the oracle's."""
    monkeypatch.setenv("CSB200_SCREEN", "1")
    for key, value in VARIANTS[variant].items():
        monkeypatch.setenv(key, value)
    rng = np.random.default_rng(4242)
    M, N, k, B = 200, 1500, 6, 4096 + 9
    A = po.gaussian_dictionary(rng, M, N) * np.exp(rng.uniform(np.log(0.05), np.log(20.0), size=N))
    A = np.asfortranarray(A)
    Bm, _ = _planted(po, rng, A, B, k, noise=1e-3)
    out = {}
    with cs.Dictionary(A) as D:
        for mode in ("0", "1"):
            monkeypatch.setenv("CSB200_SCREEN", mode)
            with cs.Batch(D, B, k) as batch:
                batch.upload(Bm)
                batch.omp(k, 1e-9)
                out[mode] = batch.download(k) + (batch.residual(),)
    for a, b in zip(out["0"], out["1"]):
        assert np.array_equal(a, b)
    sel, coef, nnz, res, its, R = out["1"]
    for s in (0, 1, 2, 3, 500, B - 1):
        t = po.Trace()
        ref = po.omp(A, Bm[:, s], k, eps=1e-9, trace=t)
        n = int(nnz[s])
        assert sel[s, :n].tolist() == t.order() and int(its[s]) == t.iterations
        o = np.argsort(sel[s, :n], kind="stable")
        assert np.allclose(coef[s, :n][o], ref.nzval, rtol=RTOL64, atol=RTOL64 * max(1.0, np.max(np.abs(ref.nzval))))


def test_screened_omp_through_the_multi_device_handle(cs, po, monkeypatch):
    """ONE csb200_omp call on a handle with several workers (all visible GPUs; entries repeat on a single-GPU box), each worker's
    share large enough for the screening path: every worker sets up the tcgen05 pass on ITS device (kernel attributes are per
    device) and the result equals the single-device call bit for bit."""
    monkeypatch.setenv("CSB200_SCREEN", "1")
    rng = np.random.default_rng(808)
    M, N, k, B = 128, 1024, 6, 3 * 4096 + 50
    A = po.gaussian_dictionary(rng, M, N)
    Bm, _ = _planted(po, rng, A, B, k, noise=1e-3)
    ndev = cs.device_count()
    devices = list(range(ndev)) if ndev >= 3 else [i % ndev for i in range(3)]
    with cs.Dictionary(A) as D1, cs.Dictionary(A, devices=devices) as Dn:
        one, many = cs.omp(D1, Bm, k), cs.omp(Dn, Bm, k)
    assert len(one) == len(many) == B
    for s in range(B):
        assert np.array_equal(one[s].nzind, many[s].nzind) and np.array_equal(one[s].nzval, many[s].nzval), s
    for s in (0, 4096, B - 1):
        ref = po.omp(A, Bm[:, s], k)
        assert many[s].nzind.tolist() == ref.nzind
        assert np.allclose(many[s].nzval, ref.nzval, rtol=RTOL64, atol=RTOL64)
