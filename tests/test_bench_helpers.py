"""CPU tests of bench.py's host-side pieces: the synthetic-input generators (same bytes as the oracle's generator), the
support sampler (distinct atoms), and the oracle-parity check it publishes in `check.oracle_parity`."""
import numpy as np

import bench
from oracle import c_oracle
from oracle import pursuit_oracle as po


def test_dictionary_bytes_equal_the_oracle_generator():
    a = bench.gaussian_dictionary_np(np.random.default_rng(1234), 64, 96)
    b = po.gaussian_dictionary(np.random.default_rng(1234), 64, 96)
    assert a.flags["F_CONTIGUOUS"] and np.array_equal(a, b)
    assert np.allclose(np.linalg.norm(a, axis=0), 1.0, rtol=0, atol=1e-14)


def test_supports_are_distinct_and_signs_are_unit():
    idx, sign = bench.draw_supports_np(np.random.default_rng(5680), 4000, 64, 32)      # k = N / 2: repeats are the rule
    assert idx.shape == sign.shape == (4000, 32)
    assert all(len(set(row.tolist())) == 32 for row in idx)
    assert set(np.unique(sign).tolist()) == {-1.0, 1.0}


def test_parity_check_accepts_the_oracle_and_rejects_a_perturbation():
    rng = np.random.default_rng(3)
    M, N, k, ns = 48, 120, 5, 6
    A = bench.gaussian_dictionary_np(rng, M, N)
    idx, sign = bench.draw_supports_np(rng, ns, N, k)
    Bm = np.asfortranarray(np.stack([A[:, idx[s]] @ sign[s] for s in range(ns)], axis=1))
    got = c_oracle.solve_batch("omp", A, Bm, k)
    # the GPU side reports (atom, coefficient) pairs in SELECTION order: rebuild that view from the oracle's own output
    sel = got["order"][:, :k].copy()
    coef = np.zeros((ns, k))
    for s in range(ns):
        lut = dict(zip(got["nzind"][s, :k].tolist(), got["nzval"][s, :k].tolist()))
        coef[s] = [lut[j] for j in sel[s].tolist()]
    ok = bench.parity_vs_c_oracle(got, sel, coef, got["nnz"], got["resnorm"], Bm)
    assert ok["pass"] and ok["selection_order_exact_frac"] == 1.0 and ok["coef_max_rel_err"] == 0.0
    bad = coef.copy(); bad[2, 1] *= 1 + 1e-8
    assert not bench.parity_vs_c_oracle(got, sel, bad, got["nnz"], got["resnorm"], Bm)["pass"]
    swapped = sel.copy(); swapped[1, [0, 1]] = swapped[1, [1, 0]]
    r = bench.parity_vs_c_oracle(got, swapped, coef, got["nnz"], got["resnorm"], Bm)
    assert not r["pass"] and r["selection_order_exact_frac"] < 1.0


def test_cli_defaults_match_the_driver_contract():
    a = bench.parse([])
    assert (a.gpus, a.config, a.impl, a.scaling, a.secondary) == (1, "c2", "ours", "weak", "auto") and a.warmup >= 3


def test_tensor_peaks_come_from_the_driver_file_or_the_stated_fallback():
    """The roofline denominators of the screening pass: MEASURED_PEAKS.json's bf16 figure for the FP16 pass (kind::f16), half of
    it for the TF32 pass (kind::tf32 runs at half the rate), else the nominal numbers of B200_PROFILING.md -- and the source
    string says which."""
    f16, f16_src = bench.f16_peak()
    tf32, tf32_src = bench.tf32_peak()
    assert f16 == 2.0 * tf32
    assert 1000.0 <= f16 <= 2500.0 and 500.0 <= tf32 <= 1250.0
    assert ("MEASURED_PEAKS.json" in f16_src) == ("MEASURED_PEAKS.json" in tf32_src)
    assert "MEASURED_PEAKS.json" in f16_src or "fallback" in f16_src
