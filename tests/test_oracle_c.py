"""The plain-C oracle (oracle/pursuit_oracle.c) against the NumPy oracle and the golden fixtures (CPU only).

Two independent restatements of the reference's mp / omp / gomp must agree: selection sequence and support exactly,
coefficients and residual norms to 1e-10 -- the same bar the CUDA path is held to.
"""
import glob
import json
import os

import numpy as np
import pytest

from oracle import c_oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-10


def _close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all(np.abs(a - b) <= rtol * np.maximum(1.0, np.abs(b))))


def _fixtures():
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        z = np.load(p, allow_pickle=False)
        meta = json.loads(str(z["meta"]))
        if meta["algo"] in ("omp", "gomp", "mp") and z["A"].dtype == np.float64:
            out.append(p)
    return out


@pytest.mark.parametrize("path", _fixtures(), ids=lambda p: os.path.basename(p)[:-4])
def test_c_oracle_reproduces_golden_fixtures(path):
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    A, Bm = z["A"], z["B"]
    k = meta["k"]
    got = c_oracle.solve_batch(meta["algo"], A, Bm, k, l=meta.get("l") or 1, eps=meta["eps"])
    for s in range(Bm.shape[1]):
        n = int(z["nnz"][s])
        assert int(got["nnz"][s]) == n, s
        assert got["nzind"][s, :n].tolist() == z["nzind"][s, :n].tolist(), s
        assert _close(got["nzval"][s, :n], z["nzval"][s, :n], 1e-9 if meta["algo"] == "mp" else RTOL), s
        if meta["algo"] != "mp":
            assert got["order"][s, :n].tolist() == z["order"][s, :n].tolist(), (s, "selection sequence")
        assert abs(got["resnorm"][s] - z["resnorm"][s]) <= RTOL * max(1.0, np.linalg.norm(Bm[:, s])) + 1e-14


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("algo", ["omp", "gomp", "mp"])
def test_c_oracle_matches_numpy_oracle_on_seeded_problems(po, algo, seed):
    rng = np.random.default_rng(1000 + seed)
    M, N, k, l = [(32, 48, 3, 2), (64, 200, 9, 4), (128, 256, 8, 3)][seed % 3]
    A = po.gaussian_dictionary(rng, M, N)
    if seed % 2 and algo != "gomp":                            # (gomp would append both twins: a singular active set,
        A[:, N - 1] = A[:, 1]                                  #  undefined in the reference)  exact tie: the lower index must win
    B = []
    for _ in range(5):
        x0 = po.sparse_vector(rng, N, k)
        B.append(po.perturb(rng, A[:, x0.nzind] @ np.asarray(x0.nzval), 5e-3))
    Bm = np.stack(B, axis=1)
    got = c_oracle.solve_batch(algo, A, Bm, k, l=l, threads=3)
    assert 1 <= got["threads"] <= 3
    for s in range(Bm.shape[1]):
        t = po.Trace()
        ref = {"omp": lambda: po.omp(A, Bm[:, s], k, trace=t), "gomp": lambda: po.gomp(A, Bm[:, s], l, k, trace=t),
               "mp": lambda: po.mp(A, Bm[:, s], k, trace=t)}[algo]()
        n = ref.nnz()
        assert int(got["nnz"][s]) == n
        assert got["nzind"][s, :n].tolist() == ref.nzind
        assert _close(got["nzval"][s, :n], ref.nzval)
        assert got["order"][s, :len(t.order())].tolist() == t.order()
        assert abs(got["resnorm"][s] - t.resnorm[-1]) < 1e-12
        assert int(got["iters"][s]) == t.iterations


def test_c_oracle_quirks(po):
    """KAT-5 / KAT-6 / KAT-7 of SURVEY 8c through the C restatement."""
    # zero signal: one stored zero at the first atom, one iteration (the eps test fires after the update)
    rng = np.random.default_rng(5)
    A = po.gaussian_dictionary(rng, 16, 24)
    got = c_oracle.solve_batch("omp", A, np.zeros((16, 1)), 4)
    assert got["nnz"][0] == 1 and got["nzind"][0, 0] == 0 and got["nzval"][0, 0] == 0.0 and got["iters"][0] == 1
    # no-op iterations: A = I, b = e0 + e1, k = 3, eps = 0: the third update re-selects atom 0 and changes nothing
    got = c_oracle.solve_batch("omp", np.eye(6), np.array([1.0, 1.0, 0, 0, 0, 0]).reshape(6, 1), 3, eps=0.0)
    assert got["nnz"][0] == 2 and got["nzind"][0, :2].tolist() == [0, 1] and got["iters"][0] == 3
    assert got["order"][0, :2].tolist() == [0, 1]
    # gomp remainder runs even after an eps-break: l = 2, k = 3 on an exactly 2-sparse signal
    Bm = (A[:, 3] - A[:, 7]).reshape(16, 1)
    got = c_oracle.solve_batch("gomp", A, Bm, 3, l=2, eps=1e-8)
    ref = po.gomp(A, Bm[:, 0], 2, 3, eps=1e-8)
    assert got["nzind"][0, :got["nnz"][0]].tolist() == ref.nzind and got["iters"][0] == 2
    with pytest.raises(ValueError):
        c_oracle.solve_batch("omp", A, Bm, 3, eps=-1.0)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line with the contract's
    keys; under torchrun every rank but 0 exits without output."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--ref-signals", "8"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "solves/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("OMP solves/sec at 1024x8192,k=32") and line["value"] > 0
    assert line["steps"] == 1 and line["warmup"] == 1 and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"})
    assert r1.returncode == 0 and r1.stdout.strip() == ""
