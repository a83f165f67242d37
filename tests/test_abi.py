"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/csb200.h declares,
and fails loudly (no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "csb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(csb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(cs):
    names = _declared_symbols()
    assert len(names) >= 25
    L = ctypes.CDLL(cs.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/csb200.h but not exported by libcsb200.so"
    assert sorted(cs.EXPORTED_SYMBOLS) == names


def test_version_and_strerror(cs):
    assert cs.lib.csb200_version() == 100
    assert cs.lib.csb200_strerror(0) == b"ok"
    assert b"non-negative" in cs.lib.csb200_strerror(-2)
    assert b"unknown" in cs.lib.csb200_strerror(-99)


def test_library_contains_only_sm100a_code(cs):
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", cs.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_sass_has_dmma_and_tma(cs):
    """Evidence that the batched kernel is the DMMA + TMA one (B200_PROFILING.md: UTMALDG = cp.async.bulk.tensor)."""
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", cs.LIB_PATH], capture_output=True, text=True).stdout
    assert "DMMA.8x8x4" in sass
    assert "UTMALDG" in sass
    assert "SYNCS" in sass          # mbarrier


def test_no_cpu_fallback_without_gpu(cs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    A = np.asfortranarray(np.eye(4))
    with pytest.raises(cs.CSB200Error):
        cs.omp(A, np.ones(4), 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "compressedsensing.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text.replace(
                    "`oracle/`", ""), (dirpath, f)


def test_argument_dispatch_mirrors_julia_methods(cs, monkeypatch):
    """omp(A,b,k::Int) vs omp(A,b,eps::Real,k) vs keywords -- checked without touching the GPU."""
    calls = []

    class FakeLib:
        def __init__(self, real): self._real = real
        def __getattr__(self, name): return getattr(self._real, name)
        def csb200_omp(self, h, B, ldb, nsig, k, eps, sel, coef, nnz, res, its):
            calls.append(("omp", k, eps)); return 0
        def csb200_gomp(self, h, B, ldb, nsig, l, k, eps, sel, coef, nnz, res, its):
            calls.append(("gomp", l, k, eps)); return 0

    class FakeDict:
        def __init__(self, A, device=0):
            self.M, self.N = A.shape; self.n_total = self.N; self.dtype = A.dtype; self._h = None
        def close(self): pass

    monkeypatch.setattr(cs, "lib", FakeLib(cs.lib))
    monkeypatch.setattr(cs, "Dictionary", FakeDict)
    monkeypatch.setattr(cs, "_to_sparse", lambda n, sel, coef, nnz: cs.SparseVector(n, np.zeros(0, np.int64), np.zeros(0)))
    A = np.zeros((8, 12)); b = np.zeros(8)
    eps64 = np.finfo(np.float64).eps
    cs.omp(A, b, 3);                       assert calls[-1] == ("omp", 3, eps64)
    cs.omp(A, b, 0.5);                     assert calls[-1] == ("omp", 8, 0.5)          # k defaults to size(A,1)
    cs.omp(A, b, 0.5, 4);                  assert calls[-1] == ("omp", 4, 0.5)
    cs.omp(A, b);                          assert calls[-1] == ("omp", 8, eps64)        # sparsity = min(size(A)...)
    cs.omp(A, b, max_residual=0.1, sparsity=2); assert calls[-1] == ("omp", 2, 0.1)
    cs.omp(A.astype(np.float32), b, 3);    assert calls[-1] == ("omp", 3, float(np.finfo(np.float32).eps))
    cs.gomp(A, b, 2, 3);                   assert calls[-1] == ("gomp", 2, 3, eps64)
    cs.gomp(A, b, 2, 0.25, 6);             assert calls[-1] == ("gomp", 2, 6, 0.25)
    cs.gomp(A, b, 2);                      assert calls[-1] == ("gomp", 2, 12, eps64)   # sparsity = size(A,2)
    with pytest.raises(ValueError, match="has to be non-negative"):
        cs.omp(A, b, -1.0)
    with pytest.raises(ValueError, match="has to be non-negative"):
        cs.gomp(A, b, 2, -1.0, 3)
    with pytest.raises(ValueError):
        cs.omp(A, np.zeros(7), 3)           # DimensionMismatch
    out = cs.omp(A, np.zeros((8, 5)), 3)
    assert isinstance(out, list) and len(out) == 5 and out[0].n == 12


def test_widened_front_ends_dispatch(cs, monkeypatch):
    """fr / ols / oomp / ormp (src/forward.jl:33-54), sp (src/twostage.jl:105), oblivious (src/oblivious.jl:3):
    argument forms of the Julia methods, checked without touching the GPU; null handles are rejected by the C ABI."""
    calls = []

    class FakeLib:
        def __init__(self, real): self._real = real
        def __getattr__(self, name): return getattr(self._real, name)
        def csb200_fr(self, h, B, ldb, nsig, k, max_eps, min_delta, sel, coef, nnz, res, its):
            calls.append(("fr", k, max_eps, min_delta)); return 0
        def csb200_sp(self, h, B, ldb, nsig, k, delta, maxiter, sel, coef, nnz, res, its):
            calls.append(("sp", k, delta, maxiter)); return 0
        def csb200_oblivious(self, h, B, ldb, nsig, k, sel, coef, nnz, res):
            calls.append(("oblivious", k)); return 0

    class FakeDict:
        def __init__(self, A, device=0):
            self.M, self.N = A.shape; self.n_total = self.N; self.dtype = A.dtype; self._h = None
        def close(self): pass

    real = cs.lib
    monkeypatch.setattr(cs, "lib", FakeLib(real))
    monkeypatch.setattr(cs, "Dictionary", FakeDict)
    monkeypatch.setattr(cs, "_to_sparse", lambda n, sel, coef, nnz: cs.SparseVector(n, np.zeros(0, np.int64), np.zeros(0)))
    A = np.zeros((8, 12)); b = np.zeros(8)
    cs.fr(A, b);                                   assert calls[-1] == ("fr", 8, 0.0, 0.0)      # sparsity = size(A,2), capped at M
    cs.fr(A, b, sparsity=3);                       assert calls[-1] == ("fr", 3, 0.0, 0.0)
    cs.fr(A, b, 0.1, 0.2);                         assert calls[-1] == ("fr", 8, 0.1, 0.2)      # k defaults to size(A,1)
    cs.ols(A, b, 0.1, 0.2, 5);                     assert calls[-1] == ("fr", 5, 0.1, 0.2)
    cs.oomp(A, b, max_residual=0.3, min_decrease=0.4); assert calls[-1] == ("fr", 8, 0.3, 0.4)
    with pytest.raises(TypeError):
        cs.ormp(A, b, 0.1)
    cs.sp(A, b, 3);                                assert calls[-1] == ("sp", 3, 1e-12, 48)     # maxiter = 16k
    cs.sp(A, b, 2, 1e-3, maxiter=5);               assert calls[-1] == ("sp", 2, 1e-3, 5)
    with pytest.raises(ValueError, match="invalid for Subspace Pursuit"):
        cs.sp(A, b, 5)                                                                        # 2k > M
    cs.oblivious(A, b, 4);                         assert calls[-1] == ("oblivious", 4)
    # the real library: null handles / bad arguments come back as status codes, never a crash
    assert real.csb200_fr(None, None, 8, 1, 3, 0.0, 0.0, None, None, None, None, None) == -1
    assert real.csb200_sp(None, None, 8, 1, 3, 1e-12, 10, None, None, None, None, None) == -1
    assert real.csb200_oblivious(None, None, 8, 1, 3, None, None, None, None) == -1
    assert real.csb200_dict_cumbabel(None, 3, None) == -1 and real.csb200_dict_colnorms(None, None) == -1
    assert real.csb200_batch_fr(None, 3, 0.0, 0.0) == -1 and real.csb200_batch_sp(None, 3, 0.0, 4) == -1
    assert real.csb200_batch_oblivious(None, 3) == -1


def test_assemble_csc_host_helper(cs):
    """Batched result format: selection-order outputs -> CSC arrays, rows ascending per column."""
    sel = np.array([[5, 2, 9, -1], [-1, -1, -1, -1], [7, 0, -1, -1]], dtype=np.int64)
    coef = np.array([[1.0, 2.0, 3.0, 0], [0, 0, 0, 0], [4.0, 5.0, 0, 0]])
    nnz = np.array([3, 0, 2], dtype=np.int64)
    colptr, rowval, nzval = cs.assemble_csc(12, sel, coef, nnz)
    assert colptr.tolist() == [0, 3, 3, 5]
    assert rowval.tolist() == [2, 5, 9, 0, 7] and nzval.tolist() == [2.0, 1.0, 3.0, 5.0, 4.0]
