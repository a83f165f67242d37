"""Per-phase clock64() timeline of the few-signal whole-solve kernel (csrc/solve_persist.cu) at the config-2 shape.

Usage (GPU box):  CSB200_PERSIST_DEBUG=1 python tools/persist_timeline.py [signals]
The library prints the averaged cycles per `update!` for the updater of signal 0 and for worker 0 to stderr.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import bench
import __graft_entry__ as ge


def main():
    ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    cs = ge.load_package()
    M, N, k = 1024, 8192, 32
    A = bench.gaussian_dictionary_np(np.random.default_rng(1234), M, N)
    idx, sign = bench.draw_supports_np(np.random.default_rng(5678 + 2), ns, N, k)
    B = np.empty((M, ns), order="F")
    for s in range(ns):
        B[:, s] = A[:, idx[s]] @ sign[s]
    with cs.Dictionary(np.asfortranarray(A), device=0) as D, cs.Batch(D, ns, k) as b:
        b.upload(B)
        for _ in range(5):
            b.omp(k, 1e-30)
        ms = []
        for _ in range(50):
            b.omp(k, 1e-30)
            ms.append(b.last_solve_ms())
        sel, _, nnz, _, _ = b.download(k)
    ok = all(set(idx[s].tolist()) == set(sel[s, :int(nnz[s])].tolist()) for s in range(ns))
    print(f"{ns} signal(s): {1e3 * float(np.mean(ms)):.1f} us per solve (device, mean of 50), supports recovered: {ok}")


if __name__ == "__main__":
    main()
