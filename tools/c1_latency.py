"""Config 1 (128 x 256, k = 8, one signal, FP64) through the batch API: device time per solve of the cluster-resident
whole-solve kernel (csrc/solve_small.cu).  Used under ncu for the per-instruction sampling in profiles/.

Usage (GPU box):  python tools/c1_latency.py [repeats]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import bench
import __graft_entry__ as ge


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    cs = ge.load_package()
    M, N, k = 128, 256, 8
    A = bench.gaussian_dictionary_np(np.random.default_rng(1234), M, N)
    idx, sign = bench.draw_supports_np(np.random.default_rng(5678 + 1), 1, N, k)
    B = np.asfortranarray((A[:, idx[0]] @ sign[0]).reshape(M, 1))
    with cs.Dictionary(np.asfortranarray(A), device=0) as D, cs.Batch(D, 1, k) as b:
        b.upload(B)
        for _ in range(10):
            b.omp(k, 1e-30)
        ms = []
        for _ in range(reps):
            b.omp(k, 1e-30)
            ms.append(b.last_solve_ms())
        sel, _, nnz, _, _ = b.download(k)
    ok = set(idx[0].tolist()) == set(sel[0, :int(nnz[0])].tolist())
    print(f"c1: {1e3 * float(np.mean(ms)):.1f} us per solve (device, mean of {reps}; median {1e3 * float(np.median(ms)):.1f}), support recovered: {ok}")


if __name__ == "__main__":
    main()
