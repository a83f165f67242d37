"""Times the correlation GEMM alone (C2 shape) through the debug hook, for CSB200_GEMM_BAND given in the env."""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
cs = ge.load_package()
M, N, B = 1024, 8192, int(os.environ.get("SWEEP_B", 65536))
rng = np.random.default_rng(0)
A = np.asfortranarray(rng.standard_normal((M, N)))
R = np.asfortranarray(rng.standard_normal((M, B)))
with cs.Dictionary(A) as D, cs.Batch(D, B, 1) as b:
    b.upload(R)
    for _ in range(2):
        b.debug_corr_topk(1, 1)
    b.profile(True)
    for _ in range(6):
        b.debug_corr_topk(1, 1)
    ms, n, _ = b.corr_time()
print(json.dumps({"band": os.environ.get("CSB200_GEMM_BAND", "default"), "mean_ms": ms / n, "tflops": 2.0 * M * N * B * n / ms / 1e9}))
