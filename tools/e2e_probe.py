"""Where the one-shot call's time goes (under gpurun): csb200_omp on 65 536 pinned signals at config 2, pipelined vs single upload,
next to the raw H2D time of the same buffer and the device-resident solve."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
cs = g.load_package()
from ctypes import POINTER, c_double, c_int64

M, N, k, B = 1024, 8192, 32, 65536
rng = np.random.default_rng(1)
A = rng.standard_normal((M, N)); A /= np.linalg.norm(A, axis=0); A = np.asfortranarray(A)
B_pin = torch.empty(B, M, dtype=torch.float64, pin_memory=True)
Bn = B_pin.numpy()
idx = np.stack([rng.choice(N, size=k, replace=False) for _ in range(B)])
for s0 in range(0, B, 4096):
    Bn[s0:s0 + 4096] = np.einsum("msk,sk->sm", A[:, idx[s0:s0 + 4096]], rng.choice(np.array([-1.0, 1.0]), size=(4096, k)))
B_np = Bn.T
dev = torch.device("cuda:0")
d = torch.empty(B, M, dtype=torch.float64, device=dev)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(B_pin, non_blocking=True); torch.cuda.synchronize(); h2d = time.perf_counter() - t0
out = {"h2d_512MiB_ms": h2d * 1e3, "h2d_GBs": B * M * 8 / h2d / 1e9}
o_sel = np.empty((B, k), dtype=np.int64); o_coef = np.empty((B, k)); o_nnz = np.empty(B, dtype=np.int64); o_res = np.empty(B); o_it = np.empty(B, dtype=np.int64)
i64p, f64p = POINTER(c_int64), POINTER(c_double)
with cs.Dictionary(A) as D:
    def once():
        rc = cs.lib.csb200_omp(D._h, B_np.ctypes.data, M, B, k, 1e-12, o_sel.ctypes.data_as(i64p), o_coef.ctypes.data_as(f64p),
                               o_nnz.ctypes.data_as(i64p), o_res.ctypes.data_as(f64p), o_it.ctypes.data_as(i64p))
        assert rc == 0, cs.lib.csb200_last_error()
    os.environ["CSB200_PIPE_DEBUG"] = "1"
    for mode in ("1",):
        os.environ["CSB200_PIPELINE"] = mode
        once()
        ts = []
        for _ in range(1):
            t0 = time.perf_counter(); once(); ts.append((time.perf_counter() - t0) * 1e3)
        out[f"csb200_omp_ms_pipeline={mode}"] = ts
    with cs.Batch(D, B, k) as batch:
        batch.upload(B_np)
        batch.omp(k, 1e-12)
        batch.omp(k, 1e-12)
        out["resident_solve_ms"] = batch.last_solve_ms()
        t0 = time.perf_counter(); r = batch.download(k); out["download_ms"] = (time.perf_counter() - t0) * 1e3
print(json.dumps(out))
