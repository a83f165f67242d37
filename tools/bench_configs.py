#!/usr/bin/env python
"""Secondary measurements kept beside bench.py (which now measures every BASELINE config itself: `bench.py --config
c1|c2s|c3|c4|c5`, and all of them inside the default line's `secondary` object).  What remains unique here: `fr2` and
`sp2` (the widened rows at the config-2 shape) and the older per-config printouts the round-1 profiles quote.

    python tools/bench_configs.py --config c1|c3|c4|c5 [--scale f]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_configs.py --config c4     (column-sharded)

Prints one JSON line per config (rank 0).  These are NOT the driver's bench line; they back the numbers quoted in
DESIGN.md.  Inputs are synthetic (Gaussian unit-norm atoms, planted k-sparse +-1 signals), generated on the GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

HBM_PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
FP64_PEAK = json.load(open(os.path.join(ROOT, "profiles", "FP64_PEAK.json")))["peak_tflops"]


def make_problem(M, N, k, B, dtype, dev, seed=1234, noise=0.0):
    """Synthetic problem generated on the device (bench.make_problem): Gaussian unit-norm atoms, planted k-sparse +-1
    signals whose k atoms are DISTINCT (sampled without replacement, `sparse_vector`, src/util.jl:13-19)."""
    import bench

    class _Ctx:
        pass
    ctx = _Ctx()
    ctx.torch, ctx.dev = torch, dev
    return bench.make_problem(ctx, M, N, k, B, dtype, seed=seed, noise=noise)


def cpu_oracle(fn, n):
    """CPU restatement of the reference (oracle/, NumPy + multi-threaded OpenBLAS) timed on n signals of the same
    workload on this box's host cores: the reported CPU baseline beside each GPU number (kind = "port")."""
    from oracle import pursuit_oracle  # noqa: F401  (checker used as the CPU baseline only)
    t0 = time.perf_counter()
    for s in range(n):
        fn(s)
    dt = time.perf_counter() - t0
    return {"solves_per_s": n / dt, "signals": n, "seconds": dt, "cores": os.cpu_count(), "kind": "port"}


def cpu_oracle_c(algo, A, Bm, n, k, l=1, eps=None):
    """The plain-C restatement (oracle/pursuit_oracle.c), one signal per host thread, on n signals (columns of Bm,
    cycled): the stronger CPU baseline for the Float64 omp / gomp / mp configs.  None when it cannot be built."""
    try:
        from oracle import c_oracle
        c_oracle.lib()
    except Exception as exc:
        sys.stderr.write(f"C oracle unavailable ({exc})\n")
        return None
    if A.dtype != np.float64:
        return None
    cols = np.asfortranarray(Bm[:, [s % Bm.shape[1] for s in range(n)]])
    t0 = time.perf_counter()
    got = c_oracle.solve_batch(algo, A, cols, k, l=l, eps=eps)
    dt = time.perf_counter() - t0
    return {"solves_per_s": n / dt, "signals": n, "seconds": dt, "cores": int(got["threads"]), "kind": "port",
            "engine": "plain-C oracle, one signal per thread"}


def c1(cs, dev, args):
    M, N, k = 128, 256, 8
    A, Bm, idx = make_problem(M, N, k, 64, np.float64, dev)
    out = {}
    with cs.Dictionary(A) as D:
        for name, nsig in [("single_signal", 1), ("batch64", 64)]:
            with cs.Batch(D, nsig, k) as b:
                b.upload(Bm[:, :nsig])
                for _ in range(20):
                    b.omp(k, 2.2e-16)
                ms = []
                for _ in range(200):
                    b.omp(k, 2.2e-16)
                    ms.append(b.last_solve_ms())
                out[name + "_us_per_solve_call"] = 1e3 * float(np.median(ms))
        t0 = time.perf_counter()
        for _ in range(200):
            cs.omp(D, Bm[:, 0], k)
        out["one_shot_host_api_us"] = 1e6 * (time.perf_counter() - t0) / 200
    out.update(config="c1 omp 128x256 k=8 f64", dict_bytes=M * N * 8, iterations=k)
    if args.cpu:
        from oracle import pursuit_oracle as po
        out["cpu_baseline"] = cpu_oracle(lambda s: po.omp(A, Bm[:, s % 64], k, ls="givens"), 256)
        out["cpu_baseline_us_per_solve"] = 1e6 / out["cpu_baseline"]["solves_per_s"]
        c = cpu_oracle_c("omp", A, Bm, 1, k)                     # one signal, one thread: the latency a caller sees
        if c:
            c1t = min(cpu_oracle_c("omp", A, Bm[:, :1], 1, k)["seconds"] for _ in range(20))
            out["cpu_c_oracle_us_per_solve_1thread"] = 1e6 * c1t
    return out


def c2s(cs, dev, args):
    """Single-signal omp on the config-2 dictionary (1024 x 8192: 64 MiB FP64 / 32 MiB FP32 -- smaller than the
    126 MB L2): the GEMV path with the dictionary L2-resident.  GB/s above the HBM peak = served from L2."""
    M, N, k = 1024, 8192, 32
    out = dict(config="c2s single-signal omp 1024x8192 k=32 (dictionary L2-resident)", hbm_peak_GBps=HBM_PEAK)
    for dt, name in [(np.float64, "f64"), (np.float32, "f32")]:
        A, Bm, idx = make_problem(M, N, k, 4, dt, dev)
        with cs.Dictionary(A) as D, cs.Batch(D, 1, k) as b:
            b.upload(Bm[:, :1])
            for _ in range(5):
                b.omp(k, 1e-30)
            b.profile(True)
            ms, corr, nl = [], 0.0, 0
            for _ in range(20):
                b.omp(k, 1e-30)
                ms.append(b.last_solve_ms())
            corr, nl, other = b.corr_time()
            b.profile(False)
            plain = []
            for _ in range(50):
                b.omp(k, 1e-30)
                plain.append(b.last_solve_ms())
            sel, coef, nnz, res, its = b.download(k)
        gemv_us = 1e3 * corr / max(1, nl)
        gbs = M * N * A.itemsize / (gemv_us * 1e-6) / 1e9
        out[name] = dict(us_per_solve=1e3 * float(np.median(plain)), us_per_solve_profiled=1e3 * float(np.median(ms)),
                         gemv_us_per_launch=gemv_us, gemv_GBps=gbs, gemv_vs_hbm_peak=gbs / HBM_PEAK,
                         gemv_share=corr / float(np.sum(ms)), dict_bytes=M * N * A.itemsize,
                         support_recovered=bool(set(idx[0]) == set(sel[0, :nnz[0]].tolist())))
        if args.cpu:
            from oracle import pursuit_oracle as po
            c = cpu_oracle(lambda s: po.omp(A, Bm[:, s % 4], k), 16)
            out[name]["cpu_baseline_us_per_solve"] = 1e6 / c["solves_per_s"]
            out[name]["cpu_cores"] = c["cores"]
            cc = cpu_oracle_c("omp", A, Bm, 1, k)
            if cc:
                out[name]["cpu_c_oracle_us_per_solve_1thread"] = 1e6 * cc["seconds"]
    return out


def c3(cs, dev, args):
    M, N, k, l, B = 2048, 32768, 64, 4, int(8192 * args.scale)
    A, Bm, idx = make_problem(M, N, k, B, np.float64, dev)
    with cs.Dictionary(A) as D, cs.Batch(D, B, k) as b:
        b.upload(Bm)
        b.gomp(l, k, 2.2e-16)
        b.profile(True)
        b.gomp(l, k, 2.2e-16)
        ms = b.last_solve_ms()
        corr_ms, n, _ = b.corr_time()
        sel, coef, nnz, res, its = b.download(k)
    rec = float(np.mean([(set(idx[s]) <= set(sel[s, :nnz[s]])) for s in range(0, B, 64)]))
    tf = 2.0 * M * N * B * n / corr_ms / 1e9
    out = dict(config="c3 gomp l=4 2048x32768 k=64 f64", signals=B, solves_per_s=B / (ms * 1e-3), ms_per_solve_batch=ms,
               corr_launches=int(n), corr_ms=corr_ms, corr_tflops=tf, frac_of_fp64_peak=tf / FP64_PEAK,
               corr_share=corr_ms / ms, support_recovered_frac=rec, max_resnorm=float(res.max()))
    if args.cpu:
        from oracle import pursuit_oracle as po
        out["cpu_baseline"] = cpu_oracle_c("gomp", A, Bm, 32, k, l=l) or \
            cpu_oracle(lambda s: po.gomp(A, Bm[:, s], l, k, ls="givens"), 12)
    return out


def fr2(cs, dev, args):
    """Forward regression / OLS at the config-2 shape: per step one DMMA pass over
    [residuals | newest directions] (4 M N B flop after the first step) + the per-signal update."""
    M, N, k, B = 1024, 8192, 32, int(65536 * args.scale)
    A, Bm, idx = make_problem(M, N, k, B, np.float64, dev)
    with cs.Dictionary(A) as D, cs.Batch(D, B, k) as b:
        b.upload(Bm)
        b.fr(k)
        b.profile(True)
        b.fr(k)
        ms = b.last_solve_ms()
        corr_ms, n, _ = b.corr_time()
        sel, coef, nnz, res, its = b.download(k)
    rec = float(np.mean([(set(idx[s]) <= set(sel[s, :nnz[s]])) for s in range(0, B, 64)]))
    flop = 2.0 * M * N * B * (2 * n - 1)                     # the first pass has no direction half
    tf = flop / corr_ms / 1e9
    out = dict(config="fr/ols 1024x8192 k=32 f64", signals=B, solves_per_s=B / (ms * 1e-3), ms_per_solve_batch=ms,
               corr_launches=int(n), corr_ms=corr_ms, corr_tflops=tf, frac_of_fp64_peak=tf / FP64_PEAK,
               corr_share=corr_ms / ms, support_recovered_frac=rec, max_resnorm=float(res.max()))
    if args.cpu:
        from oracle import pursuit_oracle as po
        out["cpu_baseline"] = cpu_oracle(lambda s: po.fr(A, Bm[:, s], 0.0, 0.0, k), 6)
    return out


def sp2(cs, dev, args):
    """Subspace pursuit at the config-2 shape, noisy signals (several update!s per signal)."""
    M, N, k, B = 1024, 8192, 32, int(16384 * args.scale)
    A, Bm, idx = make_problem(M, N, k, B, np.float64, dev, noise=1e-2)
    with cs.Dictionary(A) as D, cs.Batch(D, B, 2 * k) as b:
        b.upload(Bm)
        b.sp(k, 1e-12, 2)
        b.profile(True)
        b.sp(k)
        ms = b.last_solve_ms()
        corr_ms, n, other = b.corr_time()
        sel, coef, nnz, res, its = b.download(k)
    rec = float(np.mean([(set(idx[s]) <= set(sel[s, :nnz[s]])) for s in range(0, B, 64)]))
    tf = 2.0 * M * N * B * n / corr_ms / 1e9
    out = dict(config="sp 1024x8192 k=32 f64 noisy", signals=B, solves_per_s=B / (ms * 1e-3), ms_per_solve_batch=ms,
               corr_launches=int(n), corr_ms=corr_ms, corr_tflops=tf, frac_of_fp64_peak=tf / FP64_PEAK,
               corr_share=corr_ms / ms, update_ms_per_launch=(ms - corr_ms) / max(other, 1),
               updates_per_signal_mean=float(its.mean()), updates_per_signal_max=int(its.max()),
               support_recovered_frac=rec, median_resnorm=float(np.median(res)))
    if args.cpu:
        from oracle import pursuit_oracle as po
        out["cpu_baseline"] = cpu_oracle(lambda s: po.sp(A, Bm[:, s], k), 48)
    return out


def c5(cs, dev, args):
    M, N, iters, B = 4096, 65536, int(200 * args.scale), 4096
    A, Bm, idx = make_problem(M, N, 32, B, np.float64, dev, noise=5e-3)
    with cs.Dictionary(A) as D, cs.Batch(D, B, iters) as b:
        b.upload(Bm)
        b.mp(2)
        b.profile(True)
        b.mp(iters)
        ms = b.last_solve_ms()
        corr_ms, n, _ = b.corr_time()
        sel, coef, nnz, res, its = b.download(iters)
    tf = 2.0 * M * N * B * n / corr_ms / 1e9
    out = dict(config="c5 mp 4096x65536 f64", iterations=iters, signals=B, solves_per_s=B / (ms * 1e-3),
               ms_per_solve_batch=ms, corr_ms=corr_ms, corr_tflops=tf, frac_of_fp64_peak=tf / FP64_PEAK,
               corr_share=corr_ms / ms, median_resnorm=float(np.median(res)))
    if args.cpu:
        from oracle import pursuit_oracle as po
        out["cpu_baseline"] = cpu_oracle_c("mp", A, Bm, 16, iters) or cpu_oracle(lambda s: po.mp(A, Bm[:, s], iters), 2)
    return out


def c4(cs, dev, args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    M, k = 8192, int(128 * args.scale)
    N_loc = 131072                                    # 4 GiB FP32 per GPU: 8 ranks = the 1 048 576-atom dictionary
    N = N_loc * world
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        uid = cs.exchange_unique_id(dist, rank)
    else:
        uid = cs.ShardComm.unique_id()
    comm = cs.ShardComm(uid, rank, world, local)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    A_t = torch.empty(N_loc, M, dtype=torch.float32, device=dev)
    for n0 in range(0, N_loc, 16384):
        blk = torch.randn(16384, M, dtype=torch.float32, device=dev, generator=g)
        blk /= blk.norm(dim=1, keepdim=True)
        A_t[n0:n0 + 16384] = blk
    # planted signal: k atoms spread over all shards (every rank contributes k/world of its own atoms)
    per = max(1, k // world)
    gi = torch.Generator(device=dev).manual_seed(7)
    loc_idx = torch.randperm(N_loc, device=dev, generator=gi)[:per]
    part = A_t[loc_idx].to(torch.float64).sum(dim=0)
    if world > 1:
        dist.all_reduce(part)
    b = part.to(torch.float32).cpu().numpy()
    A_np = A_t.cpu().numpy().T
    del A_t
    torch.cuda.empty_cache()
    shard = cs.Dictionary(A_np, device=local, n_offset=rank * N_loc, n_total=N)
    del A_np
    cs.omp_sharded(shard, comm, b, k)                 # warm-up: the same call (sizes the communicator's scratch, sets up the exchange)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, info = cs.omp_sharded(shard, comm, b, k)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    tmax = torch.tensor([wall, info["corr_ms"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    wall, corr_ms = float(tmax[0]), float(tmax[1])
    planted = set((loc_idx.cpu().numpy() + rank * N_loc).tolist())
    mine = len(planted & set(x.nzind.tolist()))
    found = torch.tensor([mine], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(found)
    gbs = N_loc * M * 4 * k / (corr_ms * 1e-3) / 1e9      # per-GPU shard bytes per correlation pass / its device time
    out = dict(config=f"c4 omp 8192x{N} f32 k={k}, column-sharded over {world} GPU(s)", solves_per_s=1.0 / wall,
               wall_s=wall, corr_ms_total=corr_ms, gemv_GBps_per_gpu=gbs, frac_of_hbm_peak=gbs / HBM_PEAK,
               hbm_peak_GBps=HBM_PEAK, gemv_share_of_wall=corr_ms * 1e-3 / wall, planted_atoms_found=int(found.item()),
               planted_atoms=per * world, resnorm=info["resnorm"], exchange=info["exchange"])
    comm.close(); shard.close()
    if world > 1:
        dist.destroy_process_group()
    return out if rank == 0 else None


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["c1", "c2s", "c3", "c4", "c5", "fr2", "sp2"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle on a bounded sample (reported baseline)")
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cs = ge.load_package()
    res = {"c1": c1, "c2s": c2s, "c3": c3, "c4": c4, "c5": c5, "fr2": fr2, "sp2": sp2}[a.config](cs, dev, a)
    if res is not None:
        print(json.dumps(res))
