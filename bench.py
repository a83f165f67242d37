#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's metric: OMP solves/sec at 1024x8192, k=32, FP64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one batched OMP solve (k = 32 update!s) of `--signals` right-hand sides per GPU
(BASELINE config 2: 65 536).  Independent signals are sharded over ranks with no data-path
collective ("scaling": "weak": every rank solves its own 65 536 signals against a replicated dictionary).

One JSON line is printed by rank 0:
  value      whole-job solves/s, signals already resident in HBM, device-timed (CUDA events on the
             library's stream, max over ranks)
  e2e        the same metric through the C-ABI one-shot call `csb200_omp` with pinned HOST buffers:
             batch allocation, H2D of the signals, the solve and D2H of the results inside the timed region
  roofline   FP64 tensor (DMMA) roofline of the dominant kernel (the correlation GEMM): algorithmic
             2*M*N*B flop per launch / its mean launch duration, timed live with CUDA events
  cpu_baseline  the CPU oracle (plain-C restatement of the reference, one signal per host thread; the NumPy oracle
             if it cannot be built) on a bounded sample
`--impl reference` times that CPU restatement alone (Julia is not installed in this image, so the
reference itself cannot run).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M, N, K_SPARSE = 1024, 8192, 32
METRIC = "OMP solves/sec at 1024x8192,k=32 FP64"
UNIT = "solves/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--signals", type=int, default=65536, help="signals per GPU per step")
    ap.add_argument("--ref-signals", type=int, default=512, help="signals per step of the CPU reference arm")
    ap.add_argument("--cpu-signals", type=int, default=1024, help="signals of the cpu_baseline sample (0 = skip)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed e2e steps (0 = same as --steps)")
    return ap.parse_args()


def fp64_peak():
    """Measured DMMA peak (tools/fp64_peak.cu run on this pool's B200, committed under profiles/)."""
    path = os.path.join(ROOT, "profiles", "FP64_PEAK.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["peak_tflops"]), d.get("source", "profiles/FP64_PEAK.json")
    return 37.0, "nominal 37 TFLOP/s (no measured FP64 peak committed)"


# ------------------------------------------------------------------------------------ CPU arms
def host_workload(nsig, seed):
    from oracle import pursuit_oracle as po
    A = po.gaussian_dictionary(np.random.default_rng(1234), M, N)
    rng = np.random.default_rng(seed)
    cols = []
    for _ in range(nsig):
        x0 = po.sparse_vector(rng, N, K_SPARSE)
        cols.append(A[:, x0.nzind] @ np.asarray(x0.nzval))
    return A, np.asfortranarray(np.stack(cols, axis=1))


def time_oracle(A, Bm):
    """Seconds the CPU restatement of the reference needs for the columns of Bm, with all host threads, and how it
    ran: the plain-C oracle (oracle/pursuit_oracle.c, one signal per thread) when it can be built, else the NumPy
    oracle (signals in sequence, multi-threaded OpenBLAS gemv)."""
    try:
        from oracle import c_oracle
        c_oracle.lib()
        t0 = time.perf_counter()
        got = c_oracle.solve_batch("omp", A, Bm, K_SPARSE)
        return time.perf_counter() - t0, "plain-C oracle, one signal per thread", int(got["threads"])
    except Exception as exc:                                   # no gcc / make on this host
        sys.stderr.write(f"C oracle unavailable ({exc}); timing the NumPy oracle\n")
    from oracle import pursuit_oracle as po
    t0 = time.perf_counter()
    for s in range(Bm.shape[1]):
        po.omp(A, Bm[:, s], K_SPARSE)
    return time.perf_counter() - t0, "NumPy/OpenBLAS oracle, multi-threaded gemv", os.cpu_count()


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ns = args.ref_signals
    A, Bm = host_workload(ns * (args.steps + args.warmup), 999)
    for w in range(args.warmup):
        time_oracle(A, Bm[:, w * ns:(w + 1) * ns])
    t, how, cores = 0.0, "", os.cpu_count()
    for s in range(args.warmup, args.warmup + args.steps):
        dt, how, cores = time_oracle(A, Bm[:, s * ns:(s + 1) * ns])
        t += dt
    value = ns * args.steps / t
    sample = f"{ns} signals per step x {args.steps} steps of the 65536-signal workload, {how}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "batched omp 1024x8192 f64 k=32 (BASELINE config 2), bounded sample", "M": M, "N": N,
                   "k": K_SPARSE, "signals_per_step": ns},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "Julia is absent from this image: the reference arm is the line-by-line CPU restatement (oracle/) of the "
                "reference's omp on all host cores",
    }))


# ------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        out, _ = self.proc.communicate(timeout=10)
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def ours_arm(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cs = ge.load_package()          # raises if libcsb200.so is missing: no fallback
    B, k = args.signals, K_SPARSE
    eps = float(np.finfo(np.float64).eps)

    # synthetic inputs (src/util.jl:21-31 distribution): Gaussian, eps-mean-shifted, unit-norm atoms; planted
    # k-sparse +-1 signals.  Row j of A_t is atom j, i.e. the memory image of a column-major M x N matrix.
    g = torch.Generator(device=dev).manual_seed(1234)
    A_t = torch.randn(N, M, dtype=torch.float64, device=dev, generator=g)
    A_t -= 1e-6 * A_t.mean(dim=1, keepdim=True)
    A_t /= A_t.norm(dim=1, keepdim=True)
    A_np = A_t.cpu().numpy().T                                   # (M, N) Fortran-ordered view
    g2 = torch.Generator(device=dev).manual_seed(5678 + rank)
    idx = torch.empty(B, k, dtype=torch.int64, device=dev)
    for s0 in range(0, B, 8192):
        s1 = min(B, s0 + 8192)
        idx[s0:s1] = torch.rand(s1 - s0, N, device=dev, generator=g2).topk(k, dim=1).indices
    sign = torch.randint(0, 2, (B, k), device=dev, generator=g2).to(torch.float64) * 2 - 1
    B_t = torch.empty(B, M, dtype=torch.float64, device=dev)
    for s0 in range(0, B, 2048):
        s1 = min(B, s0 + 2048)
        B_t[s0:s1] = (A_t[idx[s0:s1]] * sign[s0:s1, :, None]).sum(dim=1)
    B_pin = torch.empty(B, M, dtype=torch.float64, pin_memory=True)
    B_pin.copy_(B_t)
    torch.cuda.synchronize()
    B_np = B_pin.numpy().T                                       # (M, B) Fortran-ordered view of pinned memory
    idx_sorted = idx.sort(dim=1).values.cpu().numpy()
    del B_t, A_t
    torch.cuda.empty_cache()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    D = cs.Dictionary(A_np, device=local)
    batch = cs.Batch(D, B, k)
    batch.upload(B_np)                                           # resident in HBM before the timed region

    # ---- device-resident throughput ("value") ----
    for _ in range(args.warmup):
        batch.omp(k, eps)
    sampler = ClockSampler(local)
    batch.profile(True)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        batch.omp(k, eps)
        dev_ms += batch.last_solve_ms()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    corr_ms, corr_launches, other_launches = batch.corr_time()
    batch.profile(False)
    dev_ms_max = maxr(dev_ms)
    wall_max = maxr(wall)
    value = world * B * args.steps / (dev_ms_max * 1e-3)

    sel, coef, nnz, res, its = batch.download(k)
    recovered = float(np.mean((np.sort(sel, axis=1) == idx_sorted).all(axis=1)))
    max_res = float(res.max())
    batch.close()

    # ---- end to end through the C ABI with host buffers ("e2e") ----
    from ctypes import POINTER, c_double, c_int64
    o_sel = np.empty((B, k), dtype=np.int64); o_coef = np.empty((B, k)); o_nnz = np.empty(B, dtype=np.int64)
    o_res = np.empty(B); o_it = np.empty(B, dtype=np.int64)
    i64p, f64p = POINTER(c_int64), POINTER(c_double)

    def e2e_once():
        rc = cs.lib.csb200_omp(D._h, B_np.ctypes.data, M, B, k, eps, o_sel.ctypes.data_as(i64p),
                               o_coef.ctypes.data_as(f64p), o_nnz.ctypes.data_as(i64p), o_res.ctypes.data_as(f64p),
                               o_it.ctypes.data_as(i64p))
        if rc != 0:
            raise RuntimeError(f"csb200_omp failed: {rc} {cs.lib.csb200_last_error().decode()}")

    e2e_steps = args.e2e_steps or args.steps
    e2e_once()                                                   # warm-up (allocator, page-in)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_once()
    barrier()
    e2e_t = maxr(time.perf_counter() - t0)
    e2e_value = world * B * e2e_steps / e2e_t
    e2e_ok = float(np.mean((np.sort(o_sel, axis=1) == idx_sorted).all(axis=1)))
    h2d = M * B * 8
    d2h = B * (4 + 4 * k + 8 * k + 8 + 4)

    # ---- roofline of the dominant kernel ----
    peak, peak_src = fp64_peak()
    flop_per_launch = 2.0 * M * N * B
    achieved = flop_per_launch * corr_launches / (corr_ms * 1e-3) / 1e12 if corr_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "corr_gemm_traffic.json")
    if os.path.exists(tpath) and B == 65536:          # the ncu capture was taken at exactly this shape
        traffic = json.load(open(tpath))["dram_bytes_per_launch"]
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_unit": "bytes of DRAM read+write per launch (ncu, profiles/corr_gemm_traffic.json)", "kernel": "corr_gemm_f64_kernel", "launches": int(corr_launches),
                "mean_launch_ms": corr_ms / max(1, corr_launches), "share_of_step": corr_ms / dev_ms if dev_ms else None,
                "flop_per_launch": flop_per_launch, "peak_source": peak_src}

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and args.cpu_signals > 0:
        ns = args.cpu_signals
        t, how, cores = time_oracle(A_np, B_np[:, :ns])
        cpu = {"value": ns / t, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {ns} of the {B} signals of this step, {how} ({t:.1f} s)"}

    D.close()
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batched omp 1024x8192 f64 k=32, 65536 signals per GPU (BASELINE config 2)", "M": M,
                       "N": N, "k": k, "signals_per_gpu": B, "global_signals": world * B,
                       "parallelism": f"signals sharded over {world} GPU(s), dictionary replicated, no collective",
                       "l2": "inputs exceed L2: signals + residuals = 2 x %d MiB per solve vs 126 MB L2" % (h2d >> 20),
                       "e2e_api": "csb200_omp (C ABI one-shot, pinned host buffers)"},
            "wall_ms_per_step": 1e3 * wall_max / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "support_recovered_frac": e2e_ok},
            "gpu_launches": int((corr_launches + other_launches + args.steps)),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "check": {"support_recovered_frac": recovered, "max_resnorm": max_res},
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours_arm(a)
