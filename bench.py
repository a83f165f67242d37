#!/usr/bin/env python
"""Benchmark of the greedy-pursuit hot path on BASELINE.json's metric and configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c1|c3|c4|c5|c2s]
                    [--secondary auto|none|c1,c3,...] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default (`--config c2`, BASELINE config 2 = the configuration the metric is quoted on): a "step" is one batched OMP
solve (k = 32 `update!`s) of 65 536 right-hand sides per GPU against the replicated 1024 x 8192 FP64 dictionary.
Independent signals are sharded over ranks with no data-path collective.  Rank 0 prints ONE JSON line:

  value      whole-job solves/s, signals already resident in HBM, device-timed (CUDA events on the library's stream,
             max over ranks); "scaling": "weak" = 65 536 signals PER GPU
  strong     (N > 1) the same measurement with BASELINE config 2's 65 536 signals SPLIT over the N GPUs
  e2e        the same metric through the C-ABI one-shot call `csb200_omp` with pinned HOST buffers: batch allocation,
             H2D of the signals, the solve and D2H of the results inside the timed region
  roofline   FP64 tensor (DMMA) roofline of the dominant kernel (the correlation GEMM): algorithmic 2*M*N*B flop per
             launch / its mean launch duration, timed live with CUDA events; the FP64 peak is re-measured inside the
             same clock-sampling window (tools/fp64_peak --quick)
  cpu_baseline  the CPU oracle (plain-C restatement of the reference, one signal per host thread) on a bounded sample
  check.oracle_parity  that sample's oracle results against the GPU results for the same signals: selection order
             exact, coefficients / residual norms within 1e-10
  secondary  the other BASELINE configs measured in the same run with the same schema (value, e2e, roofline,
             cpu_baseline, check): c1 (128 x 256 single signal), c2s (single signal on the config-2 dictionary, L2
             regime), c3 (gomp), c5 (mp) at N = 1; c4 (column-sharded single-signal omp, one 4 GiB FP32 shard per
             GPU, per-iteration exchange over NVLink) at every N.
`--config X` makes X the headline line instead.  `--impl reference` times the CPU restatement alone (Julia is not
installed in this image, so the reference itself cannot run).
Inputs: NumPy PCG64 streams (dictionary seed 1234, signals seed 5678 + config id) for c1 / c2; seeded torch
generators on the device for the large dictionaries of c3 / c4 / c5.  The oracle always sees the same bytes.
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
EPS64 = float(np.finfo(np.float64).eps)


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c1", "c2s", "c3", "c4", "c5"])
    ap.add_argument("--secondary", default="auto",
                    help="other configs measured into the line's `secondary` object: auto | none | comma list")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="c2 headline: 65536 signals per GPU (weak) or 65536 split over the GPUs (strong); the other "
                         "one is reported in the `strong` / `weak` sub-object")
    ap.add_argument("--signals", type=int, default=65536, help="c2: signals per GPU per step (weak) / in total (strong)")
    ap.add_argument("--ref-signals", type=int, default=512, help="signals per step of the CPU reference arm")
    ap.add_argument("--cpu-signals", type=int, default=1024, help="signals of the cpu_baseline sample (0 = skip)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed e2e steps (0 = same as --steps)")
    ap.add_argument("--fp64-steps", type=int, default=2,
                    help="c2: timed steps of the same workload through the FP64 DMMA pass (CSB200_SCREEN=0), reported as `fp64_path` (0 = skip)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the secondary configs (smoke runs)")
    return ap.parse_args(argv)


# ------------------------------------------------------------------------------------ peaks
def fp64_peak_committed():
    """Measured DMMA peak (tools/fp64_peak.cu run on this pool's B200, committed under profiles/)."""
    path = os.path.join(ROOT, "profiles", "FP64_PEAK.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["peak_tflops"]), d.get("source", "profiles/FP64_PEAK.json")
    return 37.0, "nominal 37 TFLOP/s (no measured FP64 peak committed)"


def fp64_peak_live(device):
    """tools/fp64_peak --quick on `device`: the register-tiled DMMA rate, measured on this box in this run."""
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    if not os.path.exists(exe):
        return None
    try:
        r = subprocess.run([exe, "--quick", str(device)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                           timeout=60)
        if r.returncode != 0:
            return None
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        return None


def tf32_peak():
    """Dense TF32 tensor peak: half of the cuBLAS bf16 figure the driver measured on this pool (MEASURED_PEAKS.json)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["bf16_tflops"]) / 2.0, ("MEASURED_PEAKS.json bf16_tflops (cuBLAS bf16 burst, driver-written) / 2: tcgen05 "
                                               "kind::tf32 runs at half the kind::f16 rate; sustained figure / 2 = %.0f"
                                               % (float(d.get("bf16_tflops_sustained", 0.0)) / 2.0))
    return 1125.0, "fallback: nominal dense TF32 1.1 PFLOP/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def f16_peak():
    """Dense FP16/BF16 tensor peak: the cuBLAS bf16 figure the driver measured on this pool (MEASURED_PEAKS.json)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["bf16_tflops"]), ("MEASURED_PEAKS.json bf16_tflops (cuBLAS bf16 burst, driver-written): tcgen05 kind::f16; "
                                         "sustained figure %.0f" % float(d.get("bf16_tflops_sustained", 0.0)))
    return 2250.0, "fallback: nominal dense FP16 2.25 PFLOP/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def f16_gemm_live(torch, dev):
    """cuBLAS FP16 GEMM (torch.matmul) 8192^3 on this box in this run: the library yardstick for the FP16 screening pass."""
    try:
        a = torch.randn(8192, 8192, device=dev, dtype=torch.float16)
        b = torch.randn(8192, 8192, device=dev, dtype=torch.float16)
        for _ in range(3):
            a @ b
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(5):
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b
        return {"tflops": 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12, "what": "torch.matmul fp16, 8192^3, best of 5 (CUDA events)"}
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"}


def tf32_gemm_live(torch, dev):
    """cuBLAS TF32 GEMM (torch.matmul, allow_tf32) 8192^3 on this box in this run: the library yardstick for the screening pass."""
    try:
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(8192, 8192, device=dev, dtype=torch.float32)
        b = torch.randn(8192, 8192, device=dev, dtype=torch.float32)
        for _ in range(3):
            a @ b
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(5):
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b
        return {"tflops": 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12, "what": "torch.matmul fp32 with allow_tf32, 8192^3, best of 5 (CUDA events)"}
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured: MEASURED_PEAKS.json hbm_gbs (device copy, read+write)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


# ------------------------------------------------------------------------------------ synthetic inputs (host, NumPy)
def gaussian_dictionary_np(rng, m, n, dtype=np.float64):
    """`sparse_data`'s dictionary (src/util.jl:21-28): Gaussian, 1e-6 x mean shifted, unit-norm columns; (m, n) Fortran."""
    A = rng.standard_normal((m, n))
    A -= 1e-6 * A.mean(axis=0, keepdims=True)
    A /= np.sqrt((A * A).sum(axis=0, keepdims=True))
    return np.asfortranarray(A.astype(dtype))


def draw_supports_np(rng, nsig, n, k):
    """k distinct atom indices per signal, uniform over k-subsets (rejection of rows with a repeat), and +-1 signs
    (`sparse_vector`, src/util.jl:13-19)."""
    idx = rng.integers(0, n, size=(nsig, k))
    for _ in range(8):
        srt = np.sort(idx, axis=1)
        bad = np.nonzero((srt[:, 1:] == srt[:, :-1]).any(axis=1))[0]
        if bad.size == 0:
            break
        idx[bad] = rng.integers(0, n, size=(bad.size, k))
    else:                                   # k*k is not small against n: draw the remaining rows one by one
        for s in bad:
            idx[s] = rng.choice(n, size=k, replace=False)
    sign = rng.integers(0, 2, size=(nsig, k)).astype(np.float64) * 2.0 - 1.0
    return idx, sign


def host_workload(nsig, seed):
    """Config-2 inputs on the host (reference arm): dictionary seed 1234, planted k-sparse +-1 signals."""
    A = gaussian_dictionary_np(np.random.default_rng(1234), M, N)
    idx, sign = draw_supports_np(np.random.default_rng(seed), nsig, N, K_SPARSE)
    Bm = np.empty((M, nsig), order="F")
    for s in range(nsig):
        Bm[:, s] = A[:, idx[s]] @ sign[s]
    return A, Bm


# ------------------------------------------------------------------------------------ CPU arms (oracle = checker / baseline only)
def run_c_oracle(algo, A, Bm, k, l=1, eps=None, threads=0):
    """(seconds, results dict, description, threads) of the plain-C oracle on the columns of Bm; None if unavailable."""
    try:
        from oracle import c_oracle
        c_oracle.lib()
    except Exception as exc:                                   # no gcc / make on this host
        sys.stderr.write(f"C oracle unavailable ({exc})\n")
        return None
    t0 = time.perf_counter()
    got = c_oracle.solve_batch(algo, A, Bm, k, l=l, eps=eps, threads=threads)
    return time.perf_counter() - t0, got, "plain-C oracle, one signal per thread", int(got["threads"])


def time_oracle(A, Bm):
    """Seconds the CPU restatement of the reference needs for the columns of Bm (config 2 omp), how it ran, on how many
    threads, and its results (None for the NumPy fallback)."""
    r = run_c_oracle("omp", A, Bm, K_SPARSE)
    if r is not None:
        return r[0], r[2], r[3], r[1]
    from oracle import pursuit_oracle as po
    t0 = time.perf_counter()
    for s in range(Bm.shape[1]):
        po.omp(A, Bm[:, s], K_SPARSE)
    return time.perf_counter() - t0, "NumPy/OpenBLAS oracle, multi-threaded gemv", os.cpu_count(), None


def parity_vs_c_oracle(got, sel, coef, nnz, res, Bm, tol=1e-10):
    """Oracle results `got` (c_oracle.solve_batch) against GPU outputs in selection order (sel, coef, nnz, res) for
    the same signals: selection sequence exact; coefficients (matched by atom) and residual norms within tol."""
    ns = sel.shape[0]
    kk = min(sel.shape[1], got["order"].shape[1])
    order_ok = (got["order"][:ns, :kk] == sel[:, :kk]).all(axis=1) & (got["nnz"][:ns] == nnz[:ns])
    coef_err, res_err = 0.0, 0.0
    for s in range(ns):
        t = int(nnz[s])
        o = np.argsort(sel[s, :t], kind="stable")
        gi, gv = sel[s, :t][o], coef[s, :t][o]
        n = int(got["nnz"][s])
        if n != t or not np.array_equal(gi, got["nzind"][s, :n]):
            coef_err = float("inf")
            continue
        ref = got["nzval"][s, :n]
        coef_err = max(coef_err, float(np.max(np.abs(gv - ref)) / max(np.max(np.abs(ref)), 1e-300)) if n else 0.0)
        nb = float(np.linalg.norm(Bm[:, s]))
        res_err = max(res_err, abs(float(res[s]) - float(got["resnorm"][s])) / max(nb, 1e-300))
    ok = bool(order_ok.all()) and coef_err <= tol and res_err <= tol
    return {"signals": int(ns), "selection_order_exact_frac": float(order_ok.mean()), "coef_max_rel_err": coef_err,
            "resnorm_max_abs_err_over_norm_b": res_err, "tolerance": tol, "pass": ok}


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
        dt, how, cores, _ = time_oracle(A, Bm[:, s * ns:(s + 1) * ns])
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


class Ctx:
    """Rank / device / library handles shared by the config functions."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import __graft_entry__ as ge
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.cs = ge.load_package()          # raises if libcsb200.so is missing: no fallback
        self.comm = None                     # column-shard communicator (created on first use, c4)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def maxr(self, *xs):
        if self.world == 1:
            return xs if len(xs) > 1 else xs[0]
        t = self.torch.tensor(list(xs), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        out = tuple(float(v) for v in t.tolist())
        return out if len(xs) > 1 else out[0]

    def sumr(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t)
        return float(t.item())

    def shard_comm(self):
        if self.comm is None:
            uid = self.cs.exchange_unique_id(self.dist, self.rank) if self.world > 1 else self.cs.ShardComm.unique_id()
            self.comm = self.cs.ShardComm(uid, self.rank, self.world, self.local)
        return self.comm

    def close(self):
        if self.comm is not None:
            self.comm.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def draw_supports_torch(torch, nsig, n, k, gen, dev):
    """Device version of draw_supports_np (same distribution, torch generator)."""
    idx = torch.randint(0, n, (nsig, k), device=dev, generator=gen)
    for _ in range(8):
        srt = idx.sort(dim=1).values
        bad = (srt[:, 1:] == srt[:, :-1]).any(dim=1).nonzero().flatten()
        if bad.numel() == 0:
            break
        idx[bad] = torch.randint(0, n, (bad.numel(), k), device=dev, generator=gen)
    else:                                   # k*k is not small against n: the k smallest of n uniform keys per row
        for s0 in range(0, bad.numel(), 1024):
            rows = bad[s0:s0 + 1024]
            idx[rows] = torch.rand(rows.numel(), n, device=dev, generator=gen).topk(k, dim=1).indices
    sign = torch.randint(0, 2, (nsig, k), device=dev, generator=gen).to(torch.float64) * 2 - 1
    return idx, sign


def planted_signals_torch(torch, A_t, idx, sign):
    """b_s = sum_j sign[s, j] * A[:, idx[s, j]] for every signal, FP64 on the device (A_t is atoms x M)."""
    B, k = idx.shape
    Mr = A_t.shape[1]
    B_t = torch.empty(B, Mr, dtype=torch.float64, device=A_t.device)
    step = max(1, (1 << 26) // (k * Mr))
    for s0 in range(0, B, step):
        s1 = min(B, s0 + step)
        B_t[s0:s1] = (A_t[idx[s0:s1]].to(torch.float64) * sign[s0:s1, :, None]).sum(dim=1)
    return B_t


def make_problem(ctx, Mr, Nc, k, B, dtype, seed=1234, noise=0.0):
    """Large synthetic problem generated on the device with a seeded torch generator (Gaussian unit-norm atoms, planted
    k-sparse +-1 signals with DISTINCT atoms, optional noise of norm `noise`); returns host arrays (Fortran views)."""
    torch, dev = ctx.torch, ctx.dev
    g = torch.Generator(device=dev).manual_seed(seed)
    A_t = torch.empty(Nc, Mr, dtype=torch.float64, device=dev)
    for n0 in range(0, Nc, 65536):
        n1 = min(Nc, n0 + 65536)
        blk = torch.randn(n1 - n0, Mr, dtype=torch.float64, device=dev, generator=g)
        blk -= 1e-6 * blk.mean(dim=1, keepdim=True)
        blk /= blk.norm(dim=1, keepdim=True)
        A_t[n0:n1] = blk
    idx, sign = draw_supports_torch(torch, B, Nc, k, g, dev)
    B_t = planted_signals_torch(torch, A_t, idx, sign)
    if noise:
        e = torch.randn(B, Mr, dtype=torch.float64, device=dev, generator=g)
        B_t += e * (noise / e.norm(dim=1, keepdim=True))
    td = torch.float32 if dtype == np.float32 else torch.float64
    A_np = A_t.to(td).cpu().numpy().T
    B_np = B_t.to(td).cpu().numpy().T
    idx_np = idx.cpu().numpy()
    del A_t, B_t
    torch.cuda.empty_cache()
    return A_np, B_np, idx_np


def supports_recovered(sel, nnz, idx):
    """Fraction of signals whose planted support is contained in the returned one."""
    ok = 0
    for s in range(sel.shape[0]):
        ok += set(idx[s].tolist()) <= set(sel[s, :int(nnz[s])].tolist())
    return ok / max(1, sel.shape[0])


# ------------------------------------------------------------------------------------ config 2 (headline)
class C2:
    """BASELINE config 2: batched omp, 1024 x 8192 FP64 dictionary (NumPy PCG64 seed 1234), k = 32."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.A_np = gaussian_dictionary_np(np.random.default_rng(1234), M, N)
        self.D = ctx.cs.Dictionary(self.A_np, device=ctx.local)
        self.A_t = ctx.torch.from_numpy(np.ascontiguousarray(self.A_np.T)).to(ctx.dev)       # atoms x M

    def signals(self, B):
        """B planted signals of this rank (NumPy PCG64 seed [5678 + 2, rank]) in pinned host memory."""
        ctx, torch = self.ctx, self.ctx.torch
        idx, sign = draw_supports_np(np.random.default_rng([5678 + 2, ctx.rank]), B, N, K_SPARSE)
        B_t = planted_signals_torch(torch, self.A_t, torch.from_numpy(idx).to(ctx.dev), torch.from_numpy(sign).to(ctx.dev))
        B_pin = torch.empty(B, M, dtype=torch.float64, pin_memory=True)
        B_pin.copy_(B_t)
        torch.cuda.synchronize()
        del B_t
        return B_pin, np.sort(idx, axis=1)

    def measure(self, B, steps, warmup, e2e_steps, with_peak=False, cpu_signals=0, screen=None):
        """Device-resident throughput, e2e through csb200_omp, roofline of the correlation kernel, optional CPU sample +
        parity.  screen: None = the library's default path (TF32 screening + exact FP64 re-evaluation for large batches),
        False = the FP64 DMMA pass (CSB200_SCREEN=0)."""
        old_env = os.environ.get("CSB200_SCREEN")
        if screen is not None:
            os.environ["CSB200_SCREEN"] = "1" if screen else "0"
        try:
            return self._measure(B, steps, warmup, e2e_steps, with_peak, cpu_signals)
        finally:
            if screen is not None:
                if old_env is None:
                    os.environ.pop("CSB200_SCREEN", None)
                else:
                    os.environ["CSB200_SCREEN"] = old_env

    def _measure(self, B, steps, warmup, e2e_steps, with_peak, cpu_signals):
        ctx, cs, k = self.ctx, self.ctx.cs, K_SPARSE
        B_pin, idx_sorted = self.signals(B)
        B_np = B_pin.numpy().T                                   # (M, B) Fortran-ordered view of pinned memory
        batch = cs.Batch(self.D, B, k)
        batch.upload(B_np)                                       # resident in HBM before the timed region
        for _ in range(warmup):
            batch.omp(k, EPS64)
        sampler = ClockSampler(ctx.local)
        batch.profile(True)
        ctx.barrier()
        sampler.start()
        t0 = time.perf_counter()
        dev_ms = 0.0
        for _ in range(steps):
            batch.omp(k, EPS64)
            dev_ms += batch.last_solve_ms()
        ctx.barrier()
        wall = time.perf_counter() - t0
        scr = batch.screen_stats(reset=True)
        screened = scr["path_id"] in (3, 4)
        f16 = scr["path_id"] == 4
        live = None
        if with_peak and ctx.rank == 0:                                                   # inside the clocks window
            live = ((f16_gemm_live if f16 else tf32_gemm_live)(ctx.torch, ctx.dev)) if screened else fp64_peak_live(ctx.local)
        ctx.barrier()
        clocks = sampler.stop()
        corr_ms, corr_launches, other_launches = batch.corr_time()
        batch.profile(False)
        dev_ms_max, wall_max = ctx.maxr(dev_ms, wall)
        value = ctx.world * B * steps / (dev_ms_max * 1e-3)
        sel, coef, nnz, res, its = batch.download(k)
        recovered = float(np.mean((np.sort(sel, axis=1) == idx_sorted).all(axis=1)))
        max_res = float(res.max())
        import hashlib
        digest = hashlib.sha256(sel.tobytes() + coef.tobytes() + nnz.tobytes()).hexdigest()[:16]   # supports in selection order + coefficients
        batch.close()

        # ---- end to end through the C ABI with host buffers ----
        from ctypes import POINTER, c_double, c_int64
        o_sel = np.empty((B, k), dtype=np.int64); o_coef = np.empty((B, k)); o_nnz = np.empty(B, dtype=np.int64)
        o_res = np.empty(B); o_it = np.empty(B, dtype=np.int64)
        i64p, f64p = POINTER(c_int64), POINTER(c_double)

        def e2e_once():
            rc = cs.lib.csb200_omp(self.D._h, B_np.ctypes.data, M, B, k, EPS64, o_sel.ctypes.data_as(i64p),
                                   o_coef.ctypes.data_as(f64p), o_nnz.ctypes.data_as(i64p), o_res.ctypes.data_as(f64p),
                                   o_it.ctypes.data_as(i64p))
            if rc != 0:
                raise RuntimeError(f"csb200_omp failed: {rc} {cs.lib.csb200_last_error().decode()}")

        e2e_once()                                               # warm-up (allocator, page-in)
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_once()
        ctx.barrier()
        e2e_t = ctx.maxr(time.perf_counter() - t0)
        e2e_value = ctx.world * B * e2e_steps / e2e_t
        e2e_ok = float(np.mean((np.sort(o_sel, axis=1) == idx_sorted).all(axis=1)))
        e2e_same = bool(np.array_equal(o_sel, sel) and np.array_equal(o_coef, coef))    # one-shot == resident path
        h2d = M * B * 8
        d2h = B * (4 + 4 * k + 8 * k + 8 + 4)
        cs.lib.csb200_dict_trim(self.D._h)

        # ---- roofline of the correlation kernel ----
        # every update! of every signal is one column of a correlation pass, however the passes are cut into launches:
        # flop per launch = total / launches.  SURVEY 8(d): algorithmic work = 2 M N flop per signal-update.
        flop_total = 2.0 * M * N * B * k * steps
        flop_per_launch = flop_total / max(1, corr_launches)
        achieved = flop_total / (corr_ms * 1e-3) / 1e12 if corr_ms > 0 else 0.0
        share = corr_ms / dev_ms if dev_ms else None
        if screened:
            # TF32 screening pass (tcgen05): the flop EXECUTED are the algorithmic 2 M N B per pass (tiles 128 x 256 x K,
            # K padded to 32: no padding at this shape), counted at TF32 -- reported against the TF32 tensor peak.
            peak, peak_src = f16_peak() if f16 else tf32_peak()
            kpad = (M + 63) // 64 * 64 if f16 else (M + 31) // 32 * 32
            tiles = ((B + 127) // 128) * ((N + 255) // 256)
            l2_bytes = tiles * kpad * (2.0 if f16 else 4.0) * (128 + 256)            # operand bytes TMA pulls from L2 per pass
            nominal = 2250.0 if f16 else 1125.0
            passes = max(1, k * steps)
            roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                        "traffic": _traffic("corr_screen_traffic.json") if B == 65536 else None,
                        "traffic_unit": "bytes of DRAM read+write per launch (ncu, profiles/corr_screen_traffic.json)",
                        "kernel": "corr_screen_tf32_kernel<4, true>" if f16 else "corr_screen_tf32_kernel<4>",
                        "dtype": ("fp16 operands scaled by powers of two, f32 accumulation (tcgen05.mma kind::f16)" if f16
                                  else "tf32 operands, f32 accumulation (tcgen05.mma kind::tf32)"),
                        "launches": int(corr_launches), "mean_launch_ms": corr_ms / max(1, corr_launches), "share_of_step": share,
                        "launches_per_update": corr_launches / passes, "flop_per_launch": flop_per_launch,
                        "peak_source": peak_src, "peak_nominal": nominal, "frac_of_nominal": achieved / nominal,
                        "library_gemm_live": live,
                        "l2_operand_bytes_per_pass": l2_bytes,
                        "l2_to_sm_tbs": l2_bytes * passes / (corr_ms * 1e-3) / 1e12 if corr_ms > 0 else None,
                        "l2_to_sm_port_tbs": 148 * 64 * 1.965e9 / 1e12,
                        "note": ("FP16 operands halve the L2->SM bytes of the TF32 pass (which ran at the SMs' L2 read-port rate); "
                                 "what bounds the FP16 pass: see DESIGN 4.14" if f16 else
                                 "the pass is bound by the SMs' L2 read ports (64 B/clk/SM), not by the tensor pipe: see DESIGN 4.14")}
            # the rest of the step is omp_update_kernel (L2-gather-bound): its share and its gather rate
            upd_ms = dev_ms - corr_ms
            gather_bytes = sum((t + 3) for t in range(k)) * M * 8.0 * B * steps      # t active columns + a_j, b, r per update!
            roofline_update = {"kernel": "omp_append_warp_kernel + omp_residual_slice_kernel (+ omp_update_list_kernel for the few signals off the common path)",
                               "bound": "l2 gather", "share_of_step": upd_ms / dev_ms if dev_ms else None,
                               "ms_per_update": upd_ms / passes, "algorithmic_gather_bytes_per_step": gather_bytes / steps,
                               "achieved_tbs": gather_bytes / (upd_ms * 1e-3) / 1e12 if upd_ms > 0 else None}
            equiv = value / ctx.world * 2.0 * M * N * k / 1e12
            fp64_equiv = {"tflops_equivalent": equiv, "of_fp64_dmma_peak": equiv / fp64_peak_committed()[0],
                          "note": "solves/s x the reference algorithm's 2 M N k flop per solve: NOT a roofline fraction (the "
                                  "FP64 flop are not executed: SURVEY 8d), stated separately"}
        else:
            committed, committed_src = fp64_peak_committed()
            # denominator: the LARGER of the committed pool measurement and this run's own (the conservative choice; the
            # live figure proves the box at hand is not faster than the committed peak)
            if live and float(live["peak_tflops"]) > committed:
                peak, peak_src = float(live["peak_tflops"]), ("measured in this run, inside the clocks window: tools/fp64_peak "
                                                              "--quick (" + live["kernel"] + ")")
            else:
                peak, peak_src = committed, committed_src + ("; re-measured in this run inside the clocks window: %.2f TFLOP/s"
                                                             % float(live["peak_tflops"]) if live else "")
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "corr_gemm_traffic.json")
            if os.path.exists(tpath) and B == 65536:          # the ncu capture was taken at exactly this shape
                traffic = json.load(open(tpath))["dram_bytes_per_launch"]
            roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                        "traffic": traffic,
                        "traffic_unit": "bytes of DRAM read+write per launch (ncu, profiles/corr_gemm_traffic.json)",
                        "kernel": "corr_gemm_f64_kernel", "launches": int(corr_launches),
                        "mean_launch_ms": corr_ms / max(1, corr_launches), "share_of_step": share,
                        "launches_per_update": corr_launches / max(1, k * steps),
                        "flop_per_launch": flop_per_launch, "peak_source": peak_src, "peak_committed": committed,
                        "peak_live": live}
            roofline_update, fp64_equiv = None, None

        # ---- CPU baseline + oracle parity (rank 0, N = 1 only) ----
        cpu, parity = None, None
        if ctx.rank == 0 and ctx.world == 1 and cpu_signals > 0:
            ns = min(cpu_signals, B)
            t, how, cores, got = time_oracle(self.A_np, B_np[:, :ns])
            cpu = {"value": ns / t, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {ns} of the {B} signals of this step, {how} ({t:.1f} s)"}
            if got is not None:
                parity = parity_vs_c_oracle(got, sel[:ns], coef[:ns], nnz[:ns], res[:ns], B_np)
        return {
            "value": value, "ms_per_step": dev_ms_max / steps, "wall_ms_per_step": 1e3 * wall_max / steps,
            "signals_per_gpu": B,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "support_recovered_frac": e2e_ok, "bit_identical_to_resident_path": e2e_same},
            "gpu_launches": int(corr_launches + other_launches + steps),
            "roofline": roofline, "roofline_update": roofline_update, "fp64_equivalent": fp64_equiv,
            "cpu_baseline": cpu, "clocks": clocks,
            "check": {"support_recovered_frac": recovered, "max_resnorm": max_res, "oracle_parity": parity,
                      "result_digest": digest, "path": scr["path"], "screening": ({key: scr[key] for key in ("signal_updates", "candidates_reevaluated", "exact_scans")}
                                                         if screened else None)},
        }

    def close(self):
        self.D.close()
        del self.A_t


def run_c2(ctx, args, headline=True):
    c2 = C2(ctx)
    e2e_steps = args.e2e_steps or args.steps
    world = ctx.world
    B_weak = args.signals
    B_strong = max(1, args.signals // world)
    first, second = ("weak", "strong") if args.scaling == "weak" else ("strong", "weak")
    sizes = {"weak": B_weak, "strong": B_strong}
    main = c2.measure(sizes[first], args.steps, args.warmup, e2e_steps, with_peak=True, cpu_signals=args.cpu_signals)
    other = None
    if world > 1:                                           # at N = 1 the two coincide
        o = c2.measure(sizes[second], max(2, min(args.steps, 5)), 2, max(2, min(e2e_steps, 5)))
        other = {"scaling": second, "value": o["value"], "unit": UNIT, "ms_per_step": o["ms_per_step"],
                 "signals_per_gpu": o["signals_per_gpu"], "global_signals": o["signals_per_gpu"] * world,
                 "e2e": o["e2e"], "roofline_frac": o["roofline"]["frac"], "gemm_share_of_step": o["roofline"]["share_of_step"],
                 "check": o["check"]}
    fp64_path = None
    if args.fp64_steps > 0 and ("screening" in main["check"]["path"]):
        # the same workload through the FP64 DMMA pass (CSB200_SCREEN=0): what the screening buys, and the FP64 kernel's roofline
        o = c2.measure(sizes[first], args.fp64_steps, 1, 1, with_peak=True, screen=False)
        fp64_path = {"value": o["value"], "unit": UNIT, "ms_per_step": o["ms_per_step"], "e2e": o["e2e"], "roofline": o["roofline"],
                     "check": o["check"],
                     "bit_identical_to_screened_path": o["check"]["result_digest"] == main["check"]["result_digest"]}
    c2.close()
    B = main.pop("signals_per_gpu")
    line = {
        "metric": METRIC, "unit": UNIT, "higher_is_better": True, "scaling": first, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": ("batched omp 1024x8192 f64 k=32, 65536 signals per GPU (BASELINE config 2)" if first == "weak"
                                else "batched omp 1024x8192 f64 k=32, 65536 signals split over the GPUs (BASELINE config 2)"),
                   "M": M, "N": N, "k": K_SPARSE, "signals_per_gpu": B, "global_signals": world * B,
                   "parallelism": f"signals sharded over {world} GPU(s), dictionary replicated, no collective",
                   "l2": "inputs exceed L2: signals + residuals = 2 x %d MiB per solve vs 126 MB L2" % ((M * B * 8) >> 20),
                   "inputs": "NumPy PCG64: dictionary seed 1234, signals seed [5680, rank] (SURVEY 8d)",
                   "e2e_api": "csb200_omp (C ABI one-shot, pinned host buffers)"},
    }
    line.update(main)
    if other is not None:
        line[second] = other
    if fp64_path is not None:
        line["fp64_path"] = fp64_path
    return line


# ------------------------------------------------------------------------------------ config 1 (128 x 256, one signal)
def run_c1(ctx, args):
    cs = ctx.cs
    Mr, Nc, k = 128, 256, 8
    A = gaussian_dictionary_np(np.random.default_rng(1234), Mr, Nc)
    idx, sign = draw_supports_np(np.random.default_rng(5678 + 1), 64, Nc, k)
    Bm = np.empty((Mr, 64), order="F")
    for s in range(64):
        Bm[:, s] = A[:, idx[s]] @ sign[s]
    hbm, hbm_src = hbm_peak()
    reps = 200
    out = {}
    with cs.Dictionary(A, device=ctx.local) as D:
        res_single = None
        for name, nsig in [("single", 1), ("batch64", 64)]:
            with cs.Batch(D, nsig, k) as b:
                b.upload(Bm[:, :nsig])
                for _ in range(20):
                    b.omp(k, EPS64)
                ms = []
                for _ in range(reps):
                    b.omp(k, EPS64)
                    ms.append(b.last_solve_ms())
                out[name] = float(np.mean(ms)), float(np.median(ms))
                if nsig == 1:
                    res_single = b.download(k)
        b0 = np.ascontiguousarray(Bm[:, 0])
        for _ in range(20):
            cs.omp(D, b0, k)
        from ctypes import POINTER, c_double, c_int64
        i64p, f64p = POINTER(c_int64), POINTER(c_double)
        o_sel = np.empty((1, k), dtype=np.int64); o_coef = np.empty((1, k)); o_nnz = np.empty(1, dtype=np.int64)
        o_res = np.empty(1); o_it = np.empty(1, dtype=np.int64)
        t0 = time.perf_counter()
        for _ in range(reps):
            rc = cs.lib.csb200_omp(D._h, b0.ctypes.data, Mr, 1, k, EPS64, o_sel.ctypes.data_as(i64p),
                                   o_coef.ctypes.data_as(f64p), o_nnz.ctypes.data_as(i64p), o_res.ctypes.data_as(f64p),
                                   o_it.ctypes.data_as(i64p))
            if rc:
                raise RuntimeError(f"csb200_omp failed: {rc}")
        e2e_s = (time.perf_counter() - t0) / reps
    mean_ms, med_ms = out["single"]
    bytes_per_solve = Mr * Nc * 8 * k
    gbs = bytes_per_solve / (mean_ms * 1e-3) / 1e9
    cpu, parity = None, None
    if ctx.rank == 0 and ctx.world == 1 and args.cpu_signals > 0:
        r = None
        best = float("inf")
        for _ in range(30):                                       # one signal on ONE host thread: the latency a caller sees
            r = run_c_oracle("omp", A, Bm[:, :1], k, threads=1)
            if r is None:
                break
            best = min(best, r[0])
        if r is not None:
            cpu = {"value": 1.0 / best, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"the same single signal, plain-C oracle on one host thread, best of 30 ({1e6 * best:.1f} us per solve)",
                   "us_per_solve": 1e6 * best}
            sel, coef, nnz, res, its = res_single
            parity = parity_vs_c_oracle(r[1], sel, coef, nnz, res, Bm)
    return {
        "metric": "omp solves/sec at 128x256,k=8 FP64, single signal", "value": 1e3 / mean_ms, "unit": UNIT,
        "ms_per_step": mean_ms, "us_per_solve_device": 1e3 * mean_ms, "us_per_solve_device_median": 1e3 * med_ms,
        "us_per_solve_host_api": 1e6 * e2e_s, "batch64_us_per_call_device": 1e3 * out["batch64"][0],
        "higher_is_better": True, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "omp 128x256 f64 k=8, single signal (BASELINE config 1)", "M": Mr, "N": Nc, "k": k,
                   "timed_solves": reps, "inputs": "NumPy PCG64: dictionary seed 1234, signals seed 5679",
                   "l2": "256 KiB dictionary: L2/L1-resident by construction, latency-bound"},
        "e2e": {"value": 1.0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": Mr * 8, "d2h_bytes_per_step": 4 + 4 + 8 + k * 12,
                "steps": reps, "api": "csb200_omp (C ABI one-shot, host buffers)"},
        "gpu_launches": reps,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None,
                     "kernel": "small_solve_kernel / cluster solve (whole solve in one launch)",
                     "bytes_per_launch": bytes_per_solve, "peak_source": hbm_src,
                     "note": "8 sequential update!s over a 256 KiB dictionary: bounded by on-chip latency, not by HBM"},
        "cpu_baseline": cpu, "check": {"oracle_parity": parity},
    }


# ------------------------------------------------------------------------------------ c2s (single signal, config-2 dictionary)
def run_c2s(ctx, args):
    cs = ctx.cs
    k = K_SPARSE
    A = gaussian_dictionary_np(np.random.default_rng(1234), M, N)
    idx, sign = draw_supports_np(np.random.default_rng(5678 + 2), 4, N, k)
    Bm = np.empty((M, 4), order="F")
    for s in range(4):
        Bm[:, s] = A[:, idx[s]] @ sign[s]
    hbm, hbm_src = hbm_peak()
    res = {}
    primary = None
    for dt, name in [(np.float64, "f64"), (np.float32, "f32")]:
        Ad, Bd = np.asfortranarray(A.astype(dt)), np.asfortranarray(Bm.astype(dt))
        with cs.Dictionary(Ad, device=ctx.local) as D, cs.Batch(D, 1, k) as b:
            b.upload(Bd[:, :1])
            for _ in range(5):
                b.omp(k, 1e-30)
            b.profile(True)
            prof = []
            for _ in range(10):
                b.omp(k, 1e-30)
                prof.append(b.last_solve_ms())
            corr, nl, other = b.corr_time()
            b.profile(False)
            plain = []
            for _ in range(100):
                b.omp(k, 1e-30)
                plain.append(b.last_solve_ms())
            out = b.download(k)
            replays = b.graph_replays()
        ms = float(np.mean(plain))
        gemv_us = 1e3 * corr / max(1, nl)
        bytes_it = M * N * Ad.itemsize
        res[name] = {"us_per_solve": 1e3 * ms, "us_per_solve_median": 1e3 * float(np.median(plain)),
                     "us_per_iteration": 1e3 * ms / k, "GBps_whole_solve": bytes_it * k / (ms * 1e-3) / 1e9,
                     "frac_of_hbm_peak_whole_solve": bytes_it * k / (ms * 1e-3) / 1e9 / hbm,
                     "gemv_us_per_launch_profiled": gemv_us, "gemv_GBps_profiled": bytes_it / (gemv_us * 1e-6) / 1e9 if gemv_us else None,
                     "graph_replays": int(replays),
                     "support_recovered": bool(set(idx[0].tolist()) == set(out[0][0, :int(out[2][0])].tolist()))}
        if name == "f64":
            primary = (ms, out, Ad, Bd)
    ms, out, Ad, Bd = primary
    # 2..8 signals share ONE pass over the dictionary per update! (multi-right-hand-side workers of solve_persist.cu)
    idx8, sign8 = draw_supports_np(np.random.default_rng(5678 + 22), 8, N, k)
    B8 = np.asfortranarray(np.stack([A[:, idx8[s]] @ sign8[s] for s in range(8)], axis=1))
    multi = {}
    with cs.Dictionary(A, device=ctx.local) as D:
        for ns in (2, 4, 8):
            with cs.Batch(D, ns, k) as b:
                b.upload(B8[:, :ns])
                for _ in range(5):
                    b.omp(k, 1e-30)
                t8 = []
                for _ in range(50):
                    b.omp(k, 1e-30)
                    t8.append(b.last_solve_ms())
                sel8, _, nnz8, _, _ = b.download(k)
            ok8 = all(set(idx8[s].tolist()) == set(sel8[s, :int(nnz8[s])].tolist()) for s in range(ns))
            multi[str(ns)] = {"us_per_call": 1e3 * float(np.mean(t8)), "vs_one_signal": float(np.mean(t8)) / ms,
                              "supports_recovered": bool(ok8)}
    cpu, parity = None, None
    if ctx.rank == 0 and ctx.world == 1 and args.cpu_signals > 0:
        r = run_c_oracle("omp", Ad, Bd[:, :1], k, eps=1e-30, threads=1)
        if r is not None:
            cpu = {"value": 1.0 / r[0], "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"the same single signal, plain-C oracle on one host thread ({1e3 * r[0]:.1f} ms per solve)"}
            parity = parity_vs_c_oracle(r[1], out[0], out[1], out[2], out[3], Bd)
    bytes_it = M * N * 8
    gbs = bytes_it * k / (ms * 1e-3) / 1e9
    return {
        "metric": "omp solves/sec at 1024x8192,k=32 FP64, single signal (dictionary L2-resident)", "value": 1e3 / ms,
        "unit": UNIT, "ms_per_step": ms, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "single-signal omp on the config-2 dictionary (64 MiB FP64 / 32 MiB FP32 < 126 MB L2)",
                   "M": M, "N": N, "k": k, "timed_solves": 100,
                   "l2": "dictionary deliberately L2-resident: this is the L2-regime single-signal path"},
        "per_dtype": res, "multi_signal_f64": multi, "gpu_launches": 100 * 2 * k,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None,
                     "kernel": "whole solve (corr_gemv_kernel + update per update!)", "bytes_per_launch": bytes_it,
                     "peak_source": hbm_src + "; the dictionary is served from L2, so this fraction of the HBM peak is "
                                              "the north-star's 'HBM/L2 roofline' figure for the single-signal path"},
        "cpu_baseline": cpu, "check": {"oracle_parity": parity},
    }


# ------------------------------------------------------------------------------------ config 3 (gomp)
def run_c3(ctx, args):
    cs = ctx.cs
    Mr, Nc, k, l = 2048, 32768, 64, 4
    B = max(256, int(8192 * args.scale))
    A, Bm, idx = make_problem(ctx, Mr, Nc, k, B, np.float64, seed=1234 + 3)
    peak, peak_src = fp64_peak_committed()
    steps = max(1, min(args.steps, 3))
    with cs.Dictionary(A, device=ctx.local) as D:
        with cs.Batch(D, B, k) as b:
            b.upload(Bm)
            b.gomp(l, k, EPS64)                                  # warm-up (builds the cached Gram matrix)
            b.profile(True)
            dev_ms = 0.0
            for _ in range(steps):
                b.gomp(l, k, EPS64)
                dev_ms += b.last_solve_ms()
            corr_ms, n, other = b.corr_time()
            b.profile(False)
            sel, coef, nnz, res, its = b.download(k)
        # e2e through the one-shot C call (pageable -> the library's own staging)
        t0 = time.perf_counter()
        xs = cs.gomp(D, Bm, l, k, result="csc")
        e2e_t = time.perf_counter() - t0
    tf = 2.0 * Mr * Nc * B * n / corr_ms / 1e9
    cpu, parity = None, None
    if ctx.rank == 0 and ctx.world == 1 and args.cpu_signals > 0:
        ns = 16
        r = run_c_oracle("gomp", A, Bm[:, :ns], k, l=l)
        if r is not None:
            cpu = {"value": ns / r[0], "unit": UNIT, "cores": r[3], "kind": "port",
                   "sample": f"first {ns} of the {B} signals, {r[2]} ({r[0]:.1f} s)"}
            parity = parity_vs_c_oracle(r[1], sel[:ns], coef[:ns], nnz[:ns], res[:ns], Bm)
    return {
        "metric": "gomp solves/sec at 2048x32768,k=64,l=4 FP64", "value": B * steps / (dev_ms * 1e-3), "unit": UNIT,
        "ms_per_step": dev_ms / steps, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "gomp l=4 2048x32768 f64 k=64, 8192 signals (BASELINE config 3)", "M": Mr, "N": Nc, "k": k,
                   "l": l, "signals": B, "l2": "512 MiB dictionary + 2 x 128 MiB signals/residuals exceed the 126 MB L2"},
        "e2e": {"value": B / e2e_t, "unit": UNIT, "h2d_bytes_per_step": Mr * B * 8, "d2h_bytes_per_step": B * (16 + 12 * k),
                "steps": 1, "api": "csb200_gomp + csb200_assemble_csc (host buffers)", "nnz_total": int(xs.nnz)},
        "gpu_launches": int(n + other + steps),
        "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "traffic": None,
                     "kernel": "corr_gemm_f64_kernel (dense |A'r| store epilogue)", "launches": int(n),
                     "mean_launch_ms": corr_ms / max(1, n), "share_of_step": corr_ms / dev_ms,
                     "flop_per_launch": 2.0 * Mr * Nc * B, "peak_source": peak_src},
        "cpu_baseline": cpu,
        "check": {"support_recovered_frac": supports_recovered(sel[::16], nnz[::16], idx[::16]), "max_resnorm": float(res.max()),
                  "oracle_parity": parity},
    }


# ------------------------------------------------------------------------------------ config 5 (mp)
def run_c5(ctx, args):
    cs = ctx.cs
    Mr, Nc, iters, B = 4096, 65536, max(4, int(200 * args.scale)), 4096
    A, Bm, idx = make_problem(ctx, Mr, Nc, 32, B, np.float64, seed=1234 + 5, noise=5e-3)
    peak, peak_src = fp64_peak_committed()
    chk_it, chk_ns = 8, 8
    with cs.Dictionary(A, device=ctx.local) as D, cs.Batch(D, B, iters) as b:
        b.upload(Bm)
        b.mp(chk_it)                                             # warm-up; its history doubles as the parity sample
        hsel, hcoef, _, hres, _ = b.download(chk_it)
        b.profile(True)
        b.mp(iters)
        dev_ms = b.last_solve_ms()
        corr_ms, n, other = b.corr_time()
        b.profile(False)
        scr = b.screen_stats(reset=True)
        sel, coef, nnz, res, its = b.download(iters)
        dev_ms64, other_leg = None, None
        # the same workload through the other correlation pass (a few iterations, scaled): FP64 DMMA when the default
        # screened, forced TF32 screening when the default is the DMMA pass (signals longer than 2048 rows)
        os.environ["CSB200_SCREEN"] = "0" if scr["path_id"] in (3, 4) else "1"
        try:
            it2 = max(2, min(iters, 16))
            b.mp(it2)
            dev_ms64 = b.last_solve_ms() * iters / it2
            scr2 = b.screen_stats(reset=True)
            sel2, coef2, *_ = b.download(it2)
            same2 = bool(np.array_equal(sel2[:, :it2], sel[:, :it2]) and np.array_equal(coef2[:, :it2], coef[:, :it2]))
            other_leg = {"value": B / (dev_ms64 * 1e-3), "unit": UNIT, "path": scr2["path"],
                         "how": f"CSB200_SCREEN={os.environ['CSB200_SCREEN']}, the first {it2} iterations scaled to {iters} "
                                "(later iterations run deeper into the noise, where the screening window holds more atoms)",
                         "bit_identical_to_default_path": same2}
        finally:
            os.environ.pop("CSB200_SCREEN", None)
    screened = scr["path_id"] in (3, 4)
    if screened:
        peak, peak_src = tf32_peak()                                 # mp is screened with TF32 operands (atoms need not be normalised)
    tf = 2.0 * Mr * Nc * B * n / corr_ms / 1e9
    cpu, parity = None, None
    if ctx.rank == 0 and ctx.world == 1 and args.cpu_signals > 0:
        r = run_c_oracle("mp", A, Bm[:, :chk_ns], chk_it)
        if r is not None:
            t, got = r[0], r[1]
            cpu = {"value": chk_ns / (t * iters / chk_it), "unit": UNIT, "cores": r[3], "kind": "port",
                   "sample": f"{chk_ns} signals x the first {chk_it} of {iters} iterations, {r[2]} ({t:.1f} s), "
                             f"extrapolated linearly to {iters} iterations"}
            order_ok = bool(np.array_equal(got["order"][:chk_ns, :chk_it], hsel[:chk_ns, :chk_it]))
            err = 0.0
            for s in range(chk_ns):                              # x[i] += <a_i, r> accumulated per atom (:29)
                acc = {}
                for i, c in zip(hsel[s, :chk_it].tolist(), hcoef[s, :chk_it].tolist()):
                    acc[i] = acc.get(i, 0.0) + c
                gi = np.array(sorted(acc)); gv = np.array([acc[i] for i in gi.tolist()])
                nn = int(got["nnz"][s])
                if nn != gi.size or not np.array_equal(gi, got["nzind"][s, :nn]):
                    err = float("inf")
                    continue
                err = max(err, float(np.max(np.abs(gv - got["nzval"][s, :nn])) / np.max(np.abs(gv))))
            rerr = float(np.max(np.abs(hres[:chk_ns] - got["resnorm"][:chk_ns]) / np.linalg.norm(Bm[:, :chk_ns], axis=0)))
            parity = {"signals": chk_ns, "iterations": chk_it, "selection_order_exact": order_ok, "coef_max_rel_err": err,
                      "resnorm_max_abs_err_over_norm_b": rerr, "tolerance": 1e-10,
                      "pass": order_ok and err <= 1e-10 and rerr <= 1e-10}
    return {
        "metric": f"mp solves/sec at 4096x65536, {iters} iterations FP64", "value": B / (dev_ms * 1e-3), "unit": UNIT,
        "ms_per_step": dev_ms, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "plain mp 4096x65536 f64, 200 iterations, 4096 signals (BASELINE config 5)", "M": Mr, "N": Nc,
                   "iterations": iters, "signals": B, "l2": "2 GiB dictionary exceeds the 126 MB L2"},
        "gpu_launches": int(n + other + 1),
        "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "traffic": None,
                     "kernel": "corr_screen_tf32_kernel" if screened else "corr_gemm_f64_kernel", "launches": int(n),
                     "mean_launch_ms": corr_ms / max(1, n),
                     "share_of_step": corr_ms / dev_ms, "flop_per_launch": 2.0 * Mr * Nc * B, "peak_source": peak_src},
        "other_path": other_leg,
        "cpu_baseline": cpu, "check": {"median_resnorm": float(np.median(res)), "oracle_parity": parity, "path": scr["path"],
                                       "screening": {key: scr[key] for key in ("signal_updates", "candidates_reevaluated", "exact_scans")} if screened else None},
    }


# ------------------------------------------------------------------------------------ config 4 (column-sharded omp)
def run_c4(ctx, args):
    """Single-signal omp, 8192 x (131072 * N) FP32, k = 128, one 4 GiB column shard per GPU (8 GPUs = BASELINE config 4;
    1 GPU = the scaled-down twin of SURVEY 8d).  Per iteration: local GEMV + arg-max, one exchange of
    {|c|, index, atom column} between all ranks (peer-memory mailboxes over NVLink, NCCL all-gather as fallback), then
    the replicated update."""
    cs, torch, dist = ctx.cs, ctx.torch, ctx.dist
    rank, world = ctx.rank, ctx.world
    Mr, k = 8192, max(8, int(128 * args.scale))
    N_loc = 131072
    Nt = N_loc * world
    comm = ctx.shard_comm()
    g = torch.Generator(device=ctx.dev).manual_seed(100 + rank)
    A_t = torch.empty(N_loc, Mr, dtype=torch.float32, device=ctx.dev)
    for n0 in range(0, N_loc, 16384):
        blk = torch.randn(16384, Mr, dtype=torch.float32, device=ctx.dev, generator=g)
        blk /= blk.norm(dim=1, keepdim=True)
        A_t[n0:n0 + 16384] = blk
    per = max(1, k // world)                              # planted atoms: k / world of every rank's own atoms
    gi = torch.Generator(device=ctx.dev).manual_seed(7)
    loc_idx = torch.randperm(N_loc, device=ctx.dev, generator=gi)[:per]
    part = A_t[loc_idx].to(torch.float64).sum(dim=0)
    if world > 1:
        dist.all_reduce(part)
    b = part.to(torch.float32).cpu().numpy()
    A_np = A_t.cpu().numpy().T
    del A_t
    torch.cuda.empty_cache()
    shard = cs.Dictionary(A_np, device=ctx.local, n_offset=rank * N_loc, n_total=Nt)
    os.environ["CSB200_SHARD_TIMING"] = "2"               # record the per-phase device times (silent)
    hbm, hbm_src = hbm_peak()
    steps = max(2, min(args.steps, 5))
    cs.omp_sharded(shard, comm, b, k)                     # warm-up: sizes the scratch, sets up the peer mailboxes
    walls, tms = [], []
    x = info = None
    for _ in range(steps):
        ctx.barrier()
        t0 = time.perf_counter()
        x, info = cs.omp_sharded(shard, comm, b, k)
        torch.cuda.synchronize()
        walls.append(time.perf_counter() - t0)
        tms.append(comm.last_timing())
    wall = ctx.maxr(float(np.mean(walls)))
    tm = {key: float(np.mean([t[key] for t in tms])) for key in ("solve_ms", "corr_ms", "exchange_ms", "update_ms", "gap_ms")}
    solve_ms, corr_ms, exch_ms, upd_ms, gap_ms = ctx.maxr(tm["solve_ms"], tm["corr_ms"], tm["exchange_ms"], tm["update_ms"],
                                                          tm["gap_ms"])
    planted = set((loc_idx.cpu().numpy() + rank * N_loc).tolist())
    found = ctx.sumr(float(len(planted & set(x.nzind.tolist()))))
    gbs = N_loc * Mr * 4 * k / (corr_ms * 1e-3) / 1e9     # one shard's bytes per correlation pass / its device time
    cpu, parity = None, None
    if rank == 0 and world == 1 and args.cpu_signals > 0:
        # the twin IS the whole dictionary at N = 1: run the FP32 NumPy oracle on the same bytes for a few update!s
        from oracle import pursuit_oracle as po
        kc = 8
        xg, ig = cs.omp_sharded(shard, comm, b, kc)
        t0 = time.perf_counter()
        tr = po.Trace()
        ref = po.omp(A_np, b, kc, trace=tr)
        t = time.perf_counter() - t0
        cpu = {"value": 1.0 / (t * k / kc), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"first {kc} of the {k} update!s on the same 4 GiB FP32 dictionary, NumPy/OpenBLAS oracle "
                         f"({t:.1f} s), extrapolated linearly to k = {k}"}
        order_ok = ig["order"].tolist() == tr.order()
        cerr = float(np.max(np.abs(xg.nzval - np.asarray(ref.nzval)) / np.max(np.abs(ref.nzval)))) if order_ok else float("inf")
        parity = {"update_steps": kc, "selection_order_exact": bool(order_ok), "coef_max_rel_err": cerr, "tolerance": 2e-5,
                  "pass": bool(order_ok and cerr <= 2e-5)}
    shard.close()
    del A_np
    per_it = {"gemv_us": 1e3 * corr_ms / k, "exchange_us": 1e3 * exch_ms / k, "update_us": 1e3 * upd_ms / k,
              "gap_us": 1e3 * gap_ms / k}
    limiter = max(per_it, key=per_it.get)
    return {
        "metric": f"single-signal omp solves/sec at 8192x{Nt} FP32, k={k}, column-sharded over {world} GPU(s)",
        "value": 1e3 / solve_ms, "unit": UNIT, "ms_per_step": solve_ms, "higher_is_better": True, "dtype": "f32",
        "data": "synthetic", "scaling": "weak",
        "config": {"workload": "single-signal omp on a column-sharded FP32 dictionary, 131072 atoms (4 GiB) per GPU "
                               "(BASELINE config 4 at 8 GPUs; its scaled-down twin at 1 GPU)", "M": Mr, "N": Nt, "k": k,
                   "atoms_per_gpu": N_loc, "l2": "4 GiB shard per GPU exceeds the 126 MB L2",
                   "exchange": info["exchange"], "timed_solves": steps},
        "e2e": {"value": 1.0 / wall, "unit": UNIT, "h2d_bytes_per_step": Mr * 4, "d2h_bytes_per_step": 12 * k + 24,
                "steps": steps, "api": "csb200_omp_sharded (host buffers; upload, input check, solve, download)"},
        "per_iteration_us": per_it, "limiter": limiter,
        "gpu_launches": int(steps * (3 * k + 2)),
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                     "traffic": _traffic("corr_gemv_traffic.json"), "kernel": "corr_gemv_kernel<float>",
                     "launches": int(steps * k), "mean_launch_ms": corr_ms / k, "share_of_step": corr_ms / solve_ms,
                     "bytes_per_launch": N_loc * Mr * 4, "peak_source": hbm_src,
                     "whole_solve_frac_of_hbm_peak": N_loc * Mr * 4 * k / (solve_ms * 1e-3) / 1e9 / hbm},
        "cpu_baseline": cpu,
        "check": {"planted_atoms_found": int(found), "planted_atoms": per * world, "resnorm": info["resnorm"],
                  "oracle_parity": parity},
    }


def _traffic(name):
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
        return json.load(open(p)).get("dram_bytes_per_launch")
    return None


RUNNERS = {"c1": run_c1, "c2s": run_c2s, "c3": run_c3, "c4": run_c4, "c5": run_c5}


def ours_arm(args):
    ctx = Ctx(args)
    if args.config == "c2":
        line = run_c2(ctx, args)
    else:
        sampler = ClockSampler(ctx.local)
        sampler.start()
        line = RUNNERS[args.config](ctx, args)
        line["clocks"] = sampler.stop()
        line.setdefault("scaling", "weak")
        line["vs_baseline"] = None
    if args.secondary == "auto":
        # single-GPU configs are measured at N = 1; the column-sharded config at every N
        names = ["c1", "c2s", "c3", "c4", "c5"] if ctx.world == 1 else ["c4"]
    elif args.secondary in ("none", ""):
        names = []
    else:
        names = [n for n in args.secondary.split(",") if n in RUNNERS]
    secondary = {}
    for name in names:
        if name == args.config:
            continue
        try:
            t0 = time.perf_counter()
            r = RUNNERS[name](ctx, args)
            r["bench_wall_s"] = time.perf_counter() - t0
            secondary[name] = r
        except Exception as exc:                              # a secondary config must not take the headline down
            secondary[name] = {"error": f"{type(exc).__name__}: {exc}"}
            if ctx.world > 1:
                raise
    if ctx.rank == 0:
        out = {"n_gpus": ctx.world, "steps": args.steps, "warmup": args.warmup}
        out.update(line)
        if secondary:
            out["secondary"] = secondary
        print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours_arm(a)
