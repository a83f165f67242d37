"""torchrun entry: column-sharded OMP over N GPUs against the oracle and the single-GPU path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/run_sharded_gpu.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import pursuit_oracle as po  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cs = ge.load_package()
    uid = cs.exchange_unique_id(dist, rank)
    comm = cs.ShardComm(uid, rank, world, local)
    ok = True
    for dtype, M, N, k in [(np.float32, 512, 40000, 24), (np.float64, 256, 9001, 10)]:
        rng = np.random.default_rng(5)                       # same bytes on every rank
        A = po.gaussian_dictionary(rng, M, N, dtype)
        A[:, N - 3] = A[:, 7]                                # a tie between the first and the last shard
        x0 = po.sparse_vector(rng, N, k - 1)
        x0.setindex(7, 1.5)
        b = (A[:, x0.nzind].astype(np.float64) @ np.asarray(x0.nzval)).astype(dtype)
        b = po.perturb(rng, b, 5e-3)
        lo, hi = cs.shard_range(N, world, rank)
        with cs.Dictionary(np.asfortranarray(A[:, lo:hi]), device=local, n_offset=lo, n_total=N) as shard:
            x, info = cs.omp_sharded(shard, comm, b, k)                 # default transport: peer-memory mailboxes
            os.environ["CSB200_SHARD_EXCHANGE"] = "nccl"
            xn, infon = cs.omp_sharded(shard, comm, b, k)               # same solve through ncclAllGather
            del os.environ["CSB200_SHARD_EXCHANGE"]
            x2, info2 = cs.omp_sharded(shard, comm, b, k)               # and back (sequence numbers keep counting)
        want = os.environ.get("CSB200_EXPECT_EXCHANGE", "peer-memory")
        ok &= info["exchange"] == want and infon["exchange"] == "nccl" and info2["exchange"] == want
        for y, iy in ((xn, infon), (x2, info2)):                       # transports are bit-identical
            ok &= y.nzind.tolist() == x.nzind.tolist() and bool(np.array_equal(y.nzval, x.nzval))
            ok &= iy["order"].tolist() == info["order"].tolist() and iy["resnorm"] == info["resnorm"]
        # every rank must hold the same answer
        mine = torch.tensor(np.concatenate([x.nzind.astype(np.float64), x.nzval, [info["resnorm"]]]), device="cuda")
        ref0 = mine.clone()
        dist.broadcast(ref0, src=0)
        ok &= bool(torch.equal(mine, ref0))
        if rank == 0:
            t = po.Trace()
            ref = po.omp(A, b, k, trace=t)
            with cs.Dictionary(A, device=local) as D:
                x1 = cs.omp(D, b, k)
            ok &= info["order"].tolist() == t.order()
            ok &= x.nzind.tolist() == ref.nzind == x1.nzind.tolist()
            ok &= bool(np.allclose(x.nzval, ref.nzval, rtol=2e-5 if dtype == np.float32 else 1e-10, atol=1e-6 if dtype == np.float32 else 1e-11))
            ok &= bool(np.allclose(x.nzval, x1.nzval, rtol=1e-12, atol=1e-13))     # sharded == unsharded GPU
            ok &= 7 in x.nzind.tolist() and (N - 3) not in x.nzind.tolist()
            print(f"dtype={np.dtype(dtype).name} world={world} ok={ok} exchange={info['exchange']} corr_ms={info['corr_ms']:.3f}", flush=True)
    # the scaled-down twin of config 4 (8192 x 131072 FP32, k = 128) split over all ranks must give the
    # single-rank result bit for bit (every c_j keeps its summation order under any column partition)
    if os.environ.get("CSB200_SKIP_TWIN", "0") != "1":
        M, N, k = 8192, 131072, 128
        g = torch.Generator(device="cuda").manual_seed(100)          # same bytes on every rank
        A_t = torch.empty(N, M, dtype=torch.float32, device="cuda")
        for n0 in range(0, N, 16384):
            blk = torch.randn(16384, M, dtype=torch.float32, device="cuda", generator=g)
            blk /= blk.norm(dim=1, keepdim=True)
            A_t[n0:n0 + 16384] = blk
        idx = torch.randperm(N, device="cuda", generator=g)[:k]
        b = A_t[idx].to(torch.float64).sum(dim=0).to(torch.float32).cpu().numpy()
        lo, hi = cs.shard_range(N, world, rank)
        A_loc = A_t[lo:hi].cpu().numpy().T
        with cs.Dictionary(A_loc, device=local, n_offset=lo, n_total=N) as shard:
            x, info = cs.omp_sharded(shard, comm, b, k)
        del A_loc
        if rank == 0:
            A_full = A_t.cpu().numpy().T
            solo = cs.ShardComm(cs.ShardComm.unique_id(), 0, 1, local)
            with cs.Dictionary(A_full, device=local) as D:
                x1, info1 = cs.omp_sharded(D, solo, b, k)
            solo.close()
            same = (info["order"].tolist() == info1["order"].tolist() and bool(np.array_equal(x.nzval, x1.nzval))
                    and info["resnorm"] == info1["resnorm"])
            ok &= same and sorted(idx.cpu().tolist()) == x.nzind.tolist()
            print(f"twin 8192x131072 f32 k=128: {world} shards vs 1 shard bit-identical={same} exchange={info['exchange']} "
                  f"corr_ms={info['corr_ms']:.2f} resnorm={info['resnorm']:.3e}", flush=True)
        del A_t
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    comm.close()
    dist.destroy_process_group()
    if rank == 0:
        print("SHARDED_OK" if int(flag.item()) == 1 else "SHARDED_FAIL", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
