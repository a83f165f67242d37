"""CPU tests of the multi-process host logic (world_size 2, gloo): shard arithmetic, unique-id exchange, and the
per-iteration exchange protocol of the column-sharded solve emulated with numpy + torch.distributed --
it must reproduce the unsharded oracle's support sequence, ties included."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as ge
    from oracle import pursuit_oracle as po
    cs = ge.load_package()
    try:
        # 1. unique id travels from rank 0 to everyone
        uid = cs.exchange_unique_id(dist, rank, make_id=lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        # 2. column-sharded OMP protocol: local argmax -> all-gather of (|c|, idx, column) -> same pick everywhere
        rng = np.random.default_rng(42)
        M, N, k = 48, 101, 6                         # N not divisible by the number of ranks
        A = po.gaussian_dictionary(rng, M, N)
        A[:, 77] = A[:, 13]                          # a tie across the shard boundary: atom 13 must win
        x0 = po.SparseVec(N, [13, 40, 90], [1.0, -1.0, 1.0])
        b = A[:, x0.nzind] @ np.array(x0.nzval)
        lo, hi = cs.shard_range(N, world, rank)
        A_loc = A[:, lo:hi]
        support, cols = [], []
        r = b.copy()
        for it in range(k):
            c = np.abs(A_loc.T @ r)
            j = int(np.argmax(c))
            rec = torch.zeros(2 + M, dtype=torch.float64)
            rec[0], rec[1] = float(c[j]), float(lo + j)
            rec[2:] = torch.from_numpy(A_loc[:, j].copy())
            gathered = [torch.zeros_like(rec) for _ in range(world)]
            dist.all_gather(gathered, rec)
            g, v, idx = cs.pick_global([(float(t[0]), int(t[1])) for t in gathered])
            assert cs.owner_of(idx, N, world) == g
            if idx in support:
                continue
            support.append(idx); cols.append(gathered[g][2:].numpy())
            AS = np.stack(cols, axis=1)
            xs, *_ = np.linalg.lstsq(AS, b, rcond=None)
            r = b - AS @ xs
        t = po.Trace()
        po.omp(A, b, k, eps=0.0, trace=t)
        assert support == t.order(), (support, t.order())
        assert support[0] in (13, 40, 90) and 77 not in support[:3]
        # 3. batched mode: signals split without communication, results concatenate in rank order
        B = 10
        slo, shi = cs.shard_range(B, world, rank)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([shi - slo]))
        assert sum(int(s) for s in sizes) == B
        out_q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out_q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_shard_arithmetic(cs):
    for n, w in [(10, 3), (1048576, 8), (7, 8), (101, 2), (0, 4)]:
        covered = []
        for r in range(w):
            lo, hi = cs.shard_range(n, w, r)
            covered += list(range(lo, hi))
            for i in (lo, hi - 1):
                if lo < hi:
                    assert cs.owner_of(i, n, w) == r
        assert covered == list(range(n))
    assert cs.pick_global([(1.0, 5), (1.0, 3), (0.5, 0)]) == (1, 1.0, 3)
    assert cs.pick_global([(0.0, -1), (0.0, 7)]) == (1, 0.0, 7)
    assert cs.pick_global([(0.0, -1)])[2] == -1
