"""N > 1 host logic on CPU: world_size-2 gloo run of the batch sharding +
solution gather (proxsuite_b200/sharding.py). The per-rank solver is the
oracle here (no GPU in this container); on the GPU box the same code path runs
with the CUDA solver and NCCL (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    from proxsuite_b200.sharding import shard_bounds

    for B in (0, 1, 7, 1024, 4097):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _oracle_solver(local):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O

    B, n = local["g"].shape
    ne, ni = local["b"].shape[1], local["u"].shape[1]
    X, Y, Z, I = np.zeros((B, n)), np.zeros((B, ne)), np.zeros((B, ni)), np.zeros((B, 7))
    for i in range(B):
        q = O.OracleQP(n, ne, ni)
        q.set(eps_abs=1e-9, eps_rel=0)
        q.init(**{k: local[k][i] for k in "HgAbClu"})
        r = q.solve()
        X[i], Y[i], Z[i] = r.x, r.y, r.z
        I[i] = [r.info.status, r.info.iter, r.info.iter_ext, r.info.mu_updates, r.info.pri_res, r.info.dua_res, r.info.objValue]
    return X, Y, Z, I


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from proxsuite_b200.sharding import solve_sharded

    B, n, ne, ni = 7, 12, 4, 6
    data = [O.generate_qp("strongly_convex", i, n, ne, ni) for i in range(B)]
    st = {k: np.stack([d[k] for d in data]) for k in "HgAbClu"}
    x, y, z, info = solve_sharded(st, _oracle_solver)
    if rank == 0:
        q.put((x, info))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_solve_matches_serial():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    x, info = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from oracle import oracle as O

    B, n, ne, ni = 7, 12, 4, 6
    st = {k: np.stack([O.generate_qp("strongly_convex", i, n, ne, ni)[k] for i in range(B)]) for k in "HgAbClu"}
    xs, _, _, infos = _oracle_solver(st)
    assert np.array_equal(x, xs)          # serial == sharded, bitwise
    assert (info[:, 0] == 0).all() and np.array_equal(info[:, 1], infos[:, 1])
