"""Multi-GPU sharding of a batch of independent QPs (one process per GPU).

QPs are independent (parallel/qp_solve.hpp:55-59 has no cross-QP state), so a
batch shards into contiguous slices [rank*B/W, (rank+1)*B/W) with no
collective inside the iteration. The only exchange is the final gather of the
solutions (x, y, z, info), a torch.distributed all_gather (NCCL over NVLink on
GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import numpy as np


def shard_bounds(batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition: the first (batch % world) ranks own one extra QP."""
    base, rem = divmod(int(batch), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def gpu_solver(settings: Dict[str, object] | None = None, box=False, hessian=None, device=-1) -> Callable:
    """Returns solve(data) -> (x, y, z, info20) running the CUDA path on this rank's GPU."""
    from . import proxqp

    def solve(data: Dict[str, np.ndarray]):
        B, n = data["g"].shape
        ne = data["b"].shape[1]
        ni = data["u"].shape[1]
        db = proxqp.dense.DenseBatch(B, n, ne, ni, box, proxqp.HessianType.Dense if hessian is None else hessian, device=device)
        for k, v in (settings or {}).items():
            setattr(db.settings, k, v)
        db.init(**data)
        db.solve()
        r = db.results()
        info = np.stack([r["info"][k].astype(np.float64) for k in ("status", "iter", "iter_ext", "mu_updates", "pri_res", "dua_res", "objValue")], axis=1)
        return r["x"], r["y"], r["z"], info
    return solve


def solve_sharded(data: Dict[str, np.ndarray], solver: Callable, group=None, device=None):
    """Every rank passes the FULL stacked batch (or at least its own slice filled
    in); rank r solves slice shard_bounds(B, W, r) with `solver` and all ranks
    receive the gathered (x, y, z, info) of the whole batch."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = data["g"].shape[0]
    lo, hi = shard_bounds(B, world, rank)
    local = {k: np.ascontiguousarray(v[lo:hi]) for k, v in data.items() if v is not None}
    x, y, z, info = solver(local)
    if world == 1:
        return x, y, z, info
    outs = []
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    maxn = max(shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0] for r in range(world))
    for arr in (x, y, z, info):
        w = arr.shape[1]
        pad = np.zeros((maxn, w))
        pad[: arr.shape[0]] = arr
        t = torch.from_numpy(pad).to(dev)
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t, group=group)
        parts = []
        for r in range(world):
            rlo, rhi = shard_bounds(B, world, r)
            parts.append(gathered[r][: rhi - rlo].cpu().numpy())
        outs.append(np.concatenate(parts, axis=0))
    return tuple(outs)
