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


def gpu_solver_device(settings: Dict[str, object] | None = None, box=False, hessian=None, device=-1) -> Callable:
    """Like gpu_solver, but the returned solve(data) leaves the solutions ON THE DEVICE: it returns ONE packed torch
    CUDA tensor [B_local, n + n_eq + n_cons + 20] = (x | y | z | info20) built from zero-copy views of the batch's
    result buffers (DenseBatch.results_device), ready for an NCCL all_gather without a host round trip. The batch
    object is kept alive on the function (`solve.batch`) so that the views stay valid."""
    from . import proxqp

    def solve(data: Dict[str, np.ndarray]):
        import torch

        B, n = data["g"].shape
        ne = data["b"].shape[1] if data.get("b") is not None else 0
        ni = data["u"].shape[1] if data.get("u") is not None else 0
        db = getattr(solve, "batch", None)
        if db is None or (db.batch, db._g.n, db._g.n_eq, db._g.n_in) != (B, n, ne, ni):
            db = proxqp.dense.DenseBatch(B, n, ne, ni, box, proxqp.HessianType.Dense if hessian is None else hessian, device=device)
            solve.batch = db
        for k, v in (settings or {}).items():
            setattr(db.settings, k, v)
        db.init(**data)
        db.solve()
        r = db.results_device()
        return torch.cat([r["x"], r["y"], r["z"], r["info"]], dim=1)
    return solve


def unpack_results(packed: np.ndarray, n: int, ne: int, nc: int):
    """(x, y, z, info7) from the packed layout of gpu_solver_device; info7 = status, iter, iter_ext, mu_updates,
    pri_res, dua_res, objValue (the columns gpu_solver returns)."""
    x, y, z, inf = packed[:, :n], packed[:, n:n + ne], packed[:, n + ne:n + ne + nc], packed[:, n + ne + nc:]
    return x, y, z, inf[:, [10, 6, 7, 8, 15, 16, 14]]


def solve_sharded_device(data: Dict[str, np.ndarray], solver: Callable, group=None):
    """solve_sharded with the solutions gathered FROM DEVICE BUFFERS: rank r solves its contiguous slice with `solver`
    (gpu_solver_device: returns the packed device tensor), one NCCL all_gather_into_tensor moves (x, y, z, info) of every
    rank over NVLink, and each rank gets the packed [B, n + n_eq + n_cons + 20] CUDA tensor of the whole batch
    (unpack_results(t.cpu().numpy(), ...) gives the host arrays). Slices are padded to the largest one for the collective."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = data["g"].shape[0]
    lo, hi = shard_bounds(B, world, rank)
    local = {k: np.ascontiguousarray(v[lo:hi]) for k, v in data.items() if v is not None}
    packed = solver(local)
    if world == 1:
        return packed
    maxn = max(shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0] for r in range(world))
    if packed.shape[0] != maxn:
        pad = torch.zeros((maxn, packed.shape[1]), dtype=packed.dtype, device=packed.device)
        pad[: packed.shape[0]] = packed
        packed = pad
    out = torch.empty((world * maxn, packed.shape[1]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed.contiguous(), group=group)
    if B == world * maxn:
        return out
    return torch.cat([out[r * maxn: r * maxn + (shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0])] for r in range(world)], dim=0)


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
