"""proxsuite.torch.qplayer.QPFunction on the B200 batch path.

Mirror of bindings/python/proxsuite/torch/qplayer.py:12-253 (the feasible variant `QPFunctionFn`): the forward pass
stacks the batch into ONE device-resident DenseBatch (the reference builds a BatchQP object by object and calls
solve_in_parallel, qplayer.py:105-170), the backward pass is ONE pqp_batch_backward launch
(solve_backward_in_parallel, qplayer.py:172-253). Same solver settings as the reference's layer
(qplayer.py:120-127: max_iter_in = 100, default_rho = refactor_rho_threshold = 5e-5, eps_abs = eps,
primal_infeasibility_solving = False) and the same argument / return order.

Differences: gradients come back in the dtype of the inputs (the reference allocates float32); the closest-feasible
variant (`structural_feasibility=False`, QPFunctionFn_infeas) is not implemented. CUDA float64 tensors never leave
the device: the layer hands their data pointers to pqp_batch_init_device / pqp_batch_results_copy_device /
pqp_batch_backward_device; CPU tensors (and other dtypes) go through the host entry points."""
from __future__ import annotations

import numpy as np
import torch
from torch.autograd import Function

import ctypes as _ct
import os as _os

from .. import _capi
from ..proxqp import dense as _dense


def _expand(t, n_batch, dims):
    """utils.expandParam (torch/utils.py): a parameter without batch dimension is shared by the whole batch."""
    if t is None or t.nelement() == 0:
        return None
    if t.dim() == dims:
        return t
    if t.dim() == dims - 1:
        return t.unsqueeze(0).expand(*([n_batch] + list(t.size())))
    raise RuntimeError("Unexpected number of dimensions.")


def _n_batch(*ts):
    dims = [3, 2, 3, 2, 3, 2, 2]
    for t, d in zip(ts, dims):
        if t is not None and t.nelement() > 0 and t.dim() == d:
            return t.size(0)
    return 1


def _np(t):
    return None if t is None else np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float64)


def _on_device(*ts):
    """Device entry points: every tensor is CUDA (PQP_QPLAYER_DEVICE_API=1 forces them for CPU tensors — only
    meaningful on the CPU emulator of tests/emu, where "device" memory is host memory)."""
    ts = [t for t in ts if t is not None]
    return all(t.is_cuda for t in ts) or _os.environ.get("PQP_QPLAYER_DEVICE_API") == "1"


def _dev(t):
    """contiguous float64 copy (kept alive by the caller) and its data pointer"""
    if t is None:
        return None, None
    c = t.detach().to(torch.float64).contiguous()
    return c, _ct.c_void_p(c.data_ptr())


def QPFunction(eps=1e-9, maxIter=1000, eps_backward=1.0e-4, rho_backward=1.0e-6, mu_backward=1.0e-6, omp_parallel=False,
               structural_feasibility=True):
    """Returns the autograd function `(Q, p, A, b, G, l, u) -> (zhats, lams, nus)` (qplayer.py:12-90).
    `omp_parallel` is accepted for signature parity: the batch always runs as one persistent kernel."""
    if not structural_feasibility:
        raise NotImplementedError("QPFunctionFn_infeas (closest feasible QP) is not available on the B200 path yet")

    class QPFunctionFn(Function):
        @staticmethod
        def forward(ctx, Q_, p_, A_, b_, G_, l_, u_):
            n_batch = _n_batch(Q_, p_, A_, b_, G_, l_, u_)
            Q, p = _expand(Q_, n_batch, 3), _expand(p_, n_batch, 2)
            G, u, l = _expand(G_, n_batch, 3), _expand(u_, n_batch, 2), _expand(l_, n_batch, 2)
            A, b = _expand(A_, n_batch, 3), _expand(b_, n_batch, 2)
            nz = Q.size(1)
            nineq = G.size(1) if G is not None else 0
            neq = A.size(1) if A is not None else 0
            assert neq > 0 or nineq > 0
            ctx.n_batch, ctx.nz, ctx.neq, ctx.nineq = n_batch, nz, neq, nineq
            db = _dense.DenseBatch(n_batch, nz, neq, nineq)
            s = db.settings  # qplayer.py:120-127
            s.primal_infeasibility_solving = False
            s.max_iter = maxIter
            s.max_iter_in = 100
            default_rho = 5.0e-5
            s.default_rho = default_rho
            s.refactor_rho_threshold = default_rho  # no refactorization
            s.eps_abs = eps
            ctx.batch = db
            ctx.device_api = _on_device(Q, p, A, b, G, l, u)
            if ctx.device_api:
                grp = db._g
                db._push()
                keep = [_dev(t) for t in (Q, p, A, b, G, l, u)]
                if Q.is_cuda:  # the library copies on its own stream: the producers of the inputs must be done
                    torch.cuda.current_stream(Q.device).synchronize()
                rho = _ct.c_double(default_rho)
                _capi.check(grp.lib.pqp_batch_init_device(grp.handle, 0, n_batch, *[k[1] for k in keep], None, None, 1,
                                                         _ct.cast(_ct.pointer(rho), _ct.c_void_p), None, None, None))
                db.solve()
                zh = torch.empty((n_batch, nz), dtype=torch.float64, device=Q.device)
                lam = torch.empty((n_batch, neq), dtype=torch.float64, device=Q.device)
                nu = torch.empty((n_batch, nineq), dtype=torch.float64, device=Q.device)
                _capi.check(grp.lib.pqp_batch_results_copy_device(grp.handle, 0, n_batch, _ct.c_void_p(zh.data_ptr()),
                                                                 _ct.c_void_p(lam.data_ptr()) if neq else None,
                                                                 _ct.c_void_p(nu.data_ptr()) if nineq else None))
                return zh.to(Q.dtype), lam.to(Q.dtype), nu.to(Q.dtype)
            db.init(H=_np(Q), g=_np(p), A=_np(A), b=_np(b), C=_np(G), l=_np(l), u=_np(u), rho=default_rho)
            db.solve()
            r = db.results()
            mk = lambda a: torch.as_tensor(a, dtype=Q.dtype, device=Q.device)  # noqa: E731
            return mk(r["x"]), mk(r["y"]), mk(r["z"])

        @staticmethod
        def backward(ctx, dl_dzhat, dl_dlams, dl_dnus):
            n_batch, dim, neq, nineq = ctx.n_batch, ctx.nz, ctx.neq, ctx.nineq
            if ctx.device_api:
                dev, dt = dl_dzhat.device, dl_dzhat.dtype
                rhs = torch.zeros((n_batch, dim + neq + nineq), dtype=torch.float64, device=dev)  # qplayer.py:197-205
                rhs[:, :dim] = dl_dzhat
                if dl_dlams is not None and neq > 0:
                    rhs[:, dim:dim + neq] = dl_dlams
                if dl_dnus is not None and nineq > 0:
                    rhs[:, dim + neq:] = dl_dnus
                new = lambda *shape: torch.empty(shape, dtype=torch.float64, device=dev)  # noqa: E731
                dQ, dp = new(n_batch, dim, dim), new(n_batch, dim)
                dA, db_ = new(n_batch, neq, dim), new(n_batch, neq)
                dG, du, dl = new(n_batch, nineq, dim), new(n_batch, nineq), new(n_batch, nineq)
                ptr = lambda t: _ct.c_void_p(t.data_ptr()) if t.numel() else None  # noqa: E731
                if rhs.is_cuda:
                    torch.cuda.current_stream(dev).synchronize()
                grp = ctx.batch._g
                _capi.check(grp.lib.pqp_batch_backward_device(grp.handle, 0, n_batch, ptr(rhs), float(eps_backward), float(rho_backward),
                                                             float(mu_backward), ptr(dQ), ptr(dp), ptr(dA), ptr(db_), ptr(dG), ptr(du), ptr(dl)))
                return (dQ.to(dt), dp.to(dt), dA.to(dt) if neq else None, db_.to(dt) if neq else None, dG.to(dt) if nineq else None,
                        dl.to(dt) if nineq else None, du.to(dt) if nineq else None)
            rhs = np.zeros((n_batch, dim + neq + nineq))  # qplayer.py:197-205
            rhs[:, :dim] = _np(dl_dzhat)
            if dl_dlams is not None and neq > 0:
                rhs[:, dim:dim + neq] = _np(dl_dlams)
            if dl_dnus is not None and nineq > 0:
                rhs[:, dim + neq:] = _np(dl_dnus)
            bd = ctx.batch.backward(rhs, eps_backward, rho_backward, mu_backward)
            mk = lambda a: torch.as_tensor(a, dtype=dl_dzhat.dtype, device=dl_dzhat.device)  # noqa: E731
            # qplayer.py:249-251: (dQs, dps, dAs, dbs, dGs, dls, dus)
            dAs = mk(bd["dL_dA"]) if neq > 0 else None
            dbs = mk(bd["dL_db"]) if neq > 0 else None
            dGs = mk(bd["dL_dC"]) if nineq > 0 else None
            dls = mk(bd["dL_dl"]) if nineq > 0 else None
            dus = mk(bd["dL_du"]) if nineq > 0 else None
            return mk(bd["dL_dH"]), mk(bd["dL_dg"]), dAs, dbs, dGs, dls, dus

    return QPFunctionFn.apply
