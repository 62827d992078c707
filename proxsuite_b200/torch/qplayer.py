"""proxsuite.torch.qplayer.QPFunction on the B200 batch path.

Mirror of bindings/python/proxsuite/torch/qplayer.py:12-253 (the feasible variant `QPFunctionFn`): the forward pass
stacks the batch into ONE device-resident DenseBatch (the reference builds a BatchQP object by object and calls
solve_in_parallel, qplayer.py:105-170), the backward pass is ONE pqp_batch_backward launch
(solve_backward_in_parallel, qplayer.py:172-253). Same solver settings as the reference's layer
(qplayer.py:120-127: max_iter_in = 100, default_rho = refactor_rho_threshold = 5e-5, eps_abs = eps,
primal_infeasibility_solving = False) and the same argument / return order.

The closest-feasible variant (`structural_feasibility=False`, QPFunctionFn_infeas, qplayer.py:255-610) is below
as well: forward = the single-sided QP with primal_infeasibility_solving on ONE DenseBatch, backward = the reference's
extended (non-square) KKT system solved in the least-squares sense by a second DenseBatch (the reference hands it to its
sparse backend as `min 0 s.t. kkt w = rhs` with primal_infeasibility_solving; the dense batch path solves the same QP).

Differences: gradients come back in the dtype of the inputs (the reference allocates float32). CUDA float64 tensors never leave
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

    class QPFunctionFn_infeas(Function):
        """qplayer.py:255-610: the QP is re-written with single-sided inequalities [-G; G] x <= [-l; u] and solved with
        primal_infeasibility_solving (closest feasible QP when it is infeasible). Returns
        (zhats, lams, nus_sol, s_e, s_i); nus_sol / s_i are folded back to the double-sided constraints."""

        @staticmethod
        def forward(ctx, Q_, p_, A_, b_, G_, l_, u_):
            n_in = G_.size(-2)
            n_batch = _n_batch(Q_, p_, A_, b_, G_, l_, u_)
            Q, p = _expand(Q_, n_batch, 3), _expand(p_, n_batch, 2)
            G, u, l = _expand(G_, n_batch, 3), _expand(u_, n_batch, 2), _expand(l_, n_batch, 2)
            A, b = _expand(A_, n_batch, 3), _expand(b_, n_batch, 2)
            h = torch.cat((-l, u), dim=1)   # single-sided inequality
            G2 = torch.cat((-G, G), dim=1)
            nz, nineq = Q.size(1), G2.size(1)
            neq = A.size(1) if A is not None else 0
            ctx.neq, ctx.nineq, ctx.nz, ctx.n_batch = neq, nineq, nz, n_batch
            db = _dense.DenseBatch(n_batch, nz, neq, nineq)
            s = db.settings  # qplayer.py:303-311
            s.primal_infeasibility_solving = True
            s.max_iter = maxIter
            s.max_iter_in = 100
            default_rho = 5.0e-5
            s.default_rho = default_rho
            s.refactor_rho_threshold = default_rho
            s.eps_abs = eps
            db.init(H=_np(Q), g=_np(p), A=_np(A), b=_np(b), C=_np(G2), l=np.full((n_batch, nineq), -1.0e20), u=_np(h), rho=default_rho)
            db.solve()
            r = db.results()
            mk = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=Q.dtype, device=Q.device)  # noqa: E731
            zhats, nus = mk(r["x"]), mk(r["z"])
            slacks = -h + torch.bmm(G2, zhats.unsqueeze(2)).squeeze(2)  # G z - h
            nus_sol = mk(-r["z"][:, :n_in] + r["z"][:, n_in:])          # de-projection to the double-sided multiplier
            s_i = mk(-r["si"][:, :n_in] + r["si"][:, n_in:])
            lams = mk(r["y"]) if neq > 0 else torch.empty((n_batch, 0), dtype=Q.dtype, device=Q.device)
            s_e = mk(r["se"]) if neq > 0 else torch.empty((n_batch, 0), dtype=Q.dtype, device=Q.device)
            ctx.lams, ctx.nus, ctx.slacks = lams, nus, slacks
            ctx.save_for_backward(zhats, s_e, Q_, p_, G_, l_, u_, A_, b_)
            return zhats, lams, nus_sol, s_e, s_i

        @staticmethod
        def backward(ctx, dl_dzhat, dl_dlams, dl_dnus, dl_ds_e, dl_ds_i):
            zhats, s_e, Q_, p_, G_, l_, u_, A_, b_ = ctx.saved_tensors
            n_batch, dim, n_eq, n_in = ctx.n_batch, ctx.nz, ctx.neq, ctx.nineq
            n_in_sol = n_in // 2
            shared = lambda t, d: t is not None and t.nelement() > 0 and t.dim() == d - 1  # noqa: E731  (expandParam's flag)
            Q_e, p_e, G_e, A_e, b_e = shared(Q_, 3), shared(p_, 2), shared(G_, 3), shared(A_, 3), shared(b_, 2)
            h_e = shared(l_, 2) or shared(u_, 2)
            Q = _np(_expand(Q_, n_batch, 3))
            G = _np(_expand(G_, n_batch, 3))
            G = np.concatenate((-G, G), axis=1)
            A = _np(_expand(A_, n_batch, 3)) if n_eq > 0 else None
            nus, slacks, lams = _np(ctx.nus), _np(ctx.slacks), _np(ctx.lams)
            zh, se = _np(zhats), _np(s_e)
            # the reference's extended KKT system (qplayer.py:399-470), one per QP, stacked
            n_row = dim + 2 * n_in + (2 * n_eq if n_eq > 0 else 0)
            n_col = 2 * dim + 2 * n_in + ((n_eq + dim) if n_eq > 0 else 0)
            kkt = np.zeros((n_batch, n_row, n_col))
            rhs = np.zeros((n_batch, n_row))
            P2c = np.zeros((n_batch, n_in))
            g = lambda t: None if t is None else _np(t)  # noqa: E731
            dz, dlm, dnu, dse, dsi = g(dl_dzhat), g(dl_dlams), g(dl_dnus), g(dl_ds_e), g(dl_ds_i)
            o_in, o_2 = dim + n_eq, dim + n_eq + n_in      # row / column offsets of the blocks
            o_se, o_si = dim + n_eq + n_in, dim + 2 * n_eq + n_in
            for i in range(n_batch):
                K = kkt[i]
                z_i, s_i = nus[i], slacks[i]
                P_1 = np.minimum(s_i, 0.0) + z_i >= 0.0
                P_2 = s_i <= 0.0
                P2c[i] = np.maximum(s_i, 0.0)
                K[:dim, :dim] = Q[i]
                if n_eq > 0:
                    K[:dim, dim:dim + n_eq] = A[i].T
                    K[dim:dim + n_eq, :dim] = A[i]
                    K[o_se:o_se + n_eq, dim:dim + n_eq] = -np.eye(n_eq)
                    K[o_se:o_se + n_eq, dim + n_eq + 2 * n_in:2 * dim + n_eq + 2 * n_in] = A[i]
                K[:dim, o_in:o_in + n_in] = G[i].T
                K[o_in:o_in + n_in, :dim] = G[i]
                D_1 = P_1.astype(np.float64)
                D_2 = P_2.astype(np.float64)
                K[o_si:, o_in:o_in + n_in] = -np.eye(n_in)
                K[o_in:o_in + n_in, o_2:o_2 + n_in] = np.diag(1.0 - D_1)
                K[o_si:, o_2:o_2 + n_in] = -np.diag(D_1 * D_2)
                dim_ = dim if n_eq > 0 else 0
                K[o_si:, dim + n_eq + 2 * n_in + dim_:] = (1.0 - D_2)[:, None] * G[i]
                r = rhs[i]
                r[:dim] = -dz[i]
                if dlm is not None and n_eq > 0:
                    r[dim:dim + n_eq] = -dlm[i]
                act = -z_i[:n_in_sol] + z_i[n_in_sol:] >= 0
                if dnu is not None and n_in > 0:
                    r[o_in:o_in + n_in_sol][~act] = dnu[i][~act]
                    r[o_in + n_in_sol:o_in + n_in][act] = -dnu[i][act]
                if dse is not None and n_eq > 0 and dse.shape[-1] != 0:
                    r[o_se:o_se + n_eq] = -dse[i]
                if dsi is not None and dsi.shape[-1] != 0:
                    r[o_si:o_si + n_in_sol][~act] = dsi[i][~act]
                    r[o_si + n_in_sol:][act] = -dsi[i][act]
            # min 0 s.t. kkt w = rhs in the closest-feasible sense (qplayer.py:503-535: H = 0, g = 0, no inequalities)
            lsq = _dense.DenseBatch(n_batch, n_col, n_row, 0)
            t = lsq.settings
            t.primal_infeasibility_solving = True
            t.eps_abs = eps_backward
            t.max_iter = 10
            t.default_rho = 1.0e-3
            t.refactor_rho_threshold = 1.0e-3
            lsq.init(H=np.zeros((n_batch, n_col, n_col)), g=np.zeros((n_batch, n_col)), A=kkt, b=rhs, rho=1.0e-3)
            lsq.solve()
            w = lsq.results()["x"]
            dx = w[:, :dim]
            dlam = w[:, dim:dim + n_eq]
            dnu_ = w[:, o_in:o_in + n_in]
            dim_ = dim if n_eq > 0 else 0
            b_5 = w[:, dim + n_eq + 2 * n_in:2 * dim + n_eq + 2 * n_in] if n_eq > 0 else None
            b_6 = w[:, dim + n_eq + 2 * n_in + dim_:]
            bger = lambda a, b: a[:, :, None] * b[:, None, :]  # noqa: E731
            dGs = bger(dnu_, zh) + bger(nus, dx) + bger(P2c, b_6)
            dhs = -dnu_
            dQs = 0.5 * (bger(dx, zh) + bger(zh, dx))
            dps = dx
            dAs = dbs = None
            if n_eq > 0:
                dAs = bger(dlam, zh) + bger(lams, dx) + bger(se, b_5)
                dbs = -dlam
                if A_e:
                    dAs = dAs.mean(0)
                if b_e:
                    dbs = dbs.mean(0)
            # fold the single-sided [-G; G], [-l; u] gradients back (the reference slices `dGs[n_in_sol:, :]`, which is
            # only right after the mean over a shared G; the constraint axis is meant)
            dG_out = dGs[:, n_in_sol:, :] - dGs[:, :n_in_sol, :]
            dl_out = -dhs[:, :n_in_sol]
            du_out = dhs[:, n_in_sol:]
            if G_e:
                dG_out = dG_out.mean(0)
            if h_e:
                dl_out, du_out = dl_out.mean(0), du_out.mean(0)
            if Q_e:
                dQs = dQs.mean(0)
            if p_e:
                dps = dps.mean(0)
            mk = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=dl_dzhat.dtype, device=dl_dzhat.device)  # noqa: E731
            return mk(dQs), mk(dps), mk(dAs), mk(dbs), mk(dG_out), mk(dl_out), mk(du_out)

    return QPFunctionFn.apply if structural_feasibility else QPFunctionFn_infeas.apply
