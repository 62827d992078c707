"""Oracle backward pass (QPLayer row f2 of SURVEY.md section 8): the restatement of
dense::compute_backward (dense/compute_ECJ.hpp:29-190) is pinned by the reference's own
acceptance test, test/src/dense_backward.cpp — Jacobians from the backward pass against
central finite differences of re-solved QPs, |difference| < 1e-5 — on the same seeded
problems. CPU only (no GPU kernel for this row yet)."""
import numpy as np
import pytest

from oracle import oracle as O

EPS_ABS = 1e-9
FD = 1e-5
TOL = 1e-5


def solved(n, ne, ni, H, g, A=None, b=None, C=None, l=None, u=None):
    qp = O.OracleQP(n, ne, ni)
    qp.set(eps_abs=EPS_ABS, eps_rel=0)
    qp.init(H=H, g=g, A=A, b=b, C=C, l=l, u=u)
    r = qp.solve()
    assert r.info.status == 0
    return qp, r


def jacobian_rows(qp, n, ne, ni, key):
    """Row i = d x_i / d(parameter `key`), from n backward passes with unit loss derivatives
    (dense_backward.cpp:44-52: eps 1e-5, rho_new = mu_new = 1e-7)."""
    rows = []
    ld = np.zeros(n + ne + ni)
    for i in range(n):
        ld[i] = 1.0
        rows.append(qp.backward(ld, 1e-5, 1e-7, 1e-7)[key].copy())
        ld[i] = 0.0
    return np.stack(rows)


def test_backward_wrt_g_equality_constrained():
    # dense_backward.cpp:16-86 (seed 1, dim 10, n_eq 5, sparsity 0.85, strong convexity 1e-1)
    n, ne, ni = 10, 5, 0
    d = O.generate_qp("strongly_convex", 1, n, ne, ni, 0.85, 1e-1)
    qp, _ = solved(n, ne, ni, d["H"], d["g"], d["A"], d["b"])
    dx_dg = jacobian_rows(qp, n, ne, ni, "dL_dg")
    fd = np.zeros((n, n))
    for i in range(n):
        gp, gm = d["g"].copy(), d["g"].copy()
        gp[i] += FD
        gm[i] -= FD
        xp = solved(n, ne, ni, d["H"], gp, d["A"], d["b"])[1].x
        xm = solved(n, ne, ni, d["H"], gm, d["A"], d["b"])[1].x
        fd[:, i] = (xp - xm) / (2 * FD)
    assert np.abs(fd - dx_dg).max() < TOL


def test_backward_wrt_b_equality_constrained():
    # dense_backward.cpp:88-146 (strong convexity 1e-2)
    n, ne, ni = 10, 5, 0
    d = O.generate_qp("strongly_convex", 1, n, ne, ni, 0.85, 1e-2)
    qp, _ = solved(n, ne, ni, d["H"], d["g"], d["A"], d["b"])
    dx_db = jacobian_rows(qp, n, ne, ni, "dL_db")
    fd = np.zeros((n, ne))
    for i in range(ne):
        bp, bm = d["b"].copy(), d["b"].copy()
        bp[i] += FD
        bm[i] -= FD
        xp = solved(n, ne, ni, d["H"], d["g"], d["A"], bp)[1].x
        xm = solved(n, ne, ni, d["H"], d["g"], d["A"], bm)[1].x
        fd[:, i] = (xp - xm) / (2 * FD)
    assert np.abs(fd - dx_db).max() < TOL


def test_backward_wrt_g_saturating_inequalities():
    # dense_backward.cpp:148-228 (dim 6, n_in 12, five lower bounds raised to 1e3, u absent)
    n, ne, ni = 6, 0, 12
    d = O.generate_qp("strongly_convex", 1, n, ne, ni, 0.85, 1e-1)
    l = d["l"].copy()
    for k in (0, 1, 2, 3, 9):
        l[k] = 1e3
    qp, r = solved(n, ne, ni, d["H"], d["g"], None, None, d["C"], l, None)
    assert (np.abs(r.z) > 0).sum() >= 1  # some constraints saturate
    dx_dg = jacobian_rows(qp, n, ne, ni, "dL_dg")
    fd = np.zeros((n, n))
    for i in range(n):
        gp, gm = d["g"].copy(), d["g"].copy()
        gp[i] += FD
        gm[i] -= FD
        xp = solved(n, ne, ni, d["H"], gp, None, None, d["C"], l, None)[1].x
        xm = solved(n, ne, ni, d["H"], gm, None, None, d["C"], l, None)[1].x
        fd[:, i] = (xp - xm) / (2 * FD)
    assert np.abs(fd - dx_dg).max() < TOL


def test_backward_outer_products_and_sign_conventions():
    """compute_backward_loss_ESG (compute_ECJ.hpp:127-188): with dL/dx = w the solve gives (dx, dy, dz) and
    dL_dH = sym(dx x^T), dL_dA = dy x^T + y dx^T, dL_db = -dy, dL_dC = dz x^T + z dx^T, dL_du / dL_dl = -dz on the
    active upper / lower rows. Checked against the directional derivative of L = w.x* along random
    perturbations of every parameter."""
    n, ne, ni = 8, 3, 6
    d = O.generate_qp("strongly_convex", 3, n, ne, ni, 0.5, 1e-1)
    qp, r = solved(n, ne, ni, d["H"], d["g"], d["A"], d["b"], d["C"], d["l"], d["u"])
    rng = np.random.default_rng(0)
    w = rng.standard_normal(n)
    bd = qp.backward(np.concatenate([w, np.zeros(ne + ni)]), 1e-9, 1e-9, 1e-9)
    assert np.allclose(bd["dL_dH"], bd["dL_dH"].T)
    assert np.all(bd["dL_du"][np.abs(r.z) == 0] == 0) and np.all(bd["dL_dl"] == 0)  # l = -1e20: never active
    t = 1e-6
    dH = rng.standard_normal((n, n)); dH = 0.5 * (dH + dH.T)
    pert = dict(H=dH, g=rng.standard_normal(n), A=rng.standard_normal((ne, n)), b=rng.standard_normal(ne),
                C=rng.standard_normal((ni, n)), u=rng.standard_normal(ni))
    for key, dv in pert.items():
        dp = {k: (d[k] + t * dv if k == key else d[k]) for k in "HgAbClu"}
        dm = {k: (d[k] - t * dv if k == key else d[k]) for k in "HgAbClu"}
        xp = solved(n, ne, ni, dp["H"], dp["g"], dp["A"], dp["b"], dp["C"], dp["l"], dp["u"])[1].x
        xm = solved(n, ne, ni, dm["H"], dm["g"], dm["A"], dm["b"], dm["C"], dm["l"], dm["u"])[1].x
        fd = w @ (xp - xm) / (2 * t)
        an = float(np.sum(bd["dL_d" + key] * dv))
        assert abs(fd - an) <= 2e-4 * max(1.0, abs(fd)), (key, fd, an)


def test_backward_rejects_dual_infeasible_status_and_bad_size():
    qp = O.OracleQP(4, 0, 0)
    with pytest.raises(ValueError):
        qp.backward(np.zeros(3))
