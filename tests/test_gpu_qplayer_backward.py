"""GPU tests of the QPLayer backward pass: pqp_batch_backward through the C-ABI and the
Python mirrors of proxsuite.proxqp.dense.{compute_backward, solve_backward_in_parallel} against the oracle's
restatement of dense/compute_ECJ.hpp (same seeded QPs, same loss derivatives; tolerance 1e-7 relative on every
Jacobian) and against the reference's own acceptance test (test/src/dense_backward.cpp: finite differences < 1e-5)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KEYS = "HgAbClu"
EPS = 1e-9
BW = ("dL_dH", "dL_dg", "dL_dA", "dL_db", "dL_dC", "dL_du", "dL_dl")


@pytest.fixture(scope="module")
def px():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from proxsuite_b200 import proxqp

    return proxqp


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as O

    O.build()
    return O


@pytest.mark.parametrize("B,n,ne,ni,with_dy", [(16, 10, 5, 0, False), (32, 20, 8, 12, False), (8, 30, 10, 30, True), (4, 12, 3, 2, False)])
def test_backward_matches_the_oracle(px, oracle, B, n, ne, ni, with_dy):
    data = [px.dense.random_qp("strongly_convex", 100 + i, n, ne, ni, 0.5, 1e-1) for i in range(B)]
    db = px.dense.DenseBatch(B, n, ne, ni)
    db.settings.eps_abs = EPS
    db.settings.eps_rel = 0
    db.init(**{k: np.stack([d[k] for d in data]) for k in KEYS})
    db.solve()
    assert (db.results()["info"]["status"] == 0).all()
    rng = np.random.default_rng(B)
    loss = np.zeros((B, n + ne + ni))
    loss[:, :n] = rng.standard_normal((B, n))
    if with_dy:
        loss[:, n:n + ne] = rng.standard_normal((B, ne))
    bd = db.backward(loss, 1e-9, 1e-7, 1e-7)
    for i, d in enumerate(data):
        q = oracle.OracleQP(n, ne, ni)
        q.set(eps_abs=EPS, eps_rel=0)
        q.init(**{k: d[k] for k in KEYS})
        assert q.solve().info.status == 0
        bo = q.backward(loss[i], 1e-9, 1e-7, 1e-7)
        for k in BW:
            if bo[k].size:
                assert np.abs(bd[k][i] - bo[k]).max() <= 1e-7 * max(1.0, np.abs(bo[k]).max()), (i, k)


def test_backward_against_finite_differences_and_api_mirrors(px):
    # test/src/dense_backward.cpp:16-86: dx/dg from n backward passes vs central differences of re-solved QPs
    n, ne, ni = 10, 5, 0
    d = px.dense.random_qp("strongly_convex", 1, n, ne, ni, 0.85, 1e-1)

    def solved(g):
        qp = px.dense.QP(n, ne, ni)
        qp.settings.eps_abs = EPS
        qp.settings.eps_rel = 0
        qp.init(d["H"], g, d["A"], d["b"], None, None, None)
        qp.solve()
        assert int(qp.results.info.status) == 0
        return qp

    qp = solved(d["g"])
    dx_dg = np.zeros((n, n))
    ld = np.zeros(n + ne + ni)
    for i in range(n):
        ld[i] = 1.0
        px.dense.compute_backward(qp, ld, 1e-5, 1e-7, 1e-7)
        dx_dg[i] = qp.model.backward_data.dL_dg
        ld[i] = 0.0
    assert qp.results.info.rho == 1e-7 and qp.results.info.mu_eq == 1e-7  # compute_ECJ.hpp:66-68
    fd = np.zeros((n, n))
    for i in range(n):
        gp, gm = d["g"].copy(), d["g"].copy()
        gp[i] += 1e-5
        gm[i] -= 1e-5
        fd[:, i] = (solved(gp).results.x - solved(gm).results.x) / 2e-5
    assert np.abs(fd - dx_dg).max() < 1e-5
    # batch entry point == single-QP entry point, bitwise
    batch = px.dense.BatchQP(3)
    for _ in range(3):
        q = batch.init_qp_in_place(n, ne, ni)
        q.settings.eps_abs = EPS
        q.settings.eps_rel = 0
        q.init(d["H"], d["g"], d["A"], d["b"], None, None, None)
    px.dense.solve_in_parallel(batch)
    losses = px.dense.VectorLossDerivatives()
    for i in range(3):
        v = np.zeros(n + ne + ni)
        v[i] = 1.0
        losses.append(v)
    px.dense.solve_backward_in_parallel(None, batch, losses, 1e-5, 1e-7, 1e-7)
    for i in range(3):
        assert np.array_equal(batch[i].model.backward_data.dL_dg, dx_dg[i])
    with pytest.raises(RuntimeError):  # an unsolved QP is refused (PQP_ESTATE)
        px.dense.compute_backward(px.dense.QP(n, ne, ni), np.zeros(n + ne + ni))


def test_torch_qp_layer_on_cuda_tensors(px):
    """proxsuite_b200.torch.QPFunction with CUDA tensors: outputs and gradients live on the device, gradients
    match central finite differences through the layer (the reference's layer: qplayer.py:93-253)."""
    import torch

    from proxsuite_b200.torch import QPFunction

    B, n, ne, ni = 4, 12, 4, 8
    data = [px.dense.random_qp("strongly_convex", 21 + i, n, ne, ni, 0.9, 1e-1) for i in range(B)]
    T = {k: torch.tensor(np.stack([d[k] for d in data]), dtype=torch.float64, device="cuda", requires_grad=True) for k in KEYS}
    layer = QPFunction(eps=1e-10, eps_backward=1e-10, rho_backward=1e-8, mu_backward=1e-8)
    w = torch.randn(B, n, dtype=torch.float64, device="cuda")

    def loss_of(par):
        z, lam, nu = layer(par["H"], par["g"], par["A"], par["b"], par["C"], par["l"], par["u"])
        assert z.is_cuda and lam.shape == (B, ne) and nu.shape == (B, ni)
        return (w * z).sum()

    loss_of(T).backward()
    rng = np.random.default_rng(1)
    t = 1e-6
    for k in ("H", "g", "A", "b", "C", "u"):
        assert T[k].grad.is_cuda and T[k].grad.shape == T[k].shape
        dv = rng.standard_normal(tuple(T[k].shape))
        if k == "H":
            dv = 0.5 * (dv + np.swapaxes(dv, 1, 2))
        dvt = torch.tensor(dv, device="cuda")
        with torch.no_grad():
            plus = {kk: (T[kk] + t * dvt if kk == k else T[kk]).detach() for kk in KEYS}
            minus = {kk: (T[kk] - t * dvt if kk == k else T[kk]).detach() for kk in KEYS}
            fd = float((loss_of(plus) - loss_of(minus)) / (2 * t))
        an = float((T[k].grad * dvt).sum())
        assert abs(fd - an) <= 5e-4 * max(1.0, abs(fd)), (k, fd, an)


def _infeas_layer_problem(rng, nb, nz, neq, nin, infeasible):
    L = rng.standard_normal((nb, nz, nz))
    Q = np.einsum("bij,bkj->bik", L, L) + 0.5 * np.eye(nz)
    p = rng.standard_normal((nb, nz))
    A = rng.standard_normal((nb, neq, nz))
    xs = rng.standard_normal((nb, nz))
    b = np.einsum("bij,bj->bi", A, xs)
    G = rng.standard_normal((nb, nin, nz))
    gx = np.einsum("bij,bj->bi", G, xs)
    l = gx - rng.uniform(0.05, 0.4, (nb, nin))
    u = gx + rng.uniform(0.05, 0.4, (nb, nin))
    if infeasible:  # two parallel inequality rows that cannot both hold
        G[:, 1] = G[:, 0]
        l[:, 1] = u[:, 0] + 1.0
        u[:, 1] = u[:, 0] + 2.0
    return dict(Q=Q, p=p, A=A, b=b, G=G, l=l, u=u)


def test_torch_qp_layer_closest_feasible_variant(px, oracle, quick=False):
    """QPFunction(structural_feasibility=False) == the reference's QPFunctionFn_infeas (qplayer.py:255-610).
    (a) On feasible QPs its outputs and gradients equal the feasible layer's and central finite differences.
    (b) On infeasible QPs the forward pass returns the oracle's closest-feasible solution (primal_infeasibility_solving
    on the single-sided QP) with non-zero slacks s_i, and the gradients match finite differences through the layer."""
    import torch

    from proxsuite_b200.torch import QPFunction

    rng = np.random.default_rng(1)
    nb, nz, neq, nin = 2, 6, 2, 3
    keys = ("Q", "p", "A", "b", "G", "l", "u")
    for infeasible in (False, True):
        d = _infeas_layer_problem(rng, nb, nz, neq, nin, infeasible)
        T = {k: torch.tensor(d[k], dtype=torch.float64, requires_grad=True) for k in keys}
        w = torch.tensor(rng.standard_normal((nb, nz)))
        layer = QPFunction(eps=1e-10, maxIter=60 if infeasible else 200, eps_backward=1e-10, structural_feasibility=False)

        def loss_of(t):
            z, lam, nu, se, si = layer(*[t[k] for k in keys])
            assert z.shape == (nb, nz) and lam.shape == (nb, neq) and nu.shape == (nb, nin) and se.shape == (nb, neq) and si.shape == (nb, nin)
            return (w * z).sum(), z, si

        loss, z, si = loss_of(T)
        loss.backward()
        if not infeasible:
            T2 = {k: v.detach().clone().requires_grad_(True) for k, v in T.items()}
            z2, _, _ = QPFunction(eps=1e-10, maxIter=200, eps_backward=1e-10)(*[T2[k] for k in keys])
            (w * z2).sum().backward()
            assert float((z - z2).abs().max()) <= 1e-7 and float(si.abs().max()) <= 1e-7
            for k in keys:
                assert float((T[k].grad - T2[k].grad).abs().max()) <= 1e-4 * max(1.0, float(T2[k].grad.abs().max())), k
        else:
            assert float(si.abs().max()) > 0.1  # the QP is infeasible: the shifted constraints are what is solved
            for i in range(nb):  # forward parity with the oracle on the single-sided QP
                G2 = np.concatenate((-d["G"][i], d["G"][i]))
                h = np.concatenate((-d["l"][i], d["u"][i]))
                qo = oracle.OracleQP(nz, neq, 2 * nin)
                qo.set(eps_abs=1e-10, eps_rel=0, primal_infeasibility_solving=1, max_iter=60, max_iter_in=100, default_rho=5e-5, refactor_rho_threshold=5e-5)
                qo.init(H=d["Q"][i], g=d["p"][i], A=d["A"][i], b=d["b"][i], C=G2, l=np.full(2 * nin, -1e20), u=h, rho=5e-5)
                ro = qo.solve()
                assert np.abs(z[i].detach().numpy() - ro.x).max() <= 1e-6 * max(1.0, np.abs(ro.x).max())
        t = 1e-6
        for k in (("p",) if quick else ("p", "b", "u") + (("G", "l") if not infeasible else ())):
            dv = torch.tensor(rng.standard_normal(tuple(T[k].shape)))
            plus = {kk: (T[kk].detach() + t * dv if kk == k else T[kk].detach()) for kk in keys}
            minus = {kk: (T[kk].detach() - t * dv if kk == k else T[kk].detach()) for kk in keys}
            fd = float((loss_of(plus)[0] - loss_of(minus)[0]) / (2 * t))
            an = float((T[k].grad * dv).sum())
            assert abs(fd - an) <= 2e-3 * max(1.0, abs(fd)), (infeasible, k, fd, an)
