"""Oracle acceptance on the reference's seeded generators, with the reference's
own criteria: optimality residuals recomputed from the
original data <= eps_abs, Ruiz identity, serial == batch, parameter plumbing,
initial-guess / update state machine."""
import numpy as np
import pytest

from helpers import kkt_residuals

KEYS = "HgAbClu"


def solve_dense(oracle, d, eps=1e-9, box=False, **settings):
    n, n_eq, n_in = d["H"].shape[0], d["A"].shape[0], d["C"].shape[0]
    qp = oracle.OracleQP(n, n_eq, n_in, box_constraints=box)
    qp.set(eps_abs=eps, eps_rel=0, **settings)
    kw = {k: d[k] for k in KEYS}
    if box:
        kw.update(l_box=d["l_box"], u_box=d["u_box"])
    qp.init(**kw)
    return qp, qp.solve()


@pytest.mark.parametrize("dim", [10, 35, 60, 110])
def test_strongly_convex_eq_and_in(oracle, dim):
    # test/src/dense_qp_with_eq_and_in.cpp:23-64 (seed 1, sparsity 0.15, eps 1e-9)
    d = oracle.generate_qp("strongly_convex", 1, dim, dim // 4, dim // 4)
    qp, r = solve_dense(oracle, d)
    pri, dua = kkt_residuals(d, r.x, r.y, r.z)
    assert r.info.status == oracle.PROXQP_SOLVED
    assert pri <= 1e-9 and dua <= 1e-9


@pytest.mark.parametrize("dim", [10, 40, 80])
def test_box_constrained_as_C(oracle, dim):
    # dense_qp_with_eq_and_in.cpp:78-115
    d = oracle.generate_qp("box_constrained", 1, dim, 0, dim)
    qp, r = solve_dense(oracle, d)
    pri, dua = kkt_residuals(d, r.x, r.y, r.z)
    assert r.info.status == oracle.PROXQP_SOLVED and pri <= 1e-9 and dua <= 1e-9


@pytest.mark.parametrize("dim", [10, 40, 80])
def test_not_strongly_convex(oracle, dim):
    # dense_qp_with_eq_and_in.cpp:128-165
    d = oracle.generate_qp("not_strongly_convex", 1, dim, dim // 2, dim // 2)
    qp, r = solve_dense(oracle, d)
    pri, dua = kkt_residuals(d, r.x, r.y, r.z)
    assert r.info.status == oracle.PROXQP_SOLVED and pri <= 1e-9 and dua <= 1e-9


@pytest.mark.parametrize("dim", [10, 40, 80])
def test_degenerate(oracle, dim):
    # dense_qp_with_eq_and_in.cpp:178+ (C duplicated)
    d = oracle.generate_qp("degenerate", 1, dim, dim // 4, dim // 4)
    qp, r = solve_dense(oracle, d)
    pri, dua = kkt_residuals(d, r.x, r.y, r.z)
    assert r.info.status == oracle.PROXQP_SOLVED and pri <= 1e-9 and dua <= 1e-9


def test_equality_only_and_unconstrained_and_lp(oracle):
    # test/src/dense_qp_eq.cpp:56-102, dense_unconstrained_qp.cpp:163-209
    d = oracle.generate_qp("strongly_convex", 1, 30, 15, 0)
    qp, r = solve_dense(oracle, d)
    pri, dua = kkt_residuals(d, r.x, r.y, r.z)
    assert r.info.status == oracle.PROXQP_SOLVED and pri <= 1e-9 and dua <= 1e-9
    d = oracle.generate_qp("strongly_convex", 1, 30, 0, 0)
    qp, r = solve_dense(oracle, d)
    assert r.info.status == oracle.PROXQP_SOLVED and kkt_residuals(d, r.x, r.y, r.z)[1] <= 1e-9
    # H = I, g = 0 => x = 0
    qp = oracle.OracleQP(10, 0, 0)
    qp.set(eps_abs=1e-9)
    qp.init(np.eye(10), np.zeros(10))
    r = qp.solve()
    assert r.info.status == oracle.PROXQP_SOLVED and np.abs(r.x).max() <= 1e-9
    # LP: H = 0 (dense_qp_eq.cpp LP case), feasible and bounded through the box C = I
    d = oracle.generate_qp("box_constrained", 3, 20, 5, 20)
    d["H"] = np.zeros((20, 20))
    qp, r = solve_dense(oracle, d)
    pri, dua = kkt_residuals(d, r.x, r.y, r.z)
    assert r.info.status == oracle.PROXQP_SOLVED and pri <= 1e-9 and dua <= 1e-9


def test_ruiz_identity(oracle):
    # test/src/dense_ruiz_equilibration.cpp:15-72
    d = oracle.generate_qp("strongly_convex", 1, 40, 20, 20)
    qp = oracle.OracleQP(40, 20, 20)
    qp.init(**{k: d[k] for k in KEYS})
    s = qp.scaled()
    D = s["delta"][:40]
    E = s["delta"][40:60]
    F = s["delta"][60:80]
    c = s["c"]
    assert np.allclose(s["H"], c * (D[:, None] * d["H"] * D[None, :]), atol=1e-10)
    assert np.allclose(s["g"], c * D * d["g"], atol=1e-10)
    assert np.allclose(s["A"], E[:, None] * d["A"] * D[None, :], atol=1e-10)
    assert np.allclose(s["b"], E * d["b"], atol=1e-10)
    assert np.allclose(s["C"], F[:, None] * d["C"] * D[None, :], atol=1e-10)


def test_box_constraints_z_ordering(oracle):
    # test/src/dense_qp_wrapper.cpp:6803-6900: z = [z_C ; z_box]
    for seed in range(20):
        d = oracle.generate_qp("box_benchmark", seed, 15, 5, 5, sparsity=0.5)
        qp, r = solve_dense(oracle, d, box=True)
        pri, dua = kkt_residuals(d, r.x, r.y, r.z)
        assert r.info.status == oracle.PROXQP_SOLVED and pri <= 1e-9 and dua <= 1e-9, seed


def test_diagonal_hessian_box(oracle):
    d = oracle.generate_qp("diagonal_benchmark", 1, 30, 15, 15, sparsity=0.5)
    n = 30
    qp = oracle.OracleQP(n, 15, 15, box_constraints=True, hessian_type=oracle.HESSIAN_DIAGONAL)
    qp.set(eps_abs=1e-9, eps_rel=0, initial_guess=oracle.NO_INITIAL_GUESS)
    qp.init(**{k: d[k] for k in KEYS}, l_box=d["l_box"], u_box=d["u_box"])
    r = qp.solve()
    pri, dua = kkt_residuals(d, r.x, r.y, r.z)
    assert r.info.status == oracle.PROXQP_SOLVED and pri <= 1e-9 and dua <= 1e-9


def test_serial_equals_batch_bitwise(oracle):
    # test/src/parallel_qp_solve.cpp:19-77 (scaled down: 16 QPs, dim 60)
    B, n, ne, ni = 16, 60, 10, 10
    data = [oracle.generate_qp("strongly_convex", i, n, ne, ni) for i in range(B)]
    b1 = oracle.OracleBatch(B, n, ne, ni)
    b2 = oracle.OracleBatch(B, n, ne, ni)
    for b in (b1, b2):
        for i in range(B):
            q = b[i]
            q.set(eps_abs=1e-9)
            q.init(**{k: data[i][k] for k in KEYS})
    b1.solve_serial()
    b2.solve(max(1, oracle.omp_max_threads() // 2))
    for i in range(B):
        assert np.array_equal(b1[i].results().x, b2[i].results().x)


def test_parameter_plumbing(oracle):
    # test/src/dense_qp_solve.cpp:164 (info.rho == 1e-7 after passing rho)
    d = oracle.generate_qp("strongly_convex", 1, 10, 2, 2)
    qp = oracle.OracleQP(10, 2, 2)
    qp.set(eps_abs=1e-9)
    qp.init(**{k: d[k] for k in KEYS}, rho=1e-7, mu_eq=1e-4)
    r = qp.solve()
    assert r.info.rho == 1e-7 and r.info.status == oracle.PROXQP_SOLVED


def test_initial_guess_modes_and_resolve(oracle):
    # test/src/dense_qp_wrapper.cpp:1372-2960, dense_maros_meszaros.cpp:160-162
    d = oracle.generate_qp("strongly_convex", 1, 20, 5, 10)
    for mode in (oracle.NO_INITIAL_GUESS, oracle.EQUALITY_CONSTRAINED_INITIAL_GUESS,
                 oracle.COLD_START_WITH_PREVIOUS_RESULT, oracle.WARM_START_WITH_PREVIOUS_RESULT):
        qp = oracle.OracleQP(20, 5, 10)
        qp.set(eps_abs=1e-9, eps_rel=0, initial_guess=mode)
        qp.init(**{k: d[k] for k in KEYS})
        r = qp.solve()
        pri, dua = kkt_residuals(d, r.x, r.y, r.z)
        assert r.info.status == oracle.PROXQP_SOLVED and pri <= 1e-9 and dua <= 1e-9
        r2 = qp.solve()  # dirty re-solve
        pri, dua = kkt_residuals(d, r2.x, r2.y, r2.z)
        assert r2.info.status == oracle.PROXQP_SOLVED and pri <= 1e-9 and dua <= 1e-9
        if mode == oracle.WARM_START_WITH_PREVIOUS_RESULT:
            assert r2.info.iter == 0
    # warm start from the solution through solve(x, y, z)
    qp = oracle.OracleQP(20, 5, 10)
    qp.set(eps_abs=1e-9, eps_rel=0)
    qp.init(**{k: d[k] for k in KEYS})
    r3 = qp.solve(r.x, r.y, r.z)
    assert r3.info.status == oracle.PROXQP_SOLVED and r3.info.iter <= 1
    assert qp.get("initial_guess") == oracle.WARM_START  # sticky (helpers.hpp:727)


def test_update_g_and_matrices(oracle):
    # test/src/dense_qp_wrapper.cpp:163-1372
    d = oracle.generate_qp("strongly_convex", 1, 20, 5, 10)
    qp = oracle.OracleQP(20, 5, 10)
    qp.set(eps_abs=1e-9, eps_rel=0)
    qp.init(**{k: d[k] for k in KEYS})
    qp.solve()
    d2 = dict(d)
    d2["g"] = d["g"] + 1.0
    qp.update(g=d2["g"])
    r = qp.solve()
    pri, dua = kkt_residuals(d2, r.x, r.y, r.z)
    assert r.info.status == oracle.PROXQP_SOLVED and pri <= 1e-9 and dua <= 1e-9
    d3 = dict(d2)
    d3["H"] = d["H"] + np.eye(20)
    qp.update(H=d3["H"], update_preconditioner=True)
    r = qp.solve()
    pri, dua = kkt_residuals(d3, r.x, r.y, r.z)
    assert r.info.status == oracle.PROXQP_SOLVED and pri <= 1e-9 and dua <= 1e-9


def test_primal_ldlt_backend(oracle):
    # test/src/dense_qp_wrapper.cpp:7618-7673; timings-dense-backend.cpp (n_eq = n_in = 2n)
    d = oracle.generate_qp("strongly_convex", 1, 10, 4, 20)
    qp = oracle.OracleQP(10, 4, 20, dense_backend=oracle.BACKEND_PRIMAL_LDLT)
    qp.set(eps_abs=1e-9, eps_rel=0)
    qp.init(**{k: d[k] for k in KEYS})
    r = qp.solve()
    pri, dua = kkt_residuals(d, r.x, r.y, r.z)
    assert r.info.status == oracle.PROXQP_SOLVED and pri <= 1e-9 and dua <= 1e-9
    assert oracle.OracleQP(10, 20, 20, dense_backend=oracle.BACKEND_AUTOMATIC).get("dense_backend") == oracle.BACKEND_PRIMAL_LDLT


def test_primal_infeasible_detected(oracle):
    # test/src/dense_qp_eq.cpp:217-258 ("infeasible qp"): x1 <= 10, x2 <= 10, x1 >= 20
    H = 2 * np.eye(2)
    g = np.array([-18.0, -12.0])
    C = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0]])
    u = np.array([10.0, 10.0, -20.0])
    l = np.full(3, -np.inf)
    qp = oracle.OracleQP(2, 0, 3)
    qp.init(H, g, None, None, C, l, u)
    qp.set(eps_rel=0, eps_abs=1e-9)
    r = qp.solve()
    assert r.info.status == oracle.PROXQP_PRIMAL_INFEASIBLE
