"""Pins the oracle (CPU restatement) against the reference's own known-answer
tests: test/src/cvxpy.py:18-74 and test/src/cvxpy.cpp:22-161."""
import numpy as np


def test_cvxpy_3dim_box_qp(oracle):
    # test/src/cvxpy.py:24-46
    H = np.array([[13.0, 12.0, -2.0], [12.0, 17.0, 6.0], [-2.0, 6.0, 12.0]])
    g = np.array([-22.0, -14.5, 13.0])
    qp = oracle.OracleQP(3, 0, 3)
    qp.set(eps_abs=1e-9, eps_rel=0)
    qp.init(H, g, None, None, np.eye(3), -np.ones(3), np.ones(3))
    r = qp.solve()
    assert r.info.status == oracle.PROXQP_SOLVED
    assert np.allclose(r.x, [1.0, 0.5, -1.0], atol=1e-7)
    assert r.info.pri_res <= 1e-9 and r.info.dua_res <= 1e-9


def test_cvxpy_1dim(oracle):
    # test/src/cvxpy.cpp:61-102
    qp = oracle.OracleQP(1, 0, 1)
    qp.set(eps_abs=1e-8)
    qp.init(np.array([[20.0]]), np.array([-10.0]), None, None, np.array([[1.0]]), np.array([0.0]), np.array([1.0]))
    r = qp.solve()
    assert r.info.status == oracle.PROXQP_SOLVED
    assert abs(r.x[0] - 0.5) <= 1e-6


def test_cvxpy_warm_start_at_solution(oracle):
    # test/src/cvxpy.cpp:104-161: start from the solution => no iteration
    H = np.array([[13.0, 12.0, -2.0], [12.0, 17.0, 6.0], [-2.0, 6.0, 12.0]])
    g = np.array([-22.0, -14.5, 13.0])
    qp = oracle.OracleQP(3, 0, 3)
    qp.set(eps_abs=1e-9, eps_rel=0)
    qp.init(H, g, None, None, np.eye(3), -np.ones(3), np.ones(3))
    r = qp.solve()
    qp2 = oracle.OracleQP(3, 0, 3)
    qp2.set(eps_abs=1e-7, eps_rel=0)
    qp2.init(H, g, None, None, np.eye(3), -np.ones(3), np.ones(3))
    r2 = qp2.solve(r.x, r.y, r.z)
    assert r2.info.status == oracle.PROXQP_SOLVED
    assert r2.info.iter <= 0
    assert np.allclose(r2.x, [1.0, 0.5, -1.0], atol=1e-6)


def test_lehmer_stream_is_deterministic(oracle):
    # utils/random_qp_problems.hpp:104-134: uniform in [0,1), reproducible
    a = oracle.lehmer_uniforms(1, 1000)
    b = oracle.lehmer_uniforms(1, 1000)
    c = oracle.lehmer_uniforms(2, 1000)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert (a >= 0).all() and (a < 1).all() and abs(a.mean() - 0.5) < 0.05
