"""The oracle against the reference's Maros-Meszaros acceptance test
(test/src/dense_maros_meszaros.cpp:87-165) on the small problems committed under
tests/golden/maros_meszaros_small.npz (built by tests/golden/make_maros_golden.py from the
reference's own .mat files): eps_abs = 2e-8, eps_rel = 0, eps_primal_inf = eps_dual_inf = 1e-12,
DenseBackend::Automatic; dual residual < 2 eps, primal feasibility within eps, and a second solve
with WARM_START_WITH_PREVIOUS_RESULT takes zero iterations."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EPS = 2e-8


def problems():
    z = np.load(os.path.join(HERE, "golden", "maros_meszaros_small.npz"))
    for name in z["names"]:
        yield str(name), {k: z[f"{name}/{k}"] for k in "HgAbClu"}


def problems_large():
    """the other 34 problems the reference's dense test runs (n up to 760), stored sparse"""
    z = np.load(os.path.join(HERE, "golden", "maros_meszaros_large.npz"))
    for name in z["names"]:
        d = {}
        for k in ("H", "A", "C"):
            shape = tuple(int(v) for v in z[f"{name}/{k}_shape"])
            m = np.zeros(shape)
            rc = z[f"{name}/{k}_rc"]
            m[rc[0], rc[1]] = z[f"{name}/{k}_v"]
            d[k] = m
        for k in ("g", "b", "l", "u"):
            d[k] = z[f"{name}/{k}"]
        yield str(name), d


def check_reference_criteria(d, x, y, z):
    ne, ni = d["A"].shape[0], d["C"].shape[0]
    dua = d["H"] @ x + d["g"]
    if ne:
        dua = dua + d["A"].T @ y
    if ni:
        dua = dua + d["C"].T @ z
    assert np.abs(dua).max() < 2 * EPS
    if ne:
        assert np.abs(d["A"] @ x - d["b"]).max() < EPS
    if ni:
        cx = d["C"] @ x
        assert (cx - d["l"]).min() > -EPS and (cx - d["u"]).max() < EPS


def test_oracle_passes_the_reference_maros_meszaros_test(oracle):
    names = []
    for name, d in problems():
        n, ne, ni = d["H"].shape[0], d["A"].shape[0], d["C"].shape[0]
        qp = oracle.OracleQP(n, ne, ni, dense_backend=oracle.BACKEND_AUTOMATIC)
        qp.set(eps_abs=EPS, eps_rel=0.0, eps_primal_inf=1e-12, eps_dual_inf=1e-12)
        qp.init(**d)
        r = qp.solve()
        assert r.info.status == oracle.PROXQP_SOLVED, name
        check_reference_criteria(d, r.x, r.y, r.z)
        qp.set(initial_guess=oracle.WARM_START_WITH_PREVIOUS_RESULT)
        r2 = qp.solve()
        assert r2.info.iter == 0, name
        names.append(name)
    assert len(names) == 28


def test_oracle_passes_the_rest_of_the_reference_maros_meszaros_list(oracle):
    """dense_maros_meszaros.cpp:95-99 skips n > 1000 / > 1000 constraint rows; these are the 34 remaining problems."""
    names = []
    for name, d in problems_large():
        n, ne, ni = d["H"].shape[0], d["A"].shape[0], d["C"].shape[0]
        qp = oracle.OracleQP(n, ne, ni, dense_backend=oracle.BACKEND_AUTOMATIC)
        qp.set(eps_abs=EPS, eps_rel=0.0, eps_primal_inf=1e-12, eps_dual_inf=1e-12)
        qp.init(**d)
        r = qp.solve()
        assert r.info.status == oracle.PROXQP_SOLVED, name
        check_reference_criteria(d, r.x, r.y, r.z)
        names.append(name)
    assert len(names) == 34
