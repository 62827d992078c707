"""CPU-only checks (no GPU needed): the oracle against the committed golden
fixtures, the C-ABI library loads and exports every symbol include/pqp.h
declares, settings defaults mirror settings.hpp:213-315, host-side argument
validation raises what the reference throws (std::invalid_argument -> ValueError)."""
import json
import os
import re

import numpy as np
import pytest

from helpers import kkt_residuals

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def load(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


def test_oracle_matches_reference_kats(oracle):
    for k in load("kat_reference.json"):
        qp = oracle.OracleQP(k["n"], k["n_eq"], k["n_in"])
        qp.init(np.array(k["H"]), np.array(k["g"]), None, None, np.array(k["C"]), np.array(k["l"]), np.array(k["u"]))
        qp.set(eps_abs=k["eps_abs"], eps_rel=0)
        r = qp.solve()
        assert r.info.status == k["status"], k["name"]
        if k["x"] is not None:
            assert np.allclose(r.x, k["x"], atol=k["tol"]), k["name"]


def test_oracle_reproduces_golden_solutions(oracle):
    for c in load("oracle_small.json"):
        d = {k: np.array(v) for k, v in c["data"].items()}
        qp = oracle.OracleQP(c["n"], c["n_eq"], c["n_in"], box_constraints=c["box"], hessian_type=c["hessian"])
        qp.set(eps_abs=1e-9, eps_rel=0, initial_guess=oracle.NO_INITIAL_GUESS)
        kw = {k: d[k] for k in "HgAbClu"}
        if c["box"]:
            kw.update(l_box=d["l_box"], u_box=d["u_box"])
        qp.init(**kw)
        r = qp.solve()
        assert r.info.status == c["status"] == 0
        assert np.allclose(r.x, c["x"], rtol=0, atol=1e-9), c["kind"]
        pri, dua = kkt_residuals(d, r.x, r.y, r.z)
        assert pri <= 1e-9 and dua <= 1e-9
        # generator is deterministic: regenerated inputs equal the stored ones
        g = oracle.generate_qp(c["kind"], c["seed"], c["n"], c["n_eq"], c["gen_n_in"], c["sparsity"])
        assert np.array_equal(g["H"], d["H"]) and np.array_equal(g["C"], d["C"]) and np.array_equal(g["u"], d["u"])


def test_cabi_library_loads_and_exports_all_declared_symbols():
    from proxsuite_b200 import _capi

    lib = _capi.lib()
    header = open(os.path.join(ROOT, "include", "pqp.h")).read()
    declared = set(re.findall(r"\b(pqp_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/pqp.h but not exported"
    assert set(_capi.EXPORTED_SYMBOLS) <= declared
    assert b"sm_100a" in lib.pqp_version()


def test_settings_defaults_match_reference():
    # settings.hpp:213-315
    from proxsuite_b200 import proxqp

    s = proxqp.Settings()
    assert s.default_rho == 1e-6 and s.default_mu_eq == 1e-3 and s.default_mu_in == 1e-1
    assert s.alpha_bcl == 0.1 and s.beta_bcl == 0.9 and s.mu_min_eq == 1e-9 and s.mu_min_in == 1e-8
    assert s.mu_update_factor == 0.1 and s.mu_update_inv_factor == 10 and s.cold_reset_mu_eq == 1 / 1.1
    assert s.eps_abs == 1e-5 and s.eps_rel == 0 and s.max_iter == 10000 and s.max_iter_in == 1500
    assert s.safe_guard == 10000 and s.nb_iterative_refinement == 10 and s.eps_refact == 1e-6
    assert s.initial_guess == proxqp.InitialGuess.EQUALITY_CONSTRAINED_INITIAL_GUESS
    assert s.compute_preconditioner and not s.update_preconditioner and not s.verbose
    assert s.preconditioner_max_iter == 10 and s.preconditioner_accuracy == 1e-3
    assert s.eps_primal_inf == 1e-4 and s.eps_dual_inf == 1e-4 and s.bcl_update
    assert s.merit_function_type == proxqp.MeritFunctionType.GPDAL and s.alpha_gpdal == 0.95
    assert proxqp.Settings(proxqp.DenseBackend.PrimalLDLT).default_rho == 1e-5
    s.eps_abs = 1e-9
    s.initial_guess = proxqp.InitialGuess.NO_INITIAL_GUESS
    assert s.eps_abs == 1e-9 and s.initial_guess == proxqp.InitialGuess.NO_INITIAL_GUESS
    with pytest.raises(AttributeError):
        s.not_a_setting = 1


def test_backend_choice_matches_reference():
    # wrapper.hpp:82-113; timings-dense-backend.cpp uses n_eq = n_in = 2 n -> PrimalLDLT
    from proxsuite_b200 import _capi

    L = _capi.lib()
    assert L.pqp_dense_backend_choice(0, 100, 50, 100, 0) == 1
    assert L.pqp_dense_backend_choice(0, 10, 20, 20, 0) == 2
    assert L.pqp_dense_backend_choice(2, 100, 50, 100, 0) == 2


def test_generator_in_product_library_matches_oracle_stream(oracle):
    from proxsuite_b200 import proxqp

    a = proxqp.dense.random_qp("strongly_convex", 5, 12, 4, 6)
    b = oracle.generate_qp("strongly_convex", 5, 12, 4, 6)
    for k in a:
        assert np.array_equal(a[k], b[k])
    # strong convexity: lambda_min(H) ~= strong_convexity_factor (random_qp_problems.hpp:326-331)
    assert abs(np.linalg.eigvalsh(a["H"]).min() - 1e-2) < 1e-8


def test_dim_zero_raises_value_error():
    # model.hpp:65-68
    from proxsuite_b200 import proxqp

    with pytest.raises(ValueError):
        proxqp.dense.QP(0, 0, 0)


def test_no_cpu_fallback_without_gpu():
    import torch

    from proxsuite_b200 import proxqp

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        proxqp.dense.QP(3, 0, 3)


def test_free_solve_binds_both_overloads_like_the_reference(monkeypatch):
    """expose-solve.hpp:54-79 / 115-142: (H, g, A, b, C, l, u, x, y, z, eps_abs, ...) and the box overload
    (H, g, A, b, C, l, u, l_box, u_box, x, y, z, eps_abs, ...). The plain overload is tried first and rejected when the
    argument at its eps_abs position is an array (host logic only: the solver behind it is stubbed)."""
    from proxsuite_b200.proxqp import dense

    seen = {}
    monkeypatch.setattr(dense, "_solve", lambda **kw: seen.update(kw) or "results")
    H, g, A, b, C, l, u = (np.eye(3), np.zeros(3), np.ones((1, 3)), np.ones(1), np.eye(3), -np.ones(3), np.ones(3))
    x, y, z = np.zeros(3), np.zeros(1), np.zeros(3)
    assert dense.solve(H, g, A, b, C, l, u, x, y, z, 1e-9, 0) == "results"
    assert seen["x"] is x and seen["y"] is y and seen["z"] is z and seen["eps_abs"] == 1e-9 and seen["eps_rel"] == 0 and "l_box" not in seen
    seen.clear()
    lb, ub, zb = -2 * np.ones(3), 2 * np.ones(3), np.zeros(6)
    dense.solve(H, g, A, b, C, l, u, lb, ub, x, y, zb, 1e-9, 0)  # index 10 is y: an array, so the box overload binds
    assert seen["l_box"] is lb and seen["u_box"] is ub and seen["x"] is x and seen["y"] is y and seen["z"] is zb and seen["eps_abs"] == 1e-9
    seen.clear()
    dense.solve(H, g, A, b, C, l, u, l_box=lb, u_box=ub, eps_abs=1e-7)
    assert seen["l_box"] is lb and seen["eps_abs"] == 1e-7 and "x" not in seen
    with pytest.raises(TypeError):
        dense.solve(H, g, A, b, C, l, u, x, y, z, 1e-9, eps_abs=1e-9)  # twice
    with pytest.raises(TypeError):
        dense.solve(H, g, nonsense=1)
