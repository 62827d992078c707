"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called
through the C-ABI (proxsuite_b200 -> libpqp_b200.so), against the oracle on
the same seeded inputs, against the committed golden fixtures, and - at
BASELINE.json's full size - through size-independent properties.

Tolerances (fp64, stated by north_star): KKT residuals recomputed from the
ORIGINAL data <= eps_abs = 1e-9; |x_gpu - x_oracle|_inf <= 1e-6 max(1,|x|_inf);
identical status. Nothing here reads /root/reference."""
import json
import os

import numpy as np
import pytest

from helpers import kkt_residuals

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = "HgAbClu"
EPS = 1e-9
XTOL = 1e-6


@pytest.fixture(scope="module")
def px():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from proxsuite_b200 import proxqp

    return proxqp


def gpu_solve(px, d, box=False, hessian=None, **settings):
    n, ne, ni = d["H"].shape[0], d["A"].shape[0], d["C"].shape[0]
    hessian = px.HessianType.Dense if hessian is None else hessian
    qp = px.dense.QP(n, ne, ni, box, hessian)
    qp.settings.eps_abs = EPS
    qp.settings.eps_rel = 0
    for k, v in settings.items():
        setattr(qp.settings, k, v)
    args = [d[k] for k in KEYS]
    if box:
        args += [d["l_box"], d["u_box"]]
    qp.init(*args)
    qp.solve()
    return qp


def oracle_solve(oracle, d, box=False, hessian=1, **settings):
    n, ne, ni = d["H"].shape[0], d["A"].shape[0], d["C"].shape[0]
    qp = oracle.OracleQP(n, ne, ni, box_constraints=box, hessian_type=int(hessian))
    qp.set(eps_abs=EPS, eps_rel=0, **{k: int(v) if not isinstance(v, float) else v for k, v in settings.items()})
    kw = {k: d[k] for k in KEYS}
    if box:
        kw.update(l_box=d["l_box"], u_box=d["u_box"])
    qp.init(**kw)
    return qp, qp.solve()


def assert_parity(d, r, ro):
    assert int(r.info.status) == ro.info.status
    pri, dua = kkt_residuals(d, r.x, r.y, r.z)
    assert pri <= EPS and dua <= EPS, (pri, dua)
    assert np.abs(r.x - ro.x).max() <= XTOL * max(1.0, np.abs(ro.x).max())
    for a, b in ((r.y, ro.y), (r.z, ro.z)):
        if a.size:
            assert np.abs(a - b).max() <= 1e-5 * max(1.0, np.abs(b).max())


def test_reference_known_answers(px):
    with open(os.path.join(HERE, "golden", "kat_reference.json")) as f:
        kats = json.load(f)
    for k in kats:
        qp = px.dense.QP(k["n"], k["n_eq"], k["n_in"])
        qp.init(np.array(k["H"]), np.array(k["g"]), None, None, np.array(k["C"]), np.array(k["l"]), np.array(k["u"]))
        qp.settings.eps_abs = k["eps_abs"]
        qp.settings.eps_rel = 0
        qp.solve()
        assert int(qp.results.info.status) == k["status"], k["name"]
        if k["x"] is not None:
            assert np.allclose(qp.results.x, k["x"], atol=k["tol"]), k["name"]


def test_golden_oracle_fixtures(px):
    with open(os.path.join(HERE, "golden", "oracle_small.json")) as f:
        cases = json.load(f)
    for c in cases:
        d = {k: np.array(v) for k, v in c["data"].items()}
        qp = gpu_solve(px, d, box=c["box"], hessian=px.HessianType(c["hessian"]), initial_guess=px.InitialGuess.NO_INITIAL_GUESS)
        r = qp.results
        assert int(r.info.status) == c["status"]
        assert np.abs(r.x - np.array(c["x"])).max() <= XTOL * max(1.0, np.abs(c["x"]).max()), c["kind"]
        pri, dua = kkt_residuals(d, r.x, r.y, r.z)
        assert pri <= EPS and dua <= EPS
        s = qp.scaled()
        assert np.allclose(s["delta"], c["delta"], rtol=1e-12, atol=0)
        assert abs(s["c"] - c["c"]) <= 1e-12


@pytest.mark.parametrize("kind,n,ne,ni", [
    ("strongly_convex", 10, 5, 5), ("strongly_convex", 35, 8, 8), ("strongly_convex", 110, 27, 27),
    ("strongly_convex", 30, 15, 0), ("strongly_convex", 30, 0, 0), ("strongly_convex", 30, 0, 40),
    ("box_constrained", 40, 0, 40), ("not_strongly_convex", 40, 20, 20), ("degenerate", 40, 10, 10),
])
def test_parity_with_oracle_on_seeded_generators(px, oracle, kind, n, ne, ni):
    # test/src/dense_qp_with_eq_and_in.cpp, dense_qp_eq.cpp, dense_unconstrained_qp.cpp
    for seed in (1, 2):
        d = oracle.generate_qp(kind, seed, n, ne, ni)
        for ig in (px.InitialGuess.NO_INITIAL_GUESS, px.InitialGuess.EQUALITY_CONSTRAINED_INITIAL_GUESS):
            qp = gpu_solve(px, d, initial_guess=ig)
            _, ro = oracle_solve(oracle, d, initial_guess=ig)
            assert_parity(d, qp.results, ro)


def test_ruiz_identity_and_oracle_scaling(px, oracle):
    # test/src/dense_ruiz_equilibration.cpp:63-71
    d = oracle.generate_qp("strongly_convex", 1, 40, 20, 20)
    qp = px.dense.QP(40, 20, 20)
    qp.init(*[d[k] for k in KEYS])
    s = qp.scaled()
    D, E, F, c = s["delta"][:40], s["delta"][40:60], s["delta"][60:80], s["c"]
    assert np.allclose(s["H"], c * (D[:, None] * d["H"] * D[None, :]), atol=1e-10)
    assert np.allclose(s["g"], c * D * d["g"], atol=1e-10)
    assert np.allclose(s["A"], E[:, None] * d["A"] * D[None, :], atol=1e-10)
    assert np.allclose(s["b"], E * d["b"], atol=1e-10)
    assert np.allclose(s["C"], F[:, None] * d["C"] * D[None, :], atol=1e-10)
    oq = oracle.OracleQP(40, 20, 20)
    oq.init(**{k: d[k] for k in KEYS})
    so = oq.scaled()
    for k in ("H", "g", "A", "b", "C", "delta"):
        assert np.allclose(s[k], so[k], rtol=1e-13, atol=1e-15), k
    # identity preconditioner
    qp2 = px.dense.QP(40, 20, 20)
    qp2.init(*[d[k] for k in KEYS], False)
    s2 = qp2.scaled()
    assert np.array_equal(s2["H"], d["H"]) and np.all(s2["delta"] == 1.0) and s2["c"] == 1.0


def test_box_constraints_and_z_ordering(px, oracle):
    # test/src/dense_qp_wrapper.cpp:6803-6900
    for seed in range(6):
        d = oracle.generate_qp("box_benchmark", seed, 15, 5, 5, sparsity=0.5)
        qp = gpu_solve(px, d, box=True)
        _, ro = oracle_solve(oracle, d, box=True)
        assert qp.results.z.shape[0] == 5 + 15
        assert_parity(d, qp.results, ro)


def test_diagonal_hessian_and_lp(px, oracle):
    d = oracle.generate_qp("diagonal_benchmark", 1, 30, 15, 15, sparsity=0.5)
    qp = gpu_solve(px, d, box=True, hessian=px.HessianType.Diagonal, initial_guess=px.InitialGuess.NO_INITIAL_GUESS)
    _, ro = oracle_solve(oracle, d, box=True, hessian=2, initial_guess=0)
    assert_parity(d, qp.results, ro)
    # LP with HessianType::Zero
    d = oracle.generate_qp("box_constrained", 3, 20, 5, 20)
    d["H"] = np.zeros((20, 20))
    qp = gpu_solve(px, d, hessian=px.HessianType.Zero)
    _, ro = oracle_solve(oracle, d, hessian=0)
    assert_parity(d, qp.results, ro)


def test_initial_guess_modes_resolve_and_update(px, oracle):
    # test/src/dense_qp_wrapper.cpp:163-2960; dense_maros_meszaros.cpp:160-162
    d = oracle.generate_qp("strongly_convex", 1, 20, 5, 10)
    for ig in px.InitialGuess:
        if ig == px.InitialGuess.WARM_START:
            continue
        qp = gpu_solve(px, d, initial_guess=ig)
        _, ro = oracle_solve(oracle, d, initial_guess=ig)
        assert_parity(d, qp.results, ro)
        qp.solve()  # dirty re-solve
        r2 = qp.results
        pri, dua = kkt_residuals(d, r2.x, r2.y, r2.z)
        assert int(r2.info.status) == 0 and pri <= EPS and dua <= EPS
        if ig == px.InitialGuess.WARM_START_WITH_PREVIOUS_RESULT:
            assert r2.info.iter == 0
    # warm start from the solution through solve(x, y, z): sticky WARM_START
    qp = gpu_solve(px, d)
    r = qp.results
    qp2 = px.dense.QP(20, 5, 10)
    qp2.settings.eps_abs = EPS
    qp2.settings.eps_rel = 0
    qp2.init(*[d[k] for k in KEYS])
    qp2.solve(r.x, r.y, r.z)
    assert int(qp2.results.info.status) == 0 and qp2.results.info.iter <= 1
    assert qp2.settings.initial_guess == px.InitialGuess.WARM_START
    # update of g, then of H with a fresh preconditioner
    d2 = dict(d)
    d2["g"] = d["g"] + 1.0
    qp.update(g=d2["g"])
    qp.solve()
    oq, _ = oracle_solve(oracle, d)
    oq.update(g=d2["g"])
    ro = oq.solve()
    assert_parity(d2, qp.results, ro)
    d3 = dict(d2)
    d3["H"] = d["H"] + np.eye(20)
    qp.update(H=d3["H"], update_preconditioner=True)
    qp.solve()
    oq.update(H=d3["H"], update_preconditioner=True)
    ro = oq.solve()
    assert_parity(d3, qp.results, ro)


def test_parameter_plumbing_and_settings_variants(px, oracle):
    # test/src/dense_qp_solve.cpp:164 (info.rho == 1e-7)
    d = oracle.generate_qp("strongly_convex", 1, 10, 2, 2)
    qp = px.dense.QP(10, 2, 2)
    qp.settings.eps_abs = EPS
    qp.init(*[d[k] for k in KEYS], True, 1e-7, 1e-4)
    qp.solve()
    assert qp.results.info.rho == 1e-7 and int(qp.results.info.status) == 0
    # PDAL merit function, Martinez update, duality-gap check
    d = oracle.generate_qp("strongly_convex", 2, 25, 8, 12)
    for st in (dict(merit_function_type=px.MeritFunctionType.PDAL), dict(bcl_update=False), dict(check_duality_gap=True)):
        qp = gpu_solve(px, d, **st)
        _, ro = oracle_solve(oracle, d, **st)
        assert_parity(d, qp.results, ro)


def test_primal_infeasibility_is_detected(px):
    # test/src/dense_qp_eq.cpp:217-258
    H = 2 * np.eye(2)
    g = np.array([-18.0, -12.0])
    C = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0]])
    qp = px.dense.QP(2, 0, 3)
    qp.init(H, g, None, None, C, np.full(3, -np.inf), np.array([10.0, 10.0, -20.0]))
    qp.settings.eps_abs = EPS
    qp.settings.eps_rel = 0
    qp.solve()
    assert qp.results.info.status == px.QPSolverOutput.PROXQP_PRIMAL_INFEASIBLE


def test_batch_equals_serial_bitwise(px, oracle):
    # test/src/parallel_qp_solve.cpp:19-133 (scaled down: 24 QPs, dim 60)
    B, n, ne, ni = 24, 60, 10, 10
    data = [oracle.generate_qp("strongly_convex", i, n, ne, ni) for i in range(B)]
    batch = px.dense.BatchQP(B)
    singles = px.dense.VectorQP()
    for d in data:
        for qp in (batch.init_qp_in_place(n, ne, ni), singles.init_qp(n, ne, ni)):
            qp.settings.eps_abs = EPS
            qp.settings.eps_rel = 0
            qp.init(*[d[k] for k in KEYS])
    px.dense.solve_in_parallel(batch)
    for q in singles:
        q.solve()
    assert batch.size() == B
    for i in range(B):
        assert np.array_equal(batch.get(i).results.x, singles[i].results.x)
        assert batch[i].results.info.iter == singles[i].results.info.iter
    px.dense.solve_in_parallel(singles, 4)  # vector overload, num_threads accepted


def test_solving_one_member_leaves_its_siblings_alone(px, oracle):
    """QP<T>::solve() touches that QP only (wrapper.hpp:922-954) and solve_in_parallel(std::vector<QP>&) the listed
    ones (parallel/qp_solve.hpp:17-38), although the members of a BatchQP share one device batch here
    (pqp_batch_select): iteration counts and results of the other members must not change. A BatchQP() built without
    a size (the reference's QP layer does that) must not open one device batch per QP."""
    n, ne, ni = 12, 4, 6
    data = [oracle.generate_qp("strongly_convex", i, n, ne, ni) for i in range(5)]
    batch = px.dense.BatchQP()
    for d in data:
        qp = batch.init_qp_in_place(n, ne, ni)
        qp.settings.eps_abs = EPS
        qp.settings.eps_rel = 0
        qp.init(*[d[k] for k in KEYS])
    assert len({id(q._group) for q in batch}) == 1
    px.dense.solve_in_parallel(batch)
    first = [(q.results.info.iter, q.results.x.copy()) for q in batch]
    assert all(it > 0 for it, _ in first)
    batch[1].settings.initial_guess = px.InitialGuess.WARM_START_WITH_PREVIOUS_RESULT
    batch[1].solve()
    assert batch[1].results.info.iter == 0
    for i in (0, 2, 3, 4):
        assert batch[i].results.info.iter == first[i][0] and np.array_equal(batch[i].results.x, first[i][1])
    batch[0].update(g=data[0]["g"] + 1.0)
    batch[3].update(g=data[3]["g"] - 1.0)
    px.dense.solve_in_parallel([batch[0], batch[3]])
    assert not np.array_equal(batch[0].results.x, first[0][1]) and not np.array_equal(batch[3].results.x, first[3][1])
    for i in (2, 4):
        assert batch[i].results.info.iter == first[i][0] and np.array_equal(batch[i].results.x, first[i][1])
    for i, dg in ((0, 1.0), (3, -1.0)):
        d2 = dict(data[i], g=data[i]["g"] + dg)
        pri, dua = kkt_residuals(d2, batch[i].results.x, batch[i].results.y, batch[i].results.z)
        assert pri <= EPS and dua <= EPS


def test_free_solve_function_and_errors(px, oracle):
    d = oracle.generate_qp("strongly_convex", 1, 12, 4, 6)
    r = px.dense.solve(d["H"], d["g"], d["A"], d["b"], d["C"], d["l"], d["u"], eps_abs=EPS, eps_rel=0)
    pri, dua = kkt_residuals(d, r.x, r.y, r.z)
    assert int(r.info.status) == 0 and pri <= EPS and dua <= EPS
    # the box overload called positionally, in the reference's argument order (expose-solve.hpp:115-126)
    db = oracle.generate_qp("box_benchmark", 2, 12, 4, 6, sparsity=0.5)
    rb = px.dense.solve(db["H"], db["g"], db["A"], db["b"], db["C"], db["l"], db["u"], db["l_box"], db["u_box"], None, np.zeros(4), None, EPS, 0)
    rk = px.dense.solve(db["H"], db["g"], db["A"], db["b"], db["C"], db["l"], db["u"], l_box=db["l_box"], u_box=db["u_box"], eps_abs=EPS, eps_rel=0)
    assert int(rb.info.status) == 0 and rb.z.shape[0] == 6 + 12 and np.abs(rb.x - rk.x).max() <= 1e-7
    assert (rb.x >= db["l_box"] - 1e-8).all() and (rb.x <= db["u_box"] + 1e-8).all()
    qp = px.dense.QP(12, 4, 6)
    with pytest.raises(ValueError):  # wrapper.hpp:380-451
        qp.init(d["H"][:5, :5], d["g"], d["A"], d["b"], d["C"], d["l"], d["u"])
    with pytest.raises(ValueError):  # box inputs on a QP built without box constraints (wrapper.hpp:540-545)
        qp.init(d["H"], d["g"], d["A"], d["b"], d["C"], d["l"], d["u"], np.zeros(12), np.ones(12))


def test_full_size_batch_properties(px):
    """BASELINE.json headline shape (n=100, n_eq=50, n_in=100), B=4096: every QP
    SOLVED with recomputed residuals <= 1e-9; solving the batch twice gives
    bit-identical results (determinism); permuting the batch permutes the results."""
    B, n, ne, ni = 4096, 100, 50, 100
    data = [px.dense.random_qp("strongly_convex", i, n, ne, ni) for i in range(B)]
    st = {k: np.stack([d[k] for d in data]) for k in KEYS}

    def run(order):
        db = px.dense.DenseBatch(B, n, ne, ni)
        db.settings.eps_abs = EPS
        db.settings.eps_rel = 0
        db.settings.initial_guess = px.InitialGuess.NO_INITIAL_GUESS
        db.init(**{k: st[k][order] for k in KEYS})
        db.solve()
        return db, db.results()

    ident = np.arange(B)
    db, r = run(ident)
    assert (r["info"]["status"] == 0).all()
    x, y, z = r["x"], r["y"], r["z"]
    cx = np.einsum("bij,bj->bi", st["C"], x)
    pri = np.maximum(np.abs(np.einsum("bij,bj->bi", st["A"], x) - st["b"]).max(1),
                     np.abs(np.maximum(cx - st["u"], 0) + np.minimum(cx - st["l"], 0)).max(1))
    dua = np.abs(np.einsum("bij,bj->bi", st["H"], x) + st["g"] + np.einsum("bji,bj->bi", st["A"], y) + np.einsum("bji,bj->bi", st["C"], z)).max(1)
    assert pri.max() <= EPS and dua.max() <= EPS
    db.solve()
    r2 = db.results()
    assert np.array_equal(r2["x"], x)
    perm = np.random.default_rng(0).permutation(B)
    _, rp = run(perm)
    assert np.array_equal(rp["x"], x[perm])


def test_compact_layout_overflow_retry(px, oracle, monkeypatch):
    """Compact layout (two CTAs per SM) with a deliberately small S^-1 capacity:
    QPs whose active set outgrows it are re-solved by the generic kernel and
    must give the same answers as the default layout."""
    B, n, ne, ni = 32, 20, 6, 40
    data = [px.dense.random_qp("strongly_convex", i, n, ne, ni) for i in range(B)]
    st = {k: np.stack([d[k] for d in data]) for k in KEYS}

    def run():
        db = px.dense.DenseBatch(B, n, ne, ni)
        db.settings.eps_abs = EPS
        db.settings.eps_rel = 0
        db.settings.initial_guess = px.InitialGuess.NO_INITIAL_GUESS
        db.init(**st)
        db.solve()
        return db.results(), db.launch_config()

    r_ref, cfg_ref = run()
    monkeypatch.setenv("PQP_LAYOUT", "compact")
    monkeypatch.setenv("PQP_SI_CAP", str(ne + 16))
    r_small, cfg_small = run()
    assert cfg_small["si_cap"] == ne + 16
    assert cfg_small["overflow_retries"] > 0, "the test shape must overflow the forced capacity"
    assert (r_small["info"]["status"] == 0).all() and (r_ref["info"]["status"] == 0).all()
    assert np.abs(r_small["x"] - r_ref["x"]).max() <= 1e-7
    for i in range(B):
        pri, dua = kkt_residuals(data[i], r_small["x"][i], r_small["y"][i], r_small["z"][i])
        assert pri <= EPS and dua <= EPS


def test_tile_layout_overflow_retry(px, oracle, monkeypatch):
    """Tile kernel with a deliberately small S^-1 capacity: QPs whose active set
    outgrows it report the internal code and are re-solved by the general kernel."""
    B, n, ne, ni = 32, 20, 6, 40
    data = [px.dense.random_qp("strongly_convex", i, n, ne, ni) for i in range(B)]
    st = {k: np.stack([d[k] for d in data]) for k in KEYS}

    def run():
        db = px.dense.DenseBatch(B, n, ne, ni)
        db.settings.eps_abs = EPS
        db.settings.eps_rel = 0
        db.settings.initial_guess = px.InitialGuess.NO_INITIAL_GUESS
        db.init(**st)
        db.solve()
        return db.results(), db.launch_config()

    r_ref, _ = run()
    monkeypatch.setenv("PQP_LAYOUT", "tile")
    monkeypatch.setenv("PQP_SI_CAP", str(ne + 16))
    r_small, cfg_small = run()
    assert cfg_small["si_cap"] == ne + 16
    assert cfg_small["overflow_retries"] > 0, "the test shape must overflow the forced capacity"
    assert (r_small["info"]["status"] == 0).all() and (r_ref["info"]["status"] == 0).all()
    assert np.abs(r_small["x"] - r_ref["x"]).max() <= 1e-7
    for i in range(B):
        pri, dua = kkt_residuals(data[i], r_small["x"][i], r_small["y"][i], r_small["z"][i])
        assert pri <= EPS and dua <= EPS


@pytest.mark.gpu
def test_fused_feed_equals_separate_launches(px, monkeypatch):
    """End-to-end path: a whole-batch init() from host buffers only uploads (in chunks, a progress word behind
    each) and the solve() issued next runs ONE persistent kernel that waits for each QP's inputs, equilibrates
    it and solves it. Results must be bit-identical to the separate set-up + solve launches (PQP_E2E=plain) and
    to the per-chunk launches (PQP_E2E=chunks); a re-init / update of the same object must not see stale state."""
    B, n, ne, ni = 384, 30, 10, 30
    data = [px.dense.random_qp("strongly_convex", i, n, ne, ni) for i in range(B)]
    st = {k: np.stack([d[k] for d in data]) for k in KEYS}

    def run(db=None):
        if db is None:
            db = px.dense.DenseBatch(B, n, ne, ni)
            db.settings.eps_abs = EPS
            db.settings.eps_rel = 0
            db.settings.initial_guess = px.InitialGuess.NO_INITIAL_GUESS
        l0 = db.timings()["kernel_launches"]
        db.init(**st)
        db.solve()
        return db, db.results(), db.timings()["kernel_launches"] - l0

    db, r_fused, l_fused = run()
    _, r_fused2, _ = run(db)  # same batch object again
    # update of the whole batch (new g, b; stored scaling kept) is deferred into the next solve as well
    g2 = st["g"] * 1.5
    db.update(g=g2)
    db.solve()
    r_upd = db.results()
    monkeypatch.setenv("PQP_E2E", "plain")
    db1, r_one, l_one = run()
    db1.update(g=g2)
    db1.solve()
    r_upd1 = db1.results()
    monkeypatch.setenv("PQP_E2E", "chunks")
    _, r_chunks, l_chunks = run()
    assert l_fused == 1 and l_one == 2 and l_chunks > l_one, (l_fused, l_one, l_chunks)
    assert (r_one["info"]["status"] == 0).all()
    for k in ("x", "y", "z"):
        assert np.array_equal(r_fused[k], r_one[k]) and np.array_equal(r_fused2[k], r_one[k]) and np.array_equal(r_chunks[k], r_one[k])
        assert np.array_equal(r_upd[k], r_upd1[k])
    assert np.array_equal(r_fused["info"]["iter"], r_one["info"]["iter"])
    assert (r_upd["info"]["status"] == 0).all() and not np.array_equal(r_upd["x"], r_one["x"])


@pytest.mark.gpu
def test_fused_feed_odd_shape_and_scaled_query(px):
    """Per-QP arrays that are not multiples of a cache line (n = 7), box constraints, and a scaled() query between
    init() and solve(), which forces the deferred set-up to run as a separate launch."""
    B, n, ne, ni = 300, 7, 3, 5
    data = [px.dense.random_qp("box_benchmark", i, n, ne, ni, 0.5) for i in range(B)]
    keys = list(KEYS) + ["l_box", "u_box"]
    st = {k: np.stack([d[k] for d in data]) for k in keys}

    def make():
        db = px.dense.DenseBatch(B, n, ne, ni, True)
        db.settings.eps_abs = EPS
        db.settings.eps_rel = 0
        db.init(**st)
        return db

    a = make()
    a.solve()
    ra = a.results()
    b = make()
    sc = b.scaled(5)  # flushes the deferred set-up
    b.solve()
    rb = b.results()
    # seed 125 of this family is primal infeasible (the oracle agrees); everything else is solved
    assert np.array_equal(ra["info"]["status"], rb["info"]["status"]) and (ra["info"]["status"] == 0).sum() >= B - 2
    for k in ("x", "y", "z"):
        assert np.array_equal(ra[k], rb[k])
    for i in (0, 5, 63, 64, 299):
        assert ra["info"]["status"][i] == 0
        pri, dua = kkt_residuals(data[i], ra["x"][i], ra["y"][i], ra["z"][i])
        assert pri <= EPS and dua <= EPS, (i, pri, dua)
    D = sc["delta"]
    assert D.shape[0] == n + ne + ni + n
    assert np.allclose(sc["A"], D[n:n + ne, None] * data[5]["A"] * D[None, :n], atol=1e-10)


def test_tile_kernel_box_constraints(px, oracle):
    """Box constraints in the tile kernel (n even: box rows materialised in the
    transposed constraint copy): oracle-identical iteration counts and solution."""
    for seed in range(6):
        d = oracle.generate_qp("box_benchmark", seed, 20, 6, 10, sparsity=0.5)
        qp = gpu_solve(px, d, box=True)
        _, ro = oracle_solve(oracle, d, box=True)
        assert qp.results.z.shape[0] == 10 + 20
        assert_parity(d, qp.results, ro)
        assert qp.results.info.iter == ro.info.iter and qp.results.info.iter_ext == ro.info.iter_ext


def test_maros_meszaros_small_problems(px, oracle):
    """The reference's Maros-Meszaros acceptance test (dense_maros_meszaros.cpp:87-165) on the
    28 small problems of tests/golden/maros_meszaros_small.npz, through the C-ABI: same status
    and objective as the oracle, the reference's residual criteria, and zero iterations on a
    warm restart from the previous result."""
    from test_oracle_maros import EPS as MEPS, check_reference_criteria, problems

    for name, d in problems():
        n, ne, ni = d["H"].shape[0], d["A"].shape[0], d["C"].shape[0]
        qo = oracle.OracleQP(n, ne, ni, dense_backend=oracle.BACKEND_AUTOMATIC)
        qo.set(eps_abs=MEPS, eps_rel=0.0, eps_primal_inf=1e-12, eps_dual_inf=1e-12)
        qo.init(**d)
        ro = qo.solve()
        qp = px.dense.QP(n, ne, ni, False, px.HessianType.Dense, px.DenseBackend.Automatic)
        qp.settings.eps_abs = MEPS
        qp.settings.eps_rel = 0
        qp.settings.eps_primal_inf = 1e-12
        qp.settings.eps_dual_inf = 1e-12
        qp.init(d["H"], d["g"], d["A"], d["b"], d["C"], d["l"], d["u"])
        qp.solve()
        r = qp.results
        assert int(r.info.status) == ro.info.status == 0, name
        check_reference_criteria(d, r.x, r.y, r.z)
        obj_o = 0.5 * ro.x @ d["H"] @ ro.x + d["g"] @ ro.x
        assert abs(r.info.objValue - obj_o) <= 1e-6 * max(1.0, abs(obj_o)), name
        qp.settings.initial_guess = px.InitialGuess.WARM_START_WITH_PREVIOUS_RESULT
        qp.solve()
        assert qp.results.info.iter == 0, name


# problems of the larger set that take tens of seconds on one CTA (thousands of Newton steps, or the whole-KKT fallback
# with a 1100 x 1100 inverse per active-set change): run with PQP_TEST_SLOW=1 (tools/mm_gpu_debug.py prints all of them)
MAROS_SLOW = ("QFORPLAN", "QSCAGR25")


def test_maros_meszaros_rest_of_the_reference_list(px, oracle, monkeypatch):
    """The other 34 problems the reference's dense Maros-Meszaros test runs (n up to 760, up to 856 constraint rows;
    tests/golden/maros_meszaros_large.npz), through the C-ABI: status SOLVED, the reference's residual criteria,
    the oracle's objective. With the 28 small ones: every problem the reference's test does not skip. QSCORPIO (degenerate,
    n = 358, 746 rows) and QSCAGR25 are the two the dual-block inverse cannot handle (cond S ~ 1e13-1e15): the big variant
    switches them to the inverse of the whole KKT matrix on its own and then needs the oracle's Newton steps (69 vs 68,
    488 vs 489)."""
    from test_oracle_maros import EPS as MEPS, check_reference_criteria, problems_large

    slow = os.environ.get("PQP_TEST_SLOW") == "1"
    monkeypatch.setenv("PQP_WATCHDOG_MS", "120000" if slow else "30000")
    done = 0
    for name, d in problems_large():
        if not slow and name in MAROS_SLOW:
            continue
        n, ne, ni = d["H"].shape[0], d["A"].shape[0], d["C"].shape[0]
        qo = oracle.OracleQP(n, ne, ni, dense_backend=oracle.BACKEND_AUTOMATIC)
        qo.set(eps_abs=MEPS, eps_rel=0.0, eps_primal_inf=1e-12, eps_dual_inf=1e-12)
        qo.init(**d)
        ro = qo.solve()
        qp = px.dense.QP(n, ne, ni, False, px.HessianType.Dense, px.DenseBackend.Automatic)
        qp.settings.eps_abs = MEPS
        qp.settings.eps_rel = 0
        qp.settings.eps_primal_inf = 1e-12
        qp.settings.eps_dual_inf = 1e-12
        qp.init(d["H"], d["g"], d["A"], d["b"], d["C"], d["l"], d["u"])
        qp.solve()
        r = qp.results
        assert int(r.info.status) == ro.info.status == 0, (name, int(r.info.status), r.info.iter)
        check_reference_criteria(d, r.x, r.y, r.z)
        obj_o = 0.5 * ro.x @ d["H"] @ ro.x + d["g"] @ ro.x
        assert abs(r.info.objValue - obj_o) <= 1e-5 * max(1.0, abs(obj_o)), name
        done += 1
    assert done == (34 if slow else 32)


@pytest.mark.gpu
def test_repeated_launches_are_stable(px):
    """Regression: 300 back-to-back launches of the persistent kernel on the headline batch (bench.py's timed loop).
    A block-divergence race in the tile kernel's active-set change (thread 0 updated the slot count while late warps
    still read it) used to end in `illegal instruction` about once per 10^5 QP solves; results must also stay
    bit-identical from launch to launch."""
    import torch

    B, n, ne, ni = 1024, 100, 50, 100
    data = [px.dense.random_qp("strongly_convex", i, n, ne, ni) for i in range(B)]
    st = {k: np.stack([d[k] for d in data]) for k in KEYS}
    db = px.dense.DenseBatch(B, n, ne, ni)
    db.settings.eps_abs = EPS
    db.settings.eps_rel = 0
    db.settings.initial_guess = px.InitialGuess.NO_INITIAL_GUESS
    db.init(**st)
    db.solve()
    ref = db.results()["x"].copy()
    stream = torch.cuda.Stream()
    for _ in range(30):
        for _ in range(10):
            db.solve_async(stream.cuda_stream)
        torch.cuda.synchronize()
        db.sync()
        assert np.array_equal(db.results()["x"], ref)
