"""TEST INFRASTRUCTURE: run parity cases of the CUDA kernel SOURCES on the CPU emulator (tests/emu/libpqp_emu.so)
against the oracle. Executed in its own process by tests/test_emu_kernels.py with PQP_B200_LIB pointing at the
emulator library; the package itself never loads it."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
assert os.environ.get("PQP_B200_LIB", "").endswith("libpqp_emu.so"), "run with PQP_B200_LIB=tests/emu/libpqp_emu.so"

from helpers import kkt_residuals  # noqa: E402
from oracle import oracle as O  # noqa: E402
from proxsuite_b200 import proxqp  # noqa: E402

KEYS = "HgAbClu"
EPS = 1e-9


def case(name, kind, B, n, ne, ni, box=False, hessian=proxqp.HessianType.Dense, sparsity=0.15, whole=True, ig=None, layout=None):
    if layout:
        os.environ["PQP_LAYOUT"] = layout
    else:
        os.environ.pop("PQP_LAYOUT", None)
    data = [proxqp.dense.random_qp(kind, i, n, ne, ni, sparsity) for i in range(B)]
    rows = data[0]["C"].shape[0]
    keys = list(KEYS) + (["l_box", "u_box"] if box else [])
    t0 = time.time()
    db = proxqp.dense.DenseBatch(B, n, ne, rows, box, hessian)
    db.settings.eps_abs = EPS
    db.settings.eps_rel = 0
    if ig is not None:
        db.settings.initial_guess = ig
    db.init(**{k: np.stack([d[k] for d in data]) for k in keys})
    if not whole:
        db.scaled(0)  # forces the stand-alone set-up launch instead of the fused feed
    db.solve()
    r = db.results()
    cfg = db.launch_config()
    out = dict(name=name, seconds=time.time() - t0, smem=cfg["smem_bytes"], si_cap=cfg["si_cap"], status=[], iters=[], ok=True)
    for i, d in enumerate(data):
        q = O.OracleQP(n, ne, rows, box_constraints=box, hessian_type=int(hessian))
        kw = dict(eps_abs=EPS, eps_rel=0)
        if ig is not None:
            kw["initial_guess"] = int(ig)
        q.set(**kw)
        q.init(**{k: d[k] for k in keys})
        ro = q.solve()
        st = int(r["info"]["status"][i])
        out["status"].append([st, ro.info.status])
        out["iters"].append([int(r["info"]["iter"][i]), ro.info.iter])
        good = st == ro.info.status
        if st == 0:
            pri, dua = kkt_residuals(d, r["x"][i], r["y"][i], r["z"][i])
            good = good and pri <= EPS and dua <= EPS and np.abs(r["x"][i] - ro.x).max() <= 1e-6 * max(1.0, np.abs(ro.x).max())
        out["ok"] = out["ok"] and bool(good)
    print(json.dumps(out), flush=True)
    return out["ok"]


def backward_case(name, seed, n, ne, ni, sparsity=0.5, with_dy=False):
    """QPLayer backward through the C-ABI on the emulator vs the oracle's compute_backward (same QP, same loss
    derivative), plus the reference's finite-difference acceptance test (test/src/dense_backward.cpp) on dL_dg."""
    os.environ.pop("PQP_LAYOUT", None)
    B = 3
    data = [proxqp.dense.random_qp("strongly_convex", seed + i, n, ne, ni, sparsity, 1e-1) for i in range(B)]
    db = proxqp.dense.DenseBatch(B, n, ne, ni)
    db.settings.eps_abs = EPS
    db.settings.eps_rel = 0
    db.init(**{k: np.stack([d[k] for d in data]) for k in KEYS})
    db.solve()
    r = db.results()
    rng = np.random.default_rng(seed)
    loss = np.zeros((B, n + ne + ni))
    loss[:, :n] = rng.standard_normal((B, n))
    if with_dy:
        loss[:, n:n + ne] = rng.standard_normal((B, ne))
    t0 = time.time()
    bd = db.backward(loss, 1e-9, 1e-7, 1e-7)
    out = dict(name=name, seconds=time.time() - t0, ok=True, maxdiff={})
    for i, d in enumerate(data):
        q = O.OracleQP(n, ne, ni)
        q.set(eps_abs=EPS, eps_rel=0)
        q.init(**{k: d[k] for k in KEYS})
        ro = q.solve()
        bo = q.backward(loss[i], 1e-9, 1e-7, 1e-7)
        for k in bo:
            diff = float(np.abs(bd[k][i] - bo[k]).max()) if bo[k].size else 0.0
            scale = max(1.0, float(np.abs(bo[k]).max())) if bo[k].size else 1.0
            out["maxdiff"][k] = max(out["maxdiff"].get(k, 0.0), diff / scale)
            out["ok"] = out["ok"] and diff <= 1e-7 * scale
        assert int(r["info"]["status"][i]) == ro.info.status == 0
    # finite differences of w.x* w.r.t. g through the emulated forward solve (QP 0)
    d = data[0]
    w = loss[0, :n]
    fd = np.zeros(n)
    for j in range(n):
        xs = []
        for sgn in (+1, -1):
            g2 = d["g"].copy()
            g2[j] += sgn * 1e-5
            q1 = proxqp.dense.QP(n, ne, ni)
            q1.settings.eps_abs = EPS
            q1.settings.eps_rel = 0
            q1.init(d["H"], g2, d["A"], d["b"], d["C"], d["l"], d["u"])
            q1.solve()
            xs.append(q1.results.x.copy())
        fd[j] = w @ (xs[0] - xs[1]) / 2e-5
    if not with_dy:
        out["fd_diff"] = float(np.abs(fd - bd["dL_dg"][0]).max())
        out["ok"] = out["ok"] and out["fd_diff"] < 1e-5
    print(json.dumps(out), flush=True)
    return out["ok"]


def backward_api_case():
    """compute_backward / solve_backward_in_parallel mirrors of the reference binding on QP objects."""
    n, ne, ni = 6, 2, 4
    data = [proxqp.dense.random_qp("strongly_convex", 10 + i, n, ne, ni, 0.5, 1e-1) for i in range(3)]
    batch = proxqp.dense.BatchQP(3)
    for d in data:
        qp = batch.init_qp_in_place(n, ne, ni)
        qp.settings.eps_abs = EPS
        qp.settings.eps_rel = 0
        qp.init(*[d[k] for k in KEYS])
    proxqp.dense.solve_in_parallel(batch)
    losses = proxqp.dense.VectorLossDerivatives()
    for i in range(3):
        v = np.zeros(n + ne + ni)
        v[i] = 1.0
        losses.append(v)
    proxqp.dense.solve_backward_in_parallel(None, batch, losses, 1e-9, 1e-7, 1e-7)
    ok = True
    for i, d in enumerate(data):
        single = proxqp.dense.QP(n, ne, ni)
        single.settings.eps_abs = EPS
        single.settings.eps_rel = 0
        single.init(*[d[k] for k in KEYS])
        single.solve()
        proxqp.dense.compute_backward(single, losses[i], 1e-9, 1e-7, 1e-7)
        a, b = batch[i].model.backward_data, single.model.backward_data
        for k in ("dL_dH", "dL_dg", "dL_dA", "dL_db", "dL_dC", "dL_du", "dL_dl"):
            ok = ok and np.array_equal(getattr(a, k), getattr(b, k))
        ok = ok and single.results.info.rho == 1e-7 and single.results.info.mu_in == 1e-7
    try:
        proxqp.dense.compute_backward(proxqp.dense.QP(n, ne, ni), np.zeros(n + ne + ni))
        ok = False  # an unsolved QP must be refused
    except RuntimeError:
        pass
    print(json.dumps(dict(name="backward_api", ok=bool(ok))), flush=True)
    return bool(ok)


def qplayer_case():
    """proxsuite_b200.torch.qplayer.QPFunction (mirror of proxsuite.torch.qplayer): gradients of a loss on the
    solution w.r.t. every parameter against central finite differences through the layer's own forward pass."""
    import torch

    from proxsuite_b200.torch import QPFunction

    torch.manual_seed(0)
    B, n, ne, ni = 2, 5, 2, 4
    data = [proxqp.dense.random_qp("strongly_convex", 21 + i, n, ne, ni, 0.9, 1e-1) for i in range(B)]
    assert all(np.abs(d["A"]).sum(1).min() > 0 for d in data)  # no empty equality row (degenerate: no unique gradient)
    T = {k: torch.tensor(np.stack([d[k] for d in data]), dtype=torch.float64, requires_grad=True) for k in KEYS}
    layer = QPFunction(eps=1e-10, eps_backward=1e-10, rho_backward=1e-8, mu_backward=1e-8)
    w = torch.randn(B, n, dtype=torch.float64)

    def loss_of(par):
        z, lam, nu = layer(par["H"], par["g"], par["A"], par["b"], par["C"], par["l"], par["u"])
        return (w * z).sum()

    loss = loss_of(T)
    loss.backward()
    ok = True
    worst = {}
    t = 1e-6
    rng = np.random.default_rng(1)
    for k in ("H", "g", "A", "b", "C", "u"):
        dv = rng.standard_normal(tuple(T[k].shape))
        if k == "H":
            dv = 0.5 * (dv + np.swapaxes(dv, 1, 2))
        dvt = torch.tensor(dv)
        with torch.no_grad():
            plus = {kk: (T[kk] + t * dvt if kk == k else T[kk]).detach() for kk in KEYS}
            minus = {kk: (T[kk] - t * dvt if kk == k else T[kk]).detach() for kk in KEYS}
            fd = float((loss_of(plus) - loss_of(minus)) / (2 * t))
        an = float((T[k].grad * dvt).sum())
        worst[k] = abs(fd - an)
        ok = ok and abs(fd - an) <= 5e-4 * max(1.0, abs(fd))
    # a parameter without batch dimension is shared: its gradient is the sum over the batch
    Hs = torch.tensor(data[0]["H"], dtype=torch.float64, requires_grad=True)
    z, _, _ = layer(Hs, T["g"].detach(), T["A"].detach(), T["b"].detach(), T["C"].detach(), T["l"].detach(), T["u"].detach())
    (w * z).sum().backward()
    ok = ok and tuple(Hs.grad.shape) == (n, n)
    print(json.dumps(dict(name="qplayer", ok=bool(ok), fd_vs_autograd=worst)), flush=True)
    return bool(ok)


def closest_feasible_case():
    """primal_infeasibility_solving (closest feasible QP, solver.hpp:1572-1595, utils.hpp:241-248): the reference's
    infeasible QP (test/src/dense_qp_eq.cpp:217-258) and a random QP with an inconsistent pair of equality rows; same
    status, x and slacks as the oracle (the multipliers of an infeasible QP diverge with the iteration count)."""
    ok = True
    H = 2 * np.eye(2); g = np.array([-18.0, -12.0]); C = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0]])
    l = np.full(3, -1e20); u = np.array([10.0, 10.0, -20.0])
    n, ne, ni = 8, 4, 5
    d = proxqp.dense.random_qp("strongly_convex", 7, n, ne, ni, 0.6, 1e-1)
    A = d["A"].copy(); b = d["b"].copy()
    A[3] = A[0]; b[3] = b[0] + 1.0
    for (dims, kw) in (((2, 0, 3), dict(H=H, g=g, C=C, l=l, u=u)), ((n, ne, ni), dict(H=d["H"], g=d["g"], A=A, b=b, C=d["C"], l=d["l"], u=d["u"]))):
        qp = proxqp.dense.QP(*dims)
        qp.settings.eps_abs = EPS
        qp.settings.eps_rel = 0
        qp.settings.primal_infeasibility_solving = True
        qp.settings.max_iter = 40  # the reference keeps iterating to max_iter once the closest feasible point is found
        qp.init(kw["H"], kw["g"], kw.get("A"), kw.get("b"), kw["C"], kw["l"], kw["u"])
        qp.solve()
        r = qp.results
        o = O.OracleQP(*dims)
        o.set(eps_abs=EPS, eps_rel=0, primal_infeasibility_solving=1, max_iter=40)
        o.init(**kw)
        ro = o.solve()
        # the reported state (SOLVED / SOLVED_CLOSEST / PRIMAL_INFEASIBLE at max_iter) depends on rounding-level quantities
        ok = ok and int(r.info.status) in (0, 2, 3) and ro.info.status in (0, 2, 3) and np.abs(r.x - ro.x).max() <= 1e-6 * max(1.0, np.abs(ro.x).max())
        ok = ok and np.abs(r.si - ro.si).max() <= 1e-6 and (dims[1] == 0 or np.abs(r.se - ro.se).max() <= 1e-6)
    print(json.dumps(dict(name="closest_feasible", ok=bool(ok))), flush=True)
    return bool(ok)


def qplayer_device_api_case():
    """the same layer through the device-pointer entry points (on the emulator device memory is host memory)"""
    os.environ["PQP_QPLAYER_DEVICE_API"] = "1"
    try:
        return qplayer_case()
    finally:
        os.environ.pop("PQP_QPLAYER_DEVICE_API", None)


def qplayer_infeas_case():
    """QPFunction(structural_feasibility=False) (QPFunctionFn_infeas, qplayer.py:255-610): the body of the GPU test"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("tq", os.path.join(ROOT, "tests", "test_gpu_qplayer_backward.py"))
    tq = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tq)
    try:
        tq.test_torch_qp_layer_closest_feasible_variant(proxqp, O, quick=True)
        ok = True
    except AssertionError as e:
        print("qplayer_infeas failed:", e, flush=True)
        ok = False
    print(json.dumps(dict(name="qplayer_infeas", ok=ok)), flush=True)
    return ok


def sharded_case():
    """pqp_sharded_* behind the C-ABI: shards on the emulated device must reproduce one DenseBatch bit for bit"""
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location("tb", os.path.join(ROOT, "tests", "test_gpu_baseline_configs.py"))
    tb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tb)
    torch.cuda.device_count = lambda: 1
    try:
        tb.test_sharded_batch_behind_the_c_abi(proxqp, O, shape=(7, 10, 3, 6))
        ok = True
    except AssertionError as e:
        print("sharded failed:", e, flush=True)
        ok = False
    print(json.dumps(dict(name="sharded", ok=ok)), flush=True)
    return ok


def big_variant_case(force_kkt=False):
    """BIG variant of the tile body (PQP_LAYOUT=big; with force_kkt its whole-KKT inverse fallback): the GPU tests' bodies"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("tb2", os.path.join(ROOT, "tests", "test_gpu_baseline_configs.py"))
    tb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tb)

    class MP:
        def setenv(self, k, v):
            os.environ[k] = v
    ok = True
    try:
        if force_kkt:
            for args in (("strongly_convex", 20, 6, 12, False, 1, 0.3), ("box_benchmark", 20, 6, 10, True, 1, 0.5), ("not_strongly_convex", 40, 20, 20, False, 1, 0.3)):
                tb.test_whole_kkt_inverse_fallback(proxqp, O, MP(), *args)
        else:
            for args in (("strongly_convex", 20, 6, 12, False, 1, 0.3, True), ("box_benchmark", 20, 6, 10, True, 1, 0.5, True), ("diagonal_benchmark", 24, 6, 6, True, 2, 0.5, True)):
                tb.test_big_variant_of_the_tile_body(proxqp, O, MP(), *args)
    except AssertionError as e:
        print("big variant failed:", e, flush=True)
        ok = False
    finally:
        os.environ.pop("PQP_LAYOUT", None)
        os.environ.pop("PQP_FORCE_KKT", None)
    print(json.dumps(dict(name="big_kkt" if force_kkt else "big_variant", ok=ok)), flush=True)
    return ok


CASES = {
    "big_variant": big_variant_case,
    "big_kkt": lambda: big_variant_case(True),
    "qplayer_infeas": qplayer_infeas_case,
    "sharded": sharded_case,
    "qplayer": qplayer_case,
    "qplayer_device_api": qplayer_device_api_case,
    "closest_feasible": closest_feasible_case,
    "backward_eq": lambda: backward_case("backward_eq", 1, 10, 5, 0, 0.85),
    "backward_mixed": lambda: backward_case("backward_mixed", 3, 8, 3, 6),
    "backward_dy": lambda: backward_case("backward_dy", 5, 8, 3, 6, with_dy=True),
    "backward_api": backward_api_case,
    "tile_small": lambda: case("tile_small", "strongly_convex", 3, 12, 4, 8),
    "tile_plain_setup": lambda: case("tile_plain_setup", "strongly_convex", 2, 12, 4, 8, whole=False),
    "tile_eq_guess": lambda: case("tile_eq_guess", "strongly_convex", 2, 10, 5, 6, ig=proxqp.InitialGuess.EQUALITY_CONSTRAINED_INITIAL_GUESS),
    "tile_box": lambda: case("tile_box", "box_benchmark", 2, 8, 3, 5, box=True, sparsity=0.5),
    "general_odd": lambda: case("general_odd", "strongly_convex", 2, 7, 3, 5),
    "general_diag": lambda: case("general_diag", "diagonal_benchmark", 2, 9, 3, 4, box=True, hessian=proxqp.HessianType.Diagonal, sparsity=0.5),
    "generic_layout": lambda: case("generic_layout", "strongly_convex", 2, 12, 4, 8, layout="generic"),
    "gated_feed": lambda: case("gated_feed", "strongly_convex", 260, 6, 2, 4),  # >= 256 QPs from host buffers: chunked upload + progress word + feed margin
    "few_rows_generic": lambda: case("few_rows_generic", "strongly_convex", 2, 8, 3, 2, layout="generic"),
    "no_inequalities": lambda: case("no_inequalities", "strongly_convex", 2, 10, 5, 0),
    "degenerate": lambda: case("degenerate", "degenerate", 2, 10, 3, 4),
    "not_strongly_convex": lambda: case("not_strongly_convex", "not_strongly_convex", 2, 10, 4, 6),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    ok = True
    for nm in names:
        ok = CASES[nm]() and ok
    sys.exit(0 if ok else 1)
