"""TEST INFRASTRUCTURE: random-shape fuzzing of the kernel sources on the CPU emulator against the oracle.
usage: PQP_B200_LIB=tests/emu/libpqp_emu.so timeout 1200 python tests/emu/fuzz_emu.py [n_cases] [seed]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
assert os.environ.get("PQP_B200_LIB", "").endswith("libpqp_emu.so")
from helpers import kkt_residuals  # noqa: E402
from oracle import oracle as O  # noqa: E402
from proxsuite_b200 import proxqp  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
EPS = 1e-9
bad = 0
t_all = time.time()
for case in range(N):
    n = int(rng.integers(2, 41))
    ne = int(rng.integers(0, n + 1)) if rng.random() < 0.8 else 0
    ni = int(rng.integers(0 if ne > 0 else 1, 2 * n + 1))
    box = bool(rng.random() < 0.3)
    hess = int(rng.choice([1, 1, 1, 2, 0])) if box or ni > 0 else 1
    kind = "strongly_convex"
    if box:
        kind = "diagonal_benchmark" if hess == 2 else "box_benchmark"
    layout = str(rng.choice(["auto", "auto", "generic", "compact"]))
    ig = int(rng.choice([O.NO_INITIAL_GUESS, O.EQUALITY_CONSTRAINED_INITIAL_GUESS]))
    B = int(rng.integers(1, 4))
    sparsity = float(rng.choice([0.15, 0.5, 0.9]))
    if layout == "auto":
        os.environ.pop("PQP_LAYOUT", None)
    else:
        os.environ["PQP_LAYOUT"] = layout
    desc = dict(case=case, n=n, ne=ne, ni=ni, box=box, hess=hess, kind=kind, layout=layout, ig=ig, B=B, sparsity=sparsity)
    try:
        data = [proxqp.dense.random_qp(kind, 1000 * case + i, n, ne, ni, sparsity) for i in range(B)]
        if hess == 0:
            for d in data:
                d["H"] = np.zeros_like(d["H"])
        elif hess == 2 and kind != "diagonal_benchmark":
            for d in data:
                d["H"] = np.diag(np.diag(d["H"]))
        keys = list("HgAbClu") + (["l_box", "u_box"] if box else [])
        db = proxqp.dense.DenseBatch(B, n, ne, ni, box, proxqp.HessianType(hess))
        db.settings.eps_abs = EPS
        db.settings.eps_rel = 0
        db.settings.initial_guess = proxqp.InitialGuess(ig)
        db.settings.max_iter = 200
        db.init(**{k: np.stack([d[k] for d in data]) for k in keys})
        db.solve()
        r = db.results()
        do_bw = (not box) and hess == 1 and bool((r["info"]["status"] == 0).all())
        if do_bw:
            loss = np.zeros((B, n + ne + ni))
            loss[:, :n] = rng.standard_normal((B, n))
            if ne and rng.random() < 0.5:
                loss[:, n:n + ne] = rng.standard_normal((B, ne))
            bd = db.backward(loss, 1e-9, 1e-7, 1e-7)
        for i, d in enumerate(data):
            q = O.OracleQP(n, ne, ni, box_constraints=box, hessian_type=hess)
            q.set(eps_abs=EPS, eps_rel=0, initial_guess=ig, max_iter=200)
            q.init(**{k: d[k] for k in keys})
            ro = q.solve()
            st = int(r["info"]["status"][i])
            ok = st == ro.info.status
            if ok and st == 0:
                pri, dua = kkt_residuals(d, r["x"][i], r["y"][i], r["z"][i])
                ok = pri <= EPS and dua <= EPS
                if hess == 1:  # strictly convex: the solution is unique
                    ok = ok and np.abs(r["x"][i] - ro.x).max() <= 1e-6 * max(1.0, np.abs(ro.x).max())
            if ok and do_bw:
                bo = q.backward(loss[i], 1e-9, 1e-7, 1e-7)
                for k in bo:
                    if bo[k].size and np.abs(bd[k][i] - bo[k]).max() > 1e-6 * max(1.0, np.abs(bo[k]).max()):
                        ok = False
                        desc["backward_key"] = k
            if not ok:
                bad += 1
                print("MISMATCH", json.dumps(desc), "qp", i, "status", st, ro.info.status, "iter", int(r["info"]["iter"][i]), ro.info.iter, flush=True)
    except Exception as e:  # noqa: BLE001
        bad += 1
        print("EXCEPTION", json.dumps(desc), repr(e), flush=True)
print("fuzz done: %d cases, %d bad, %.0f s" % (N, bad, time.time() - t_all), flush=True)
sys.exit(1 if bad else 0)
