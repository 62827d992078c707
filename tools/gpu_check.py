"""Developer check run on the GPU box (through gpurun): GPU path vs oracle on
seeded QPs, Ruiz identity, rough timing. Writes gpurun_out/gpu_check.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import kkt_residuals  # noqa: E402
from oracle import oracle as O  # noqa: E402
from proxsuite_b200 import proxqp  # noqa: E402

KEYS = "HgAbClu"
out = {}


def stack(data, k):
    return np.stack([d[k] for d in data])


def run_case(name, kind, B, n, ne, ni, box=False, hessian=proxqp.HessianType.Dense, eps=1e-9,
             ig=proxqp.InitialGuess.NO_INITIAL_GUESS, sparsity=0.15, compare=True, reps=1):
    data = [proxqp.dense.random_qp(kind, i, n, ne, ni, sparsity) for i in range(B)]
    rows = data[0]["C"].shape[0]
    db = proxqp.dense.DenseBatch(B, n, ne, rows, box, hessian)
    db.settings.eps_abs = eps
    db.settings.eps_rel = 0
    db.settings.initial_guess = ig
    kw = {k_: stack(data, k_) for k_ in KEYS}
    names = dict(H="H", g="g", A="A", b="b", C="C", l="l", u="u")
    if box:
        kw["l_box"] = stack(data, "l_box")
        kw["u_box"] = stack(data, "u_box")
    t0 = time.time()
    db.init(**kw)
    db.sync()
    t_init = time.time() - t0
    best = None
    for _ in range(reps):
        t0 = time.time()
        db.solve()
        t = time.time() - t0
        best = t if best is None else min(best, t)
    res = db.results()
    tm = db.timings()
    st = res["info"]["status"]
    pri = np.zeros(B)
    dua = np.zeros(B)
    for i in range(B):
        pri[i], dua[i] = kkt_residuals(data[i], res["x"][i], res["y"][i], res["z"][i])
    rec = dict(B=B, n=n, ne=ne, ni=rows, box=box, solved=int((st == 0).sum()), max_pri=float(pri.max()), max_dua=float(dua.max()),
               iter_mean=float(res["info"]["iter"].mean()), iter_ext_mean=float(res["info"]["iter_ext"].mean()),
               mu_updates_mean=float(res["info"]["mu_updates"].mean()), t_init_s=t_init, t_solve_wall_s=best,
               solve_ms_dev=tm["solve_ms"], setup_ms_dev=tm["setup_ms"], qps_per_s_dev=B / (tm["solve_ms"] * 1e-3) if tm["solve_ms"] > 0 else 0,
               launch=db.launch_config())
    if compare:
        nb = min(B, 16)
        xd, its = [], []
        for i in range(nb):
            q = O.OracleQP(n, ne, rows, box_constraints=box, hessian_type=int(hessian))
            q.set(eps_abs=eps, eps_rel=0, initial_guess=int(ig))
            kk = {k_: data[i][k_] for k_ in KEYS}
            if box:
                kk.update(l_box=data[i]["l_box"], u_box=data[i]["u_box"])
            q.init(**kk)
            r = q.solve()
            xd.append(float(np.abs(r.x - res["x"][i]).max() / max(1.0, np.abs(r.x).max())))
            its.append((r.info.iter, int(res["info"]["iter"][i]), r.info.iter_ext, int(res["info"]["iter_ext"][i]), r.info.status, int(st[i])))
        rec["x_rel_diff_max"] = max(xd)
        rec["iters_oracle_vs_gpu"] = its[:8]
    out[name] = rec
    print(name, json.dumps(rec))
    return db, data, res


def ruiz_check():
    n, ne, ni = 40, 20, 20
    d = proxqp.dense.random_qp("strongly_convex", 1, n, ne, ni)
    qp = proxqp.dense.QP(n, ne, ni)
    qp.init(*[d[k] for k in KEYS])
    s = qp.scaled()
    oq = O.OracleQP(n, ne, ni)
    oq.init(**{k: d[k] for k in KEYS})
    so = oq.scaled()
    rec = {k: float(np.abs(s[k] - so[k]).max()) for k in ("H", "g", "A", "b", "C", "u", "l", "delta")}
    rec["c"] = abs(s["c"] - so["c"])
    out["ruiz_vs_oracle"] = rec
    print("ruiz", rec)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    try:
        if which in ("all", "small"):
            ruiz_check()
            run_case("tiny", "strongly_convex", 4, 10, 5, 5)
            run_case("tiny_eqguess", "strongly_convex", 4, 10, 5, 5, ig=proxqp.InitialGuess.EQUALITY_CONSTRAINED_INITIAL_GUESS)
            run_case("small", "strongly_convex", 16, 30, 10, 20)
        if which in ("all", "mid"):
            run_case("cfg2_64", "strongly_convex", 64, 100, 50, 100)
            run_case("box_small", "box_benchmark", 8, 15, 5, 5, box=True, sparsity=0.5)
            run_case("degenerate", "degenerate", 8, 20, 5, 5)
            run_case("not_strongly_convex", "not_strongly_convex", 8, 20, 10, 10)
        if which == "prof":
            db, data, res = run_case("cfg2_prof", "strongly_convex", 1024, 100, 50, 100, compare=False, reps=1)
            pr = db.profile()
            tot = pr["total"]
            out["profile_cycles_per_qp"] = {k: v / 1024 for k, v in pr.items()}
            print("cycles per QP:", {k: round(v / 1024) for k, v in pr.items()})
            print("share:", {k: round(v / tot, 3) for k, v in pr.items()})
        if which in ("all", "perf"):
            run_case("cfg2_%d" % int(os.environ.get("PERF_B", "1024")), "strongly_convex", int(os.environ.get("PERF_B", "1024")), 100, 50, 100, compare=False, reps=3)
    finally:
        with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
            json.dump(out, f, indent=1)
