"""BASELINE.json configs 3-5 (and cfg 2b) through the public API at reduced batch sizes: every QP must
be SOLVED with recomputed residuals <= 1e-9. Prints QP/s of the device-resident solve per shape."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from proxsuite_b200 import proxqp  # noqa: E402

px = proxqp


def run(name, kind, B, n, ne, ni, box=False, hessian=None, sparsity=0.15):
    data = [px.dense.random_qp(kind, i, n, ne, ni, sparsity) for i in range(B)]
    keys = [k for k in ("H", "g", "A", "b", "C", "l", "u", "l_box", "u_box") if k in data[0]]
    st = {k: np.stack([d[k] for d in data]) for k in keys}
    hessian = px.HessianType.Dense if hessian is None else hessian
    db = px.dense.DenseBatch(B, n, ne, ni, box_constraints=box, hessian_type=hessian)
    db.settings.eps_abs = 1e-9
    db.settings.eps_rel = 0
    db.settings.initial_guess = px.InitialGuess.NO_INITIAL_GUESS
    db.init(**st)
    db.solve()
    t0 = time.perf_counter()
    db.solve()
    dt = time.perf_counter() - t0
    r = db.results()
    x, y, z = r["x"], r["y"], r["z"]
    cx = np.einsum("bij,bj->bi", st["C"], x)
    zc = z[:, :ni]
    pri = np.abs(np.einsum("bij,bj->bi", st["A"], x) - st["b"]).max() if ne else 0.0
    pri = max(pri, np.abs(np.maximum(cx - st["u"], 0) + np.minimum(cx - st["l"], 0)).max())
    dua = np.einsum("bij,bj->bi", st["H"], x) + st["g"] + np.einsum("bji,bj->bi", st["C"], zc)
    if ne:
        dua = dua + np.einsum("bji,bj->bi", st["A"], y)
    if box:
        dua = dua + z[:, ni:]
        pri = max(pri, np.abs(np.maximum(x - st["u_box"], 0) + np.minimum(x - st["l_box"], 0)).max())
    cfg = db.launch_config()
    print(f"{name}: B={B} n={n} n_eq={ne} n_in={ni} box={box} solved {int((r['info']['status'] == 0).sum())}/{B} pri {pri:.2e} dua {np.abs(dua).max():.2e} "
          f"iter {r['info']['iter'].mean():.1f} kernel {db.timings()['solve_ms']:.2f} ms -> {B / (db.timings()['solve_ms'] * 1e-3):.0f} QP/s (wall {dt * 1e3:.1f} ms) "
          f"smem {cfg['smem_bytes']} si_cap {cfg['si_cap']} retries {cfg['overflow_retries']}", flush=True)
    if os.environ.get("PQP_PROFILE"):
        pr = db.profile()
        if pr:
            tot = max(pr["total"], 1)
            print("   cycles/QP:", {k: round(v / (2 * B)) for k, v in pr.items()}, flush=True)  # two solves since the batch was created
            print("   share    :", {k: round(v / tot, 3) for k, v in pr.items()}, flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["2b", "3", "4", "5"]
    if "2b" in which:
        run("cfg2b", "strongly_convex", 4096, 100, 50, 100)
    full = os.environ.get("SWEEP_FULL") == "1"  # BASELINE.json's own per-GPU batch sizes (cfg 5: 592 instead of 4096, 2 QPs per CTA)
    if "3" in which:
        run("cfg3", "box_benchmark", 4096 if full else 592, 100, 50, 50, box=True, sparsity=0.75)
    if "4" in which:
        run("cfg4", "strongly_convex", 1024 if full else 296, 256, 128, 256)
    if "5" in which:
        run("cfg5", "diagonal_benchmark", 592 if full else 148, 500, 250, 250, box=True, hessian=px.HessianType.Diagonal, sparsity=0.75)
