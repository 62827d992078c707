"""Small workload for compute-sanitizer (racecheck / synccheck / memcheck): a few QPs through every kernel family
(tile kernel, general kernel with box constraints, diagonal Hessian, fused feed, backward)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from proxsuite_b200 import proxqp as px  # noqa: E402


def run(kind, B, n, ne, ni, box=False, hessian=px.HessianType.Dense, sparsity=0.3):
    data = [px.dense.random_qp(kind, i, n, ne, ni, sparsity) for i in range(B)]
    keys = [k for k in ("H", "g", "A", "b", "C", "l", "u", "l_box", "u_box") if k in data[0] and (box or "box" not in k)]
    st = {k: np.stack([d[k] for d in data]) for k in keys}
    db = px.dense.DenseBatch(B, n, ne, ni, box_constraints=box, hessian_type=hessian)
    db.settings.eps_abs = 1e-9
    db.settings.eps_rel = 0
    db.init(**st)
    db.solve()
    r = db.results()
    print(kind, n, "solved", int((r["info"]["status"] == 0).sum()), "/", B, flush=True)


if __name__ == "__main__":
    run("strongly_convex", 4, 20, 6, 12)
    run("box_benchmark", 2, 15, 5, 5, box=True)
    run("diagonal_benchmark", 2, 14, 4, 4, box=True, hessian=px.HessianType.Diagonal)
