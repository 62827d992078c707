"""Generates tests/golden/*.json (committed fixtures).

Two kinds of fixtures:
  kat_reference.json   known-answer cases copied from the reference's own tests
                       (test/src/cvxpy.py:24-46, test/src/cvxpy.cpp:61-102,
                       test/src/dense_qp_eq.cpp:217-258): inputs + expected x / status.
  oracle_small.json    small seeded QPs (reference generators) with the oracle's
                       solution, iteration counts and Ruiz scaling, so that the
                       GPU box can check parity without /root/reference and so
                       that oracle regressions are caught.
Run from the repo root:  python tests/golden/make_golden.py
(The reference itself cannot be imported or compiled here - no Eigen - so the
fixtures come from its test sources and from the pinned oracle.)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def tolist(d):
    return {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in d.items()}


def main():
    kat = [
        dict(name="cvxpy_3dim_box", source="test/src/cvxpy.py:24-46", n=3, n_eq=0, n_in=3,
             H=[[13.0, 12.0, -2.0], [12.0, 17.0, 6.0], [-2.0, 6.0, 12.0]], g=[-22.0, -14.5, 13.0],
             C=np.eye(3).tolist(), l=[-1.0] * 3, u=[1.0] * 3, eps_abs=1e-9, x=[1.0, 0.5, -1.0], status=0, tol=1e-7),
        dict(name="cvxpy_1dim", source="test/src/cvxpy.cpp:61-102", n=1, n_eq=0, n_in=1, H=[[20.0]], g=[-10.0],
             C=[[1.0]], l=[0.0], u=[1.0], eps_abs=1e-8, x=[0.5], status=0, tol=1e-6),
        dict(name="infeasible_qp", source="test/src/dense_qp_eq.cpp:217-258", n=2, n_eq=0, n_in=3,
             H=[[2.0, 0.0], [0.0, 2.0]], g=[-18.0, -12.0], C=[[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0]],
             l=[-1e30] * 3, u=[10.0, 10.0, -20.0], eps_abs=1e-9, x=None, status=2, tol=0.0),
    ]
    with open(os.path.join(HERE, "kat_reference.json"), "w") as f:
        json.dump(kat, f, indent=1)

    cases = []
    specs = [("strongly_convex", 1, 10, 5, 5, 0.15, False), ("strongly_convex", 2, 30, 10, 20, 0.15, False),
             ("not_strongly_convex", 1, 20, 10, 10, 0.15, False), ("degenerate", 1, 20, 5, 5, 0.15, False),
             ("box_constrained", 1, 15, 0, 15, 0.15, False), ("box_benchmark", 3, 15, 5, 5, 0.5, True),
             ("diagonal_benchmark", 1, 20, 10, 10, 0.5, True)]
    for kind, seed, n, ne, ni, sp, box in specs:
        d = O.generate_qp(kind, seed, n, ne, ni, sp)
        rows = d["C"].shape[0]
        hess = O.HESSIAN_DIAGONAL if kind == "diagonal_benchmark" else O.HESSIAN_DENSE
        qp = O.OracleQP(n, ne, rows, box_constraints=box, hessian_type=hess)
        qp.set(eps_abs=1e-9, eps_rel=0, initial_guess=O.NO_INITIAL_GUESS)
        kw = {k: d[k] for k in "HgAbClu"}
        if box:
            kw.update(l_box=d["l_box"], u_box=d["u_box"])
        qp.init(**kw)
        sc = qp.scaled()
        r = qp.solve()
        cases.append(dict(kind=kind, seed=seed, n=n, n_eq=ne, n_in=rows, gen_n_in=ni, sparsity=sp, box=box, hessian=hess,
                          data=tolist(d), x=r.x.tolist(), y=r.y.tolist(), z=r.z.tolist(), status=r.info.status,
                          iter=r.info.iter, iter_ext=r.info.iter_ext, mu_updates=r.info.mu_updates,
                          delta=sc["delta"].tolist(), c=sc["c"]))
    with open(os.path.join(HERE, "oracle_small.json"), "w") as f:
        json.dump(cases, f)
    print("wrote", len(kat), "KATs and", len(cases), "oracle cases")


if __name__ == "__main__":
    main()
