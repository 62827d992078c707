"""Runs the larger Maros-Meszaros problems one by one on the GPU and prints status / iterations / time or the error."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_oracle_maros import check_reference_criteria, problems_large  # noqa: E402
from proxsuite_b200 import proxqp as px  # noqa: E402

for name, d in problems_large():
    n, ne, ni = d["H"].shape[0], d["A"].shape[0], d["C"].shape[0]
    try:
        qp = px.dense.QP(n, ne, ni, False, px.HessianType.Dense, px.DenseBackend.Automatic)
        qp.settings.eps_abs = 2e-8
        qp.settings.eps_rel = 0
        qp.settings.eps_primal_inf = 1e-12
        qp.settings.eps_dual_inf = 1e-12
        qp.init(d["H"], d["g"], d["A"], d["b"], d["C"], d["l"], d["u"])
        t = time.time()
        qp.solve()
        r = qp.results
        ok = "criteria ok"
        try:
            check_reference_criteria(d, r.x, r.y, r.z)
        except AssertionError:
            ok = "CRITERIA FAILED"
        cfg = qp._group
        print(name, n, ne, ni, "status", int(r.info.status), "iter", r.info.iter, r.info.iter_ext, "%.2fs" % (time.time() - t), ok, flush=True)
    except Exception as e:  # noqa: BLE001
        print(name, n, ne, ni, "ERROR", repr(e)[:400], flush=True)
        break
