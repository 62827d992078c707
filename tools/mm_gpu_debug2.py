"""QSCORPIO and the problems after it, each with a watchdog and the outer-iteration trace (PQP_DEBUG_TRACE=0 PQP_WATCHDOG_MS=...)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_oracle_maros import check_reference_criteria, problems_large  # noqa: E402
from proxsuite_b200 import proxqp as px  # noqa: E402

want = sys.argv[1:] or ["QSCORPIO", "QSCSD1", "QSCTAP1", "QSHARE1B", "QSTAIR", "VALUES"]
np.set_printoptions(linewidth=220, precision=3)
for name, d in problems_large():
    if name not in want:
        continue
    n, ne, ni = d["H"].shape[0], d["A"].shape[0], d["C"].shape[0]
    db = px.dense.DenseBatch(1, n, ne, ni)
    s = db.settings
    s.eps_abs = 2e-8
    s.eps_rel = 0
    s.eps_primal_inf = 1e-12
    s.eps_dual_inf = 1e-12
    db.init(**{k: v[None] for k, v in d.items()})
    t = time.time()
    db.solve()
    r = db.results()
    print(name, n, ne, ni, os.environ.get("PQP_LAYOUT"), "status", r["info"]["status"], "iter", r["info"]["iter"], r["info"]["iter_ext"], "%.2fs" % (time.time() - t), db.launch_config(), flush=True)
    tr = db.debug_trace()
    print(tr[:14], flush=True)
