"""Short target for ncu captures: one init + N solves of a cfg-2 shaped batch (third argument `cfg4`: n = 256 / 128 / 256)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from proxsuite_b200 import proxqp  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n, ne, ni = 100, 50, 100
if len(sys.argv) > 3 and sys.argv[3] == "cfg4":  # BASELINE configs[3]: the big variant's shape
    n, ne, ni = 256, 128, 256
data = [proxqp.dense.random_qp("strongly_convex", i, n, ne, ni) for i in range(B)]
db = proxqp.dense.DenseBatch(B, n, ne, ni)
db.settings.eps_abs = 1e-9
db.settings.eps_rel = 0
db.settings.initial_guess = proxqp.InitialGuess.NO_INITIAL_GUESS
db.init(**{k: np.stack([d[k] for d in data]) for k in "HgAbClu"})
for _ in range(reps):
    db.solve()
r = db.results()
print("solved", int((r["info"]["status"] == 0).sum()), "of", B, "solve_ms", db.timings()["solve_ms"])
