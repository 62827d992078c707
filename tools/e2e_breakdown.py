"""Where the end-to-end time of one init + solve + results cycle goes (cfg-2 batch, pinned host inputs)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from proxsuite_b200 import proxqp  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n, ne, ni = 100, 50, 100
data = [proxqp.dense.random_qp("strongly_convex", i, n, ne, ni) for i in range(B)]
host = {k: torch.from_numpy(np.stack([d[k] for d in data])).pin_memory().numpy() for k in "HgAbClu"}
db = proxqp.dense.DenseBatch(B, n, ne, ni)
db.settings.eps_abs = 1e-9
db.settings.eps_rel = 0
db.settings.initial_guess = proxqp.InitialGuess.NO_INITIAL_GUESS


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, r


for rep in range(4):
    t_init, _ = timed(lambda: db.init(**host))
    t_solve, _ = timed(db.solve)
    t_res, r = timed(db.results)
    tm = db.timings()
    print(f"rep {rep}: init {t_init:.2f} ms (setup kernel {tm['setup_ms']:.2f})  solve {t_solve:.2f} ms (kernel {tm['solve_ms']:.2f})  results {t_res:.2f} ms  total {t_init + t_solve + t_res:.2f} ms -> {B / (t_init + t_solve + t_res) * 1e3:.0f} QP/s")
h2d = sum(v.nbytes for v in host.values())
t0 = time.perf_counter()
for k, v in host.items():
    torch.from_numpy(v).cuda(non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"plain H2D of the inputs: {dt * 1e3:.2f} ms = {h2d / dt / 1e9:.1f} GB/s")
