"""BASELINE cfg 4 (n=256, n_eq=128, n_in=256; 1024 QPs per GPU) through the one-process sharded batch of the C-ABI
(pqp_sharded_*): every visible GPU gets 1024 QPs (weak scaling); prints QP/s for 1 device and for all."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from proxsuite_b200 import proxqp as px  # noqa: E402

per_gpu = int(os.environ.get("PER_GPU", "1024"))
n, ne, ni = 256, 128, 256
ndev = torch.cuda.device_count()
uniq = [px.dense.random_qp("strongly_convex", i, n, ne, ni) for i in range(64)]  # 64 distinct QPs, tiled (host memory)
for devices in ([0], list(range(ndev))):
    B = per_gpu * len(devices)
    st = {k: np.stack([uniq[i % 64][k] for i in range(B)]) for k in "HgAbClu"}
    sb = px.dense.ShardedBatch(B, n, ne, ni, devices=devices)
    sb.settings.eps_abs = 1e-9
    sb.settings.eps_rel = 0
    sb.settings.initial_guess = px.InitialGuess.NO_INITIAL_GUESS
    sb.init(**st)
    sb.solve()
    t0 = time.perf_counter()
    sb.solve()
    dt = time.perf_counter() - t0
    r = sb.results()
    print(f"cfg4 sharded over {len(devices)} GPU(s): B={B} solved {int((r['info']['status'] == 0).sum())}/{B} in {dt * 1e3:.1f} ms -> {B / dt:.0f} QP/s (device-resident re-solve, wall)", flush=True)
