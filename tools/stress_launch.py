"""Reproducer for launch-sequence problems: back-to-back asynchronous solves on a torch stream (the pattern of
bench.py's timed `value` loop, nvidia-smi sampler running), alternating with end-to-end init + solve cycles."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from proxsuite_b200 import proxqp  # noqa: E402
from bench import ClockSampler  # noqa: E402

loops = int(sys.argv[1]) if len(sys.argv) > 1 else 20
e2e = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B, n, ne, ni = 1024, 100, 50, 100
data = [proxqp.dense.random_qp("strongly_convex", i, n, ne, ni) for i in range(B)]
host = {k: torch.from_numpy(np.stack([d[k] for d in data])).pin_memory().numpy() for k in "HgAbClu"}
db = proxqp.dense.DenseBatch(B, n, ne, ni)
db.settings.eps_abs = 1e-9
db.settings.eps_rel = 0
db.settings.initial_guess = proxqp.InitialGuess.NO_INITIAL_GUESS
db.init(**host)
db.solve()
ref = db.results()["x"].copy()
stream = torch.cuda.Stream()
use_sampler = os.environ.get("STRESS_SAMPLER", "1") == "1"
sync_every = os.environ.get("STRESS_SYNC", "0") == "1"
own_stream = os.environ.get("STRESS_OWN_STREAM", "0") == "1"
sampler = ClockSampler(0)
if use_sampler:
    sampler.start()
t0 = time.time()
bad = 0
for it in range(loops):
    for _ in range(10):
        db.solve_async(None if own_stream else stream.cuda_stream)
        if sync_every:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    db.sync()
    if e2e:
        db.init(**host)
        db.solve()
    x = db.results()["x"]
    if not np.array_equal(x, ref):
        bad += 1
if use_sampler:
    sampler.stop()
print("stress ok: %d loops x (10 async solves%s), %d result mismatches, %.1f s, lib %s" % (loops, " + 1 e2e cycle" if e2e else "", bad, time.time() - t0, os.environ.get("PQP_B200_LIB", "default")))
