"""cfg 5 (diagonal Hessian, n = 500, box): Newton / outer / mu-update counters of the GPU path next to the oracle's, per seed."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from proxsuite_b200 import proxqp as px  # noqa: E402
from oracle import oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n, ne, ni = 500, 250, 250
data = [O.generate_qp("diagonal_benchmark", i, n, ne, ni, 0.75) for i in range(B)]
keys = ["H", "g", "A", "b", "C", "l", "u", "l_box", "u_box"]
st = {k: np.stack([d[k] for d in data]) for k in keys}
db = px.dense.DenseBatch(B, n, ne, ni, box_constraints=True, hessian_type=px.HessianType.Diagonal)
db.settings.eps_abs = 1e-9
db.settings.eps_rel = 0
db.settings.initial_guess = px.InitialGuess.NO_INITIAL_GUESS
db.init(**st)
db.solve()
r = db.results()
for i, d in enumerate(data):
    qo = O.OracleQP(n, ne, ni, box_constraints=True, hessian_type=2)
    qo.set(eps_abs=1e-9, eps_rel=0, initial_guess=O.NO_INITIAL_GUESS)
    qo.init(**{k: d[k] for k in keys})
    ro = qo.solve()
    inf = r["info"]
    print(i, "gpu", int(inf["status"][i]), int(inf["iter"][i]), int(inf["iter_ext"][i]), int(inf["mu_updates"][i]),
          "oracle", ro.info.status, ro.info.iter, ro.info.iter_ext, ro.info.mu_updates, "dx %.2e" % np.abs(r["x"][i] - ro.x).max(), flush=True)
