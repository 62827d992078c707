"""proxsuite_b200 — B200-native batched dense ProxQP.

Drop-in for ONE path of Simple-Robotics/proxsuite: `proxsuite.proxqp.dense.QP`
/ `BatchQP` / `VectorQP` + `solve_in_parallel` (and the free `dense.solve`),
with the same names, argument meaning and error behaviour
(reference: bindings/python/src/expose-all.cpp:76-123).  Every solve runs the
hand-written sm_100a CUDA kernels behind the C-ABI of include/pqp.h; there is
no CPU fallback.

    from proxsuite_b200 import proxqp
    qp = proxqp.dense.QP(n, n_eq, n_in)
    qp.init(H, g, A, b, C, l, u)
    qp.solve()
"""
from . import proxqp  # noqa: F401

__version__ = "0.1.0"
