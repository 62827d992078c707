"""`proxsuite_b200.proxqp.dense` — QP / BatchQP / VectorQP / solve / solve_in_parallel.

Host-side mirror (Python, like the reference's binding layer) of
  proxsuite::proxqp::dense::QP<T>        dense/wrapper.hpp:115-963
  proxsuite::proxqp::dense::BatchQP<T>   dense/wrapper.hpp:1253-1311
  dense::solve_in_parallel               parallel/qp_solve.hpp:17-60
  dense::solve (free function)           dense/wrapper.hpp:1000-1233
as exposed by bindings/python/src/expose-qpobject.hpp:26-232,
expose-qpvector.hpp:19-39, expose-parallel.hpp:24-83, expose-solve.hpp:20.

A QP lives in a device-resident batch of same-shaped QPs (`pqp_batch`, see
include/pqp.h). A stand-alone `QP(...)` is a batch of one; `BatchQP` packs
every `init_qp_in_place(dim, n_eq, n_in)` of the same shape into one batch so
that `solve_in_parallel` is ONE persistent-kernel launch per shape.
"""
from __future__ import annotations

import ctypes as ct
from typing import Iterable, List, Optional, Sequence

import numpy as np

from .. import _capi
from . import DenseBackend, HessianType, Info, InitialGuess, Results, Settings

_VP = ct.c_void_p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_VP)


def _mat(a, rows, cols, what, allow_empty=True):
    """numpy / None -> C-contiguous float64 [rows, cols] or None (absent).

    Size-0 inputs count as absent (wrapper.hpp:380-451); wrong sizes raise
    ValueError like PROXSUITE_CHECK_ARGUMENT_SIZE (macros.hpp:28-35)."""
    if a is None:
        return None
    a = np.asarray(a, dtype=np.float64)
    if a.size == 0 and allow_empty:
        return None
    if a.ndim != 2 or a.shape[0] != rows or a.shape[1] != cols:
        raise ValueError(f"wrong argument size: expected {rows}x{cols} for {what}, got {tuple(a.shape)}")
    return np.ascontiguousarray(a)


def _vec(a, size, what, allow_empty=True):
    if a is None:
        return None
    a = np.asarray(a, dtype=np.float64)
    if a.size == 0 and allow_empty:
        return None
    if a.ndim == 2 and 1 in a.shape:  # column / row vectors are vectors (Eigen::Ref<Vec> accepts n x 1); a k x m block is not
        a = a.reshape(-1)
    if a.ndim != 1 or a.shape[0] != size:
        raise ValueError(f"wrong argument size: expected {size} for {what}, got {a.shape[0] if a.ndim else 0}")
    return np.ascontiguousarray(a)


def _opt_scalar(v):
    if v is None:
        return None, None
    c = ct.c_double(float(v))
    return c, ct.cast(ct.pointer(c), _VP)


class _Group:
    """One device batch of same-shaped QPs (owns a `pqp_batch*`)."""

    def __init__(self, capacity, n, n_eq, n_in, box, hessian, backend, device=-1):
        self.lib = _capi.lib()
        self.capacity = int(capacity)
        self.key = (int(n), int(n_eq), int(n_in), bool(box), int(hessian), int(backend))
        self.n, self.n_eq, self.n_in, self.box = int(n), int(n_eq), int(n_in), bool(box)
        self.nc = self.n_in + (self.n if self.box else 0)
        self.device = int(device)
        self.handle = self.lib.pqp_batch_create(self.capacity, self.n, self.n_eq, self.n_in, int(self.box), int(hessian), int(backend), int(device))
        if not self.handle:
            msg = _capi.last_error()
            if "wrong argument" in msg:
                raise ValueError(msg)
            raise RuntimeError(f"proxsuite_b200: cannot create device batch: {msg}")
        be = ct.c_int(0)
        self.lib.pqp_batch_dims(self.handle, None, None, None, None, None, ct.cast(ct.pointer(be), _VP))
        self.backend = DenseBackend(be.value)
        self.used = 0
        self.members: List["QP"] = []
        self.generation = 0  # bumped by every solve

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                self.lib.pqp_batch_destroy(h)
            except Exception:
                pass
            self.handle = None

    # -- settings ---------------------------------------------------------
    def push_settings(self, index, settings: Settings):
        _capi.check(self.lib.pqp_batch_settings_set(self.handle, index, ct.byref(settings._c)))

    def pull_settings(self, index, settings: Settings):
        _capi.check(self.lib.pqp_batch_settings_get(self.handle, index, ct.byref(settings._c)))

    def push_all(self):
        for q in self.members:
            self.push_settings(q._index, q.settings)

    def pull_all(self):
        for q in self.members:
            self.pull_settings(q._index, q.settings)

    # -- solve ------------------------------------------------------------
    def solve_async(self, indices=None):
        """`indices`: solve only these members (QP.solve() of one member, solve_in_parallel over a subset); None: all."""
        self.push_all()
        if indices is not None and len(indices) < self.used:
            arr = np.ascontiguousarray(np.asarray(sorted(indices), dtype=np.int64))
            _capi.check(self.lib.pqp_batch_select(self.handle, _ptr(arr), arr.size))
        _capi.check(self.lib.pqp_batch_solve_async(self.handle, None))
        self.generation += 1

    def sync(self):
        _capi.check(self.lib.pqp_batch_sync(self.handle))
        self.pull_all()

    def fetch(self, first, count):
        x = np.empty((count, self.n))
        y = np.empty((count, self.n_eq))
        z = np.empty((count, self.nc))
        se = np.empty((count, self.n_eq))
        si = np.empty((count, self.nc))
        info = (_capi.pqp_info * count)()
        _capi.check(self.lib.pqp_batch_results(self.handle, first, count, _ptr(x), _ptr(y), _ptr(z), _ptr(se), _ptr(si), ct.cast(info, _VP)))
        return x, y, z, se, si, info

    def timings(self):
        a, b, n = ct.c_double(0), ct.c_double(0), ct.c_int64(0)
        self.lib.pqp_batch_timings(self.handle, ct.cast(ct.pointer(a), _VP), ct.cast(ct.pointer(b), _VP), ct.cast(ct.pointer(n), _VP))
        return dict(setup_ms=a.value, solve_ms=b.value, kernel_launches=n.value)


class BackwardData:
    """dense::BackwardData<T> (dense/backward_data.hpp:27-129, expose-model.hpp:25-41)."""

    def __init__(self, dim=0, n_eq=0, n_in=0):
        self.initialize(dim, n_eq, n_in)

    def initialize(self, dim, n_eq, n_in):
        self.dL_dH = np.zeros((dim, dim))
        self.dL_dg = np.zeros(dim)
        self.dL_dA = np.zeros((n_eq, dim))
        self.dL_db = np.zeros(n_eq)
        self.dL_dC = np.zeros((n_in, dim))
        self.dL_du = np.zeros(n_in)
        self.dL_dl = np.zeros(n_in)


class _Model:
    """dense::Model<T> dimensions (dense/model.hpp:23-61) and the backward data it owns (model.hpp:45)."""

    def __init__(self, dim, n_eq, n_in):
        self.dim, self.n_eq, self.n_in = dim, n_eq, n_in
        self.n_total = dim + n_eq + n_in
        self.backward_data = BackwardData(dim, n_eq, n_in)


class QP:
    """proxsuite.proxqp.dense.QP (wrapper.hpp:115-963, expose-qpobject.hpp:26-232)."""

    def __init__(self, n: int, n_eq: int, n_in: int, box_constraints: bool = False,
                 hessian_type=HessianType.Dense, dense_backend=DenseBackend.PrimalDualLDLT,
                 *, device: int = -1, _group: Optional[_Group] = None):
        # the reference has both (box, HessianType, DenseBackend) and
        # (box, DenseBackend, HessianType) overloads (wrapper.hpp:140-333)
        if isinstance(hessian_type, DenseBackend) or isinstance(dense_backend, HessianType):
            hessian_type, dense_backend = (dense_backend if isinstance(dense_backend, HessianType) else HessianType.Dense,
                                           hessian_type if isinstance(hessian_type, DenseBackend) else DenseBackend.PrimalDualLDLT)
        if int(n) == 0:
            raise ValueError("wrong argument size: the dimension wrt the primal variable x should be strictly positive.")
        if _group is None:
            _group = _Group(1, n, n_eq, n_in, box_constraints, hessian_type, dense_backend, device)
        self._group = _group
        self._index = _group.used
        _group.used += 1
        _group.members.append(self)
        self._n, self._n_eq, self._n_in, self._box = int(n), int(n_eq), int(n_in), bool(box_constraints)
        self._nc = self._n_in + (self._n if self._box else 0)
        self.settings = Settings(_group.backend)
        self.model = _Model(self._n, self._n_eq, self._n_in)
        self._results = Results(self._n, self._n_eq, self._n_in, self._box)
        self._results.info.rho = self.settings.default_rho
        self._fetched_generation = 0
        self._host_results_dirty = False

    # -- helpers ------------------------------------------------------------
    def _gather(self, H, g, A, b, Cm, l, u, l_box, u_box):
        n, ne, ni = self._n, self._n_eq, self._n_in
        return [_mat(H, n, n, "H"), _vec(g, n, "g"), _mat(A, ne, n, "A"), _vec(b, ne, "b"), _mat(Cm, ni, n, "C"),
                _vec(l, ni, "l"), _vec(u, ni, "u"), _vec(l_box, n, "l_box"), _vec(u_box, n, "u_box")]

    def _split_args(self, args, kw, pre_name):
        """Reference overloads: (..., l, u, [l_box, u_box,] compute_preconditioner, rho, mu_eq, mu_in, manual_minimal_H_eigenvalue)."""
        args = list(args)
        self._box_args_given = "l_box" in kw or "u_box" in kw
        l_box = kw.pop("l_box", None)
        u_box = kw.pop("u_box", None)
        if len(args) >= 2 and not isinstance(args[0], (bool, np.bool_)) and (args[0] is None or np.ndim(args[0]) >= 1) \
                and not isinstance(args[1], (bool, np.bool_)) and (args[1] is None or np.ndim(args[1]) >= 1):
            l_box, u_box = args[0], args[1]
            args = args[2:]
            self._box_args_given = True
        names = [pre_name, "rho", "mu_eq", "mu_in", "manual_minimal_H_eigenvalue"]
        vals = {pre_name: kw.pop(pre_name, None), "rho": kw.pop("rho", None), "mu_eq": kw.pop("mu_eq", None),
                "mu_in": kw.pop("mu_in", None), "manual_minimal_H_eigenvalue": kw.pop("manual_minimal_H_eigenvalue", None)}
        for name, v in zip(names, args):
            vals[name] = v
        if kw:
            raise TypeError(f"unexpected keyword arguments {sorted(kw)}")
        return l_box, u_box, vals

    def _call(self, fn, arrs, flag, vals):
        keep = [_opt_scalar(vals[k]) for k in ("rho", "mu_eq", "mu_in", "manual_minimal_H_eigenvalue")]
        g = self._group
        g.push_settings(self._index, self.settings)
        rc = fn(g.handle, self._index, 1, *[_ptr(a) for a in arrs], int(flag), *[k[1] for k in keep])
        _capi.check(rc)
        g.pull_settings(self._index, self.settings)
        self._mark_results_stale_from_device()

    def _mark_results_stale_from_device(self):
        self._fetched_generation = -1

    # -- reference API --------------------------------------------------------
    def init(self, H=None, g=None, A=None, b=None, C=None, l=None, u=None, *args, **kw):
        """QP::init (wrapper.hpp:354-498, 520-703)."""
        l_box, u_box, vals = self._split_args(args, kw, "compute_preconditioner")
        if not self._box and (l_box is not None or u_box is not None):
            raise ValueError("wrong model setup: the QP object is designed without box constraints, but is initialized with lower or upper box inequalities.")
        if self._box and not self._box_args_given:  # the 7-argument overload on a box QP (wrapper.hpp:367-372)
            raise ValueError("wrong model setup: the QP object is designed with box constraints, but is initialized without lower or upper box inequalities.")
        arrs = self._gather(H, g, A, b, C, l, u, l_box, u_box)
        pre = True if vals["compute_preconditioner"] is None else bool(vals["compute_preconditioner"])
        self._call(self._group.lib.pqp_batch_init, arrs, pre, vals)

    def update(self, H=None, g=None, A=None, b=None, C=None, l=None, u=None, *args, **kw):
        """QP::update (wrapper.hpp:723-918)."""
        l_box, u_box, vals = self._split_args(args, kw, "update_preconditioner")
        if not self._box and (l_box is not None or u_box is not None):
            raise ValueError("wrong model setup: the QP object is designed without box constraints, but the update includes lower or upper box inequalities.")
        arrs = self._gather(H, g, A, b, C, l, u, l_box, u_box)
        pre = False if vals["update_preconditioner"] is None else bool(vals["update_preconditioner"])
        self._call(self._group.lib.pqp_batch_update, arrs, pre, vals)

    def _warm_start(self, x, y, z):
        if x is None and y is None and z is None:
            return
        xs = [_vec(x, self._n, "x", False), _vec(y, self._n_eq, "y", False), _vec(z, self._nc, "z", False)]
        g = self._group
        g.push_settings(self._index, self.settings)
        _capi.check(g.lib.pqp_batch_warm_start(g.handle, self._index, 1, *[_ptr(a) for a in xs]))
        g.pull_settings(self._index, self.settings)

    def solve(self, x=None, y=None, z=None):
        """QP::solve() / solve(x, y, z) (wrapper.hpp:922-954)."""
        self._warm_start(x, y, z)
        g = self._group
        g.solve_async([self._index])  # this QP only: siblings of a shared device batch keep their results
        g.sync()

    def cleanup(self):
        """QP::cleanup (wrapper.hpp:958-962)."""
        g = self._group
        g.push_settings(self._index, self.settings)
        _capi.check(g.lib.pqp_batch_cleanup(g.handle, self._index, 1))
        self._mark_results_stale_from_device()

    @property
    def results(self) -> Results:
        if self._fetched_generation != self._group.generation:
            x, y, z, se, si, info = self._group.fetch(self._index, 1)
            r = self._results
            r.x, r.y, r.z, r.se, r.si = x[0], y[0], z[0], se[0], si[0]
            r.info = Info(info[0])
            self._fetched_generation = self._group.generation
        return self._results

    def scaled(self):
        """qp.work scaled data and qp.ruiz (delta, c); test/debug helper."""
        n, ne, ni, nc = self._n, self._n_eq, self._n_in, self._nc
        H = np.zeros((n, n)); g = np.zeros(n); A = np.zeros((ne, n)); b = np.zeros(ne)
        Cm = np.zeros((ni, n)); u = np.zeros(nc); l = np.zeros(nc); delta = np.zeros(n + ne + nc)
        c = ct.c_double(0)
        _capi.check(self._group.lib.pqp_batch_scaled(self._group.handle, self._index, _ptr(H), _ptr(g), _ptr(A), _ptr(b), _ptr(Cm), _ptr(u), _ptr(l), _ptr(delta), ct.cast(ct.pointer(c), _VP)))
        return dict(H=H, g=g, A=A, b=b, C=Cm, u=u, l=l, delta=delta, c=c.value)


class BatchQP:
    """proxsuite.proxqp.dense.BatchQP (wrapper.hpp:1253-1311, expose-qpvector.hpp:19-39).

    `init_qp_in_place` accepts the QP constructor's optional arguments
    (box_constraints, hessian_type, dense_backend) that the reference's
    `init_qp_in_place(dim, n_eq, n_in)` cannot express."""

    def __init__(self, batch_size: int = 0, device: int = -1):
        self._capacity = int(batch_size)
        self._device = device
        self._groups = {}
        self._count = {}
        self._qps: List[QP] = []

    def _group_for(self, key):
        g = self._groups.get(key)
        if g is None or g.used >= g.capacity:
            # first QP of this shape, or the reserved capacity is exhausted:
            # open a new device batch (earlier QPs keep their own)
            # geometric growth (a BatchQP() built without a size, as the reference's QP layer does, would otherwise
            # open one capacity-1 device batch - streams, events, ~35 allocations, one launch - per QP)
            total = self._count.get(key, 0)
            n, ne, ni, box, ht, be = key
            dflt = int(max(1, min(64, (256 << 20) // max(1, 16 * (n * n + (ne + ni) * n)))))  # <= 256 MB of device copies
            left = self._capacity - len(self._qps)
            cap = (left if left > 0 else dflt) if g is None else max(2 * total, dflt)
            g = _Group(cap, n, ne, ni, box, ht, be, self._device)
            self._groups[key] = g
        return g

    def init_qp_in_place(self, dim: int, n_eq: int, n_in: int, box_constraints: bool = False,
                         hessian_type=HessianType.Dense, dense_backend=DenseBackend.PrimalDualLDLT) -> QP:
        if int(dim) == 0:
            raise ValueError("wrong argument size: the dimension wrt the primal variable x should be strictly positive.")
        key = (int(dim), int(n_eq), int(n_in), bool(box_constraints), int(hessian_type), int(dense_backend))
        g = self._group_for(key)
        self._count[key] = self._count.get(key, 0) + 1
        qp = QP(dim, n_eq, n_in, box_constraints, HessianType(int(hessian_type)), DenseBackend(int(dense_backend)), _group=g)
        self._qps.append(qp)
        return qp

    def insert(self, qp: QP):
        """BatchQP::insert (wrapper.hpp:1288); unlike the reference (quirk 6 of
        SURVEY Appendix A) the inserted QP is counted and solved."""
        self._qps.append(qp)

    def get(self, i: int) -> QP:
        return self._qps[i]

    def __getitem__(self, i: int) -> QP:
        return self._qps[i]

    def size(self) -> int:
        return len(self._qps)

    def __len__(self):
        return len(self._qps)

    def __iter__(self):
        return iter(self._qps)


class VectorQP(list):
    """std::vector<dense::QP<double>> (expose-qpvector.hpp:13, 30-31)."""

    def init_qp(self, dim, n_eq, n_in):
        qp = QP(dim, n_eq, n_in)
        self.append(qp)
        return qp


def _groups_of(qps: Iterable[QP]):
    seen, out = set(), []
    for q in qps:
        g = q._group
        if id(g) not in seen:
            seen.add(id(g))
            out.append(g)
    return out


def solve_in_parallel(qps, num_threads: Optional[int] = None):
    """dense::solve_in_parallel (parallel/qp_solve.hpp:17-60, expose-parallel.hpp:33-46).

    `num_threads` is accepted for signature parity and ignored: the work
    distribution is the kernel's atomic work queue over persistent CTAs."""
    qps = list(qps)
    groups = _groups_of(qps)
    members = {}
    for q in qps:
        members.setdefault(id(q._group), []).append(q._index)
    for g in groups:
        g.solve_async(members[id(g)])  # only the listed QPs of each device batch
    for g in groups:
        g.sync()


class VectorLossDerivatives(list):
    """std::vector<dense::Vec<double>> (expose-qpvector.hpp:34-37)."""


def _group_backward(group, first, count, loss, eps, rho_backward, mu_backward):
    n, ne, ni = group.n, group.n_eq, group.n_in
    loss = np.ascontiguousarray(np.asarray(loss, dtype=np.float64).reshape(count, n + ne + ni))
    out = dict(dL_dH=np.zeros((count, n, n)), dL_dg=np.zeros((count, n)), dL_dA=np.zeros((count, ne, n)), dL_db=np.zeros((count, ne)),
               dL_dC=np.zeros((count, ni, n)), dL_du=np.zeros((count, ni)), dL_dl=np.zeros((count, ni)))
    _capi.check(group.lib.pqp_batch_backward(group.handle, int(first), int(count), _ptr(loss), float(eps), float(rho_backward), float(mu_backward),
                                             *[_ptr(out[k]) for k in ("dL_dH", "dL_dg", "dL_dA", "dL_db", "dL_dC", "dL_du", "dL_dl")]))
    return out


def _store_backward(qp, out, k):
    bd = qp.model.backward_data
    for name in ("dL_dH", "dL_dg", "dL_dA", "dL_db", "dL_dC", "dL_du", "dL_dl"):
        setattr(bd, name, out[name][k].copy())
    qp._results.info.rho = float(qp._bw_rho)
    qp._results.info.mu_eq = qp._results.info.mu_in = float(qp._bw_mu)


def compute_backward(qp: QP, loss_derivative, eps: float = 1e-4, rho_backward: float = 1e-6, mu_backward: float = 1e-6):
    """dense::compute_backward (dense/compute_ECJ.hpp:29-125, expose-backward.hpp:22-30): fills
    qp.model.backward_data with the derivatives of the loss w.r.t. H, g, A, b, C, u, l."""
    ld = np.asarray(loss_derivative, dtype=np.float64)
    if ld.size != qp.model.n_total:
        raise ValueError("wrong argument size: loss_derivative must have dim + n_eq + n_in entries")
    qp.results  # pull the latest results / status before the info fields are touched
    out = _group_backward(qp._group, qp._index, 1, ld, eps, rho_backward, mu_backward)
    qp._bw_rho, qp._bw_mu = rho_backward, mu_backward
    _store_backward(qp, out, 0)


def solve_backward_in_parallel(num_threads=None, qps=None, loss_derivatives=None, eps: float = 1e-4, rho_backward: float = 1e-6,
                               mu_backward: float = 1e-6):
    """dense::solve_backward_in_parallel (parallel/qp_solve.hpp:84-138, expose-parallel.hpp:55-81): the backward pass
    of every QP of a BatchQP / VectorQP; one kernel launch per run of QPs that are contiguous in a device batch.
    `num_threads` is accepted for signature parity and ignored."""
    qps = list(qps)
    if len(qps) != len(loss_derivatives):
        raise ValueError("wrong argument size: one loss derivative per QP is required")
    k = 0
    while k < len(qps):
        g, first = qps[k]._group, qps[k]._index
        e = k
        while e + 1 < len(qps) and qps[e + 1]._group is g and qps[e + 1]._index == qps[e]._index + 1:
            e += 1
        for q in qps[k:e + 1]:
            q.results
        loss = np.stack([np.asarray(loss_derivatives[j], dtype=np.float64) for j in range(k, e + 1)])
        out = _group_backward(g, first, e - k + 1, loss, eps, rho_backward, mu_backward)
        for j in range(k, e + 1):
            qps[j]._bw_rho, qps[j]._bw_mu = rho_backward, mu_backward
            _store_backward(qps[j], out, j - k)
        k = e + 1


_SOLVE_TAIL = ("x", "y", "z", "eps_abs", "eps_rel", "rho", "mu_eq", "mu_in", "verbose", "compute_preconditioner", "compute_timings", "max_iter",
               "initial_guess", "check_duality_gap", "eps_duality_gap_abs", "eps_duality_gap_rel", "primal_infeasibility_solving",
               "default_H_eigenvalue_estimate")
_SOLVE_PLAIN = ("H", "g", "A", "b", "C", "l", "u") + _SOLVE_TAIL  # expose-solve.hpp:54-79
_SOLVE_BOX = ("H", "g", "A", "b", "C", "l", "u", "l_box", "u_box") + _SOLVE_TAIL  # expose-solve.hpp:115-142


def solve(*args, **kwargs) -> Results:
    """Free function dense::solve (wrapper.hpp:1000-1233), both overloads of expose-solve.hpp:20-142 with their positional
    argument orders: (H, g, A, b, C, l, u, x, y, z, eps_abs, ...) and (H, g, A, b, C, l, u, l_box, u_box, x, y, z, eps_abs, ...).
    Overload resolution as the reference's binding does it: the plain overload is tried first and is rejected when the
    argument at its `eps_abs` position (index 10, `y` of the box overload) is an array, or when l_box / u_box are given
    by keyword."""
    box_call = "l_box" in kwargs or "u_box" in kwargs or (len(args) > 10 and args[10] is not None and np.ndim(args[10]) >= 1)
    names = _SOLVE_BOX if box_call else _SOLVE_PLAIN
    if len(args) > len(names):
        raise TypeError(f"solve() takes at most {len(names)} positional arguments ({len(args)} given)")
    kw = dict(zip(names, args))
    for k, v in kwargs.items():
        if k not in _SOLVE_BOX:
            raise TypeError(f"solve() got an unexpected keyword argument '{k}'")
        if k in kw:
            raise TypeError(f"solve() got multiple values for argument '{k}'")
        kw[k] = v
    return _solve(**kw)


def _solve(H=None, g=None, A=None, b=None, C=None, l=None, u=None, x=None, y=None, z=None, eps_abs=None, eps_rel=None,
           rho=None, mu_eq=None, mu_in=None, verbose=None, compute_preconditioner=True, compute_timings=False,
           max_iter=None, initial_guess=InitialGuess.EQUALITY_CONSTRAINED_INITIAL_GUESS, check_duality_gap=False,
           eps_duality_gap_abs=None, eps_duality_gap_rel=None, primal_infeasibility_solving=False,
           default_H_eigenvalue_estimate=0.0, l_box=None, u_box=None) -> Results:
    n = 0
    if H is not None:
        n = np.asarray(H).shape[0]
    elif g is not None:
        n = np.asarray(g).shape[0]
    elif A is not None:
        n = np.asarray(A).shape[1]
    elif C is not None:
        n = np.asarray(C).shape[1]
    n_eq = 0 if A is None else np.asarray(A).shape[0]
    n_in = 0 if C is None else np.asarray(C).shape[0]
    box = l_box is not None or u_box is not None
    qp = QP(n, n_eq, n_in, box)
    qp.settings.initial_guess = initial_guess
    qp.settings.check_duality_gap = check_duality_gap
    if eps_abs is not None:
        qp.settings.eps_abs = eps_abs
    if eps_rel is not None:
        qp.settings.eps_rel = eps_rel
    if verbose is not None:
        qp.settings.verbose = verbose
    if max_iter is not None:
        qp.settings.max_iter = max_iter
    if eps_duality_gap_abs is not None:
        qp.settings.eps_duality_gap_abs = eps_duality_gap_abs
    if eps_duality_gap_rel is not None:
        qp.settings.eps_duality_gap_rel = eps_duality_gap_rel
    qp.settings.compute_timings = compute_timings
    qp.settings.primal_infeasibility_solving = primal_infeasibility_solving
    kw = dict(rho=rho, mu_eq=mu_eq, mu_in=mu_in, manual_minimal_H_eigenvalue=default_H_eigenvalue_estimate or None)
    if box:
        qp.init(H, g, A, b, C, l, u, l_box, u_box, compute_preconditioner, **kw)
    else:
        qp.init(H, g, A, b, C, l, u, compute_preconditioner, **kw)
    qp.solve(x, y, z)
    return qp.results


# ---------------------------------------------------------------------------
# Bulk (array-of-QPs) API: one H2D copy + one set-up launch + one solve launch.
# This is what bench.py times end to end.
# ---------------------------------------------------------------------------
class DenseBatch:
    """A batch of B same-shaped QPs given as stacked arrays
    (H[B,n,n], g[B,n], A[B,n_eq,n], ...). Semantically B QP objects sharing one
    Settings; the per-QP object API above is built on the same C-ABI."""

    def __init__(self, batch, n, n_eq, n_in, box_constraints=False, hessian_type=HessianType.Dense,
                 dense_backend=DenseBackend.PrimalDualLDLT, device=-1):
        self._g = _Group(batch, n, n_eq, n_in, box_constraints, hessian_type, dense_backend, device)
        self.batch = int(batch)
        self.settings = Settings(self._g.backend)

    def _push(self):
        _capi.check(self._g.lib.pqp_batch_settings_set(self._g.handle, -1, ct.byref(self.settings._c)))

    def init(self, H=None, g=None, A=None, b=None, C=None, l=None, u=None, l_box=None, u_box=None,
             compute_preconditioner=True, rho=None, mu_eq=None, mu_in=None, manual_minimal_H_eigenvalue=None, update=False):
        G = self._g
        B, n, ne, ni = self.batch, G.n, G.n_eq, G.n_in

        def chk(a, shape, what):
            if a is None:
                return None
            a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
            if a.size == 0:
                return None
            if a.size != int(np.prod(shape)):
                raise ValueError(f"wrong argument size: expected {shape} for {what}, got {a.shape}")
            return a
        arrs = [chk(H, (B, n, n), "H"), chk(g, (B, n), "g"), chk(A, (B, ne, n), "A"), chk(b, (B, ne), "b"),
                chk(C, (B, ni, n), "C"), chk(l, (B, ni), "l"), chk(u, (B, ni), "u"), chk(l_box, (B, n), "l_box"),
                chk(u_box, (B, n), "u_box")]
        keep = [_opt_scalar(v) for v in (rho, mu_eq, mu_in, manual_minimal_H_eigenvalue)]
        self._push()
        self._inputs_in_flight = arrs  # pinned inputs are uploaded asynchronously: keep them alive until the next call
        fn = G.lib.pqp_batch_update if update else G.lib.pqp_batch_init
        _capi.check(fn(G.handle, 0, B, *[_ptr(a) for a in arrs], int(compute_preconditioner), *[k[1] for k in keep]))
        _capi.check(G.lib.pqp_batch_settings_get(G.handle, 0, ct.byref(self.settings._c)))

    def update(self, **kw):
        kw.setdefault("compute_preconditioner", kw.pop("update_preconditioner", False))
        self.init(update=True, **kw)

    def warm_start(self, x=None, y=None, z=None):
        G = self._g
        xs = [None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float64)) for a in (x, y, z)]
        self._push()
        _capi.check(G.lib.pqp_batch_warm_start(G.handle, 0, self.batch, *[_ptr(a) for a in xs]))
        _capi.check(G.lib.pqp_batch_settings_get(G.handle, 0, ct.byref(self.settings._c)))

    def solve(self, x=None, y=None, z=None):
        if x is not None or y is not None or z is not None:
            self.warm_start(x, y, z)
        self._push()
        _capi.check(self._g.lib.pqp_batch_solve(self._g.handle))
        _capi.check(self._g.lib.pqp_batch_settings_get(self._g.handle, 0, ct.byref(self.settings._c)))

    def solve_async(self, stream=None):
        self._push()
        _capi.check(self._g.lib.pqp_batch_solve_async(self._g.handle, stream))

    def sync(self):
        _capi.check(self._g.lib.pqp_batch_sync(self._g.handle))

    def results(self):
        x, y, z, se, si, info = self._g.fetch(0, self.batch)
        # one structured view over the C array instead of batch x fields attribute reads
        rec = np.frombuffer(info, dtype=_capi.INFO_DTYPE, count=self.batch)
        inf = {k: np.ascontiguousarray(rec[k]) for k in rec.dtype.names}
        return dict(x=x, y=y, z=z, se=se, si=si, info=inf)

    def results_device(self):
        """Zero-copy torch views of the solutions where the solve kernel wrote them (device memory of this batch):
        dict(x[B, n], y[B, n_eq], z[B, n_cons], info[B, 20]); valid after sync() and until the next solve / init.
        info columns follow pqp_info (6 iter, 7 iter_ext, 8 mu_updates, 10 status, 14 objValue, 15 pri_res, 16 dua_res).
        Lets a caller hand the results to NCCL (sharding.solve_sharded) or to torch code without a host round trip."""
        import torch

        G = self._g
        ptrs = [ct.c_void_p(0) for _ in range(4)]
        _capi.check(G.lib.pqp_batch_results_device(G.handle, *[ct.byref(q) for q in ptrs]))
        nc = G.n_in + (G.n if G.box else 0)

        class _View:
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = dict(shape=shape, typestr="<f8", data=(int(ptr), False), version=3, strides=None)

        dev = torch.device("cuda", G.device if G.device >= 0 else torch.cuda.current_device())
        out = {}
        for name, q, w in (("x", ptrs[0], G.n), ("y", ptrs[1], G.n_eq), ("z", ptrs[2], nc), ("info", ptrs[3], 20)):
            if w == 0 or not q.value:
                out[name] = torch.empty((self.batch, 0), dtype=torch.float64, device=dev)
            else:
                out[name] = torch.as_tensor(_View(q.value, (self.batch, w)), device=dev)
        return out

    def timings(self):
        return self._g.timings()

    def backward(self, loss_derivatives, eps=1e-4, rho_backward=1e-6, mu_backward=1e-6):
        """solve_backward_in_parallel over the whole batch: loss_derivatives[B, n + n_eq + n_in] -> dict of stacked
        BackwardData arrays (dL_dH[B, n, n], dL_dg, dL_dA, dL_db, dL_dC, dL_du, dL_dl)."""
        return _group_backward(self._g, 0, self.batch, loss_derivatives, eps, rho_backward, mu_backward)

    def scaled(self, index):
        """qp.work scaled data and qp.ruiz (delta, c) of QP `index`; test/debug helper."""
        G = self._g
        n, ne, ni = G.n, G.n_eq, G.n_in
        nc = ni + (n if G.box else 0)
        H = np.zeros((n, n)); g = np.zeros(n); A = np.zeros((ne, n)); b = np.zeros(ne)
        Cm = np.zeros((ni, n)); u = np.zeros(nc); l = np.zeros(nc); delta = np.zeros(n + ne + nc)
        c = ct.c_double(0)
        _capi.check(G.lib.pqp_batch_scaled(G.handle, int(index), _ptr(H), _ptr(g), _ptr(A), _ptr(b), _ptr(Cm), _ptr(u), _ptr(l), _ptr(delta), ct.cast(ct.pointer(c), _VP)))
        return dict(H=H, g=g, A=A, b=b, C=Cm, u=u, l=l, delta=delta, c=c.value)

    def launch_config(self):
        grid, smem, mask, ws = ct.c_int(0), ct.c_int(0), ct.c_int(0), ct.c_int64(0)
        G = self._g
        G.lib.pqp_batch_launch_config(G.handle, ct.cast(ct.pointer(grid), _VP), ct.cast(ct.pointer(smem), _VP), ct.cast(ct.pointer(mask), _VP), ct.cast(ct.pointer(ws), _VP))
        return dict(grid=grid.value, smem_bytes=smem.value, in_smem_mask=mask.value, ws_doubles=ws.value & 0xffffffff,
                    si_cap=(ws.value >> 32) & 0xffff, overflow_retries=ws.value >> 48)

    PROFILE_PHASES = ["stage", "build_M1", "eq_block", "insert", "delete", "solve_kkt", "kkt_residual", "linesearch",
                      "mu_update", "global_residuals", "newton_misc", "total"]

    def occupancy(self, fused=False):
        """resident CTAs per SM the runtime grants the solve kernel of this batch's layout (plain / fused instantiation)"""
        return int(self._g.lib.pqp_batch_occupancy(self._g.handle, int(bool(fused))))

    def profile(self, reset=True):
        """Per-phase SM cycles summed over all QPs solved so far (PQP_PROFILE=1)."""
        out = (ct.c_longlong * 12)()
        k = self._g.lib.pqp_batch_profile(self._g.handle, out, int(reset))
        return {n: int(out[i]) for i, n in enumerate(self.PROFILE_PHASES)} if k else None

    def debug_trace(self):
        out = np.zeros(6 * 4096)
        k = self._g.lib.pqp_batch_debug_trace(self._g.handle, _ptr(out), out.size)
        return out[:k].reshape(-1, 6)


class ShardedBatch:
    """One batch of B same-shaped QPs sharded over several GPUs of this node from ONE process (pqp_sharded_* of the
    C-ABI): contiguous slices, one device batch each, no cross-device dependency inside the iteration
    (parallel/qp_solve.hpp:55-59). `devices`: CUDA ordinals (default: every visible device). The torch.distributed
    variant (one process per GPU, NCCL gather of the solutions) is proxsuite_b200.sharding."""

    def __init__(self, batch, n, n_eq, n_in, box_constraints=False, hessian_type=HessianType.Dense,
                 dense_backend=DenseBackend.PrimalDualLDLT, devices=None):
        self.lib = _capi.lib()
        if devices is None:
            import torch

            devices = list(range(max(1, torch.cuda.device_count())))
        self.devices = [int(d) for d in devices]
        arr = (ct.c_int * len(self.devices))(*self.devices)
        self.batch, self.n, self.n_eq, self.n_in, self.box = int(batch), int(n), int(n_eq), int(n_in), bool(box_constraints)
        self.nc = self.n_in + (self.n if self.box else 0)
        self.handle = self.lib.pqp_sharded_create(self.batch, self.n, self.n_eq, self.n_in, int(self.box), int(hessian_type), int(dense_backend), arr, len(self.devices))
        if not self.handle:
            msg = _capi.last_error()
            raise (ValueError if "wrong argument" in msg else RuntimeError)(f"proxsuite_b200: cannot create sharded batch: {msg}")
        self.settings = Settings(DenseBackend(self.lib.pqp_dense_backend_choice(int(dense_backend), self.n, self.n_eq, self.n_in, int(self.box))))

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            self.lib.pqp_sharded_destroy(h)
            self.handle = None

    def shards(self):
        """[(device, first, count)] of the slices"""
        out = []
        for k in range(self.lib.pqp_sharded_count(self.handle)):
            f, c = ct.c_int64(0), ct.c_int64(0)
            self.lib.pqp_sharded_shard(self.handle, k, ct.byref(f), ct.byref(c))
            out.append((self.devices[k], f.value, c.value))
        return out

    def _feed(self, fn, flag, H, g, A, b, C, l, u, l_box, u_box, rho, mu_eq, mu_in, eig):
        B, n, ne, ni = self.batch, self.n, self.n_eq, self.n_in

        def chk(a, shape, what):
            if a is None:
                return None
            a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
            if a.size == 0:
                return None
            if a.size != int(np.prod(shape)):
                raise ValueError(f"wrong argument size: expected {shape} for {what}, got {a.shape}")
            return a
        arrs = [chk(H, (B, n, n), "H"), chk(g, (B, n), "g"), chk(A, (B, ne, n), "A"), chk(b, (B, ne), "b"), chk(C, (B, ni, n), "C"),
                chk(l, (B, ni), "l"), chk(u, (B, ni), "u"), chk(l_box, (B, n), "l_box"), chk(u_box, (B, n), "u_box")]
        keep = [_opt_scalar(v) for v in (rho, mu_eq, mu_in, eig)]
        _capi.check(self.lib.pqp_sharded_settings_set(self.handle, ct.byref(self.settings._c)))
        _capi.check(fn(self.handle, *[_ptr(a) for a in arrs], int(flag), *[k[1] for k in keep]))

    def init(self, H=None, g=None, A=None, b=None, C=None, l=None, u=None, l_box=None, u_box=None, compute_preconditioner=True,
             rho=None, mu_eq=None, mu_in=None, manual_minimal_H_eigenvalue=None):
        self._feed(self.lib.pqp_sharded_init, compute_preconditioner, H, g, A, b, C, l, u, l_box, u_box, rho, mu_eq, mu_in, manual_minimal_H_eigenvalue)

    def update(self, H=None, g=None, A=None, b=None, C=None, l=None, u=None, l_box=None, u_box=None, update_preconditioner=False,
               rho=None, mu_eq=None, mu_in=None, manual_minimal_H_eigenvalue=None):
        self._feed(self.lib.pqp_sharded_update, update_preconditioner, H, g, A, b, C, l, u, l_box, u_box, rho, mu_eq, mu_in, manual_minimal_H_eigenvalue)

    def solve(self):
        _capi.check(self.lib.pqp_sharded_settings_set(self.handle, ct.byref(self.settings._c)))
        _capi.check(self.lib.pqp_sharded_solve(self.handle))

    def results(self):
        B = self.batch
        x, y, z = np.zeros((B, self.n)), np.zeros((B, self.n_eq)), np.zeros((B, self.nc))
        se, si = np.zeros((B, self.n_eq)), np.zeros((B, self.nc))
        info = (_capi.pqp_info * B)()
        _capi.check(self.lib.pqp_sharded_results(self.handle, _ptr(x), _ptr(y), _ptr(z), _ptr(se), _ptr(si), info))
        rec = np.frombuffer(info, dtype=_capi.INFO_DTYPE, count=B)
        return dict(x=x, y=y, z=z, se=se, si=si, info={k: np.ascontiguousarray(rec[k]) for k in rec.dtype.names})


_GEN_KINDS = {"strongly_convex": 0, "not_strongly_convex": 1, "degenerate": 2, "box_constrained": 3,
              "box_benchmark": 4, "diagonal_benchmark": 5}


def random_qp(kind: str, seed: int, n: int, n_eq: int, n_in: int, sparsity_factor: float = 0.15,
              strong_convexity_factor: float = 1e-2):
    """Reference-specified synthetic QP (utils/random_qp_problems.hpp:463-628 and the
    box / diagonal benchmark recipes); host code, deterministic Lehmer-64 stream."""
    k = _GEN_KINDS[kind]
    rows = 2 * n_in if k == 2 else n_in
    H = np.zeros((n, n)); g = np.zeros(n); A = np.zeros((n_eq, n)); b = np.zeros(n_eq)
    Cm = np.zeros((rows, n)); u = np.zeros(rows); l = np.zeros(rows); ub = np.zeros(n); lb = np.zeros(n)
    rc = _capi.lib().pqp_random_qp(k, int(seed), n, n_eq, n_in, float(sparsity_factor), float(strong_convexity_factor),
                                   _ptr(H), _ptr(g), _ptr(A), _ptr(b), _ptr(Cm), _ptr(u), _ptr(l), _ptr(ub), _ptr(lb))
    _capi.check(rc)
    out = dict(H=H, g=g, A=A, b=b, C=Cm, u=u, l=l)
    if k >= 4:
        out.update(u_box=ub, l_box=lb)
    return out
