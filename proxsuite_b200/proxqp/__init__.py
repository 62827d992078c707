"""`proxsuite_b200.proxqp` — mirrors `proxsuite.proxqp` for the dense batch path.

Names follow the reference bindings:
  Settings / Results / Info      bindings/python/src/expose-settings.hpp:22-103,
                                 expose-results.hpp:23-113
  QPSolverOutput, InitialGuess, MeritFunctionType, DenseBackend, HessianType
                                 status.hpp:17-35, settings.hpp:26-45
  dense.{QP, BatchQP, VectorQP, solve, solve_in_parallel}
                                 expose-qpobject.hpp:26-232, expose-qpvector.hpp:19-39,
                                 expose-parallel.hpp:24-83, expose-solve.hpp:20
"""
from __future__ import annotations

import enum

from .. import _capi


class QPSolverOutput(enum.IntEnum):  # status.hpp:17-26
    PROXQP_SOLVED = 0
    PROXQP_MAX_ITER_REACHED = 1
    PROXQP_PRIMAL_INFEASIBLE = 2
    PROXQP_SOLVED_CLOSEST_PRIMAL_FEASIBLE = 3
    PROXQP_DUAL_INFEASIBLE = 4
    PROXQP_NOT_RUN = 5


class InitialGuess(enum.IntEnum):  # status.hpp:28-35
    NO_INITIAL_GUESS = 0
    EQUALITY_CONSTRAINED_INITIAL_GUESS = 1
    WARM_START_WITH_PREVIOUS_RESULT = 2
    WARM_START = 3
    COLD_START_WITH_PREVIOUS_RESULT = 4


class DenseBackend(enum.IntEnum):  # settings.hpp:33-39
    Automatic = 0
    PrimalDualLDLT = 1
    PrimalLDLT = 2


class MeritFunctionType(enum.IntEnum):  # settings.hpp:41-45
    GPDAL = 0
    PDAL = 1


class HessianType(enum.IntEnum):  # settings.hpp:47-52
    Zero = 0
    Dense = 1
    Diagonal = 2


_BOOL_FIELDS = {"verbose", "update_preconditioner", "compute_preconditioner", "compute_timings",
                "check_duality_gap", "bcl_update", "primal_infeasibility_solving"}
_ENUM_FIELDS = {"initial_guess": InitialGuess, "merit_function_type": MeritFunctionType}


class Settings:
    """Settings<double> (settings.hpp:88-316). Attribute names are the reference's."""

    def __init__(self, dense_backend: DenseBackend = DenseBackend.PrimalDualLDLT):
        object.__setattr__(self, "_c", _capi.pqp_settings())
        _capi.lib().pqp_settings_default(self._c, int(dense_backend))

    def __getattr__(self, name):
        c = object.__getattribute__(self, "_c")
        if name.startswith("_") or not hasattr(c, name):
            raise AttributeError(name)
        v = getattr(c, name)
        if name in _BOOL_FIELDS:
            return bool(v)
        if name in _ENUM_FIELDS:
            return _ENUM_FIELDS[name](v)
        return v

    def __setattr__(self, name, value):
        c = object.__getattribute__(self, "_c")
        if not hasattr(c, name) or name == "reserved_":
            raise AttributeError(f"Settings has no attribute {name!r}")
        if name in _BOOL_FIELDS or name in _ENUM_FIELDS:
            value = int(value)
        setattr(c, name, value)

    def _copy_from(self, other: "Settings"):
        import ctypes
        ctypes.memmove(ctypes.byref(self._c), ctypes.byref(other._c), ctypes.sizeof(_capi.pqp_settings))

    def __eq__(self, other):
        return isinstance(other, Settings) and bytes(self._c) == bytes(other._c)

    def __repr__(self):
        return "Settings(" + ", ".join(f"{k}={getattr(self, k)!r}" for k, _ in _capi.pqp_settings._fields_ if k != "reserved_") + ")"


class Info:
    """Info<double> (results.hpp:28-58)."""

    _FIELDS = [k for k, _ in _capi.pqp_info._fields_]

    def __init__(self, c_info=None):
        c = c_info if c_info is not None else _capi.pqp_info()
        for k in self._FIELDS:
            v = getattr(c, k)
            if k == "status":
                v = QPSolverOutput(int(v))
            setattr(self, k, v)
        if c_info is None:
            self.status = QPSolverOutput.PROXQP_NOT_RUN
            self.mu_eq, self.mu_eq_inv, self.mu_in, self.mu_in_inv, self.rho, self.nu = 1e-3, 1e3, 1e-1, 1e1, 1e-6, 1.0

    def __repr__(self):
        return "Info(" + ", ".join(f"{k}={getattr(self, k)!r}" for k in self._FIELDS) + ")"


class Results:
    """Results<double> (results.hpp:67-203): x, y, z, se, si, info."""

    def __init__(self, n=0, n_eq=0, n_in=0, box_constraints=False):
        import numpy as np
        nc = n_in + (n if box_constraints else 0)
        self.x = np.zeros(n)
        self.y = np.zeros(n_eq)
        self.z = np.zeros(nc)
        self.se = np.zeros(n_eq)
        self.si = np.zeros(nc)
        self.info = Info()


def omp_get_max_threads() -> int:
    """Kept for API parity (expose-all.cpp:121): the GPU path has no OpenMP
    team; returns the number of persistent CTAs a solve would launch on the
    current device, or 0 when no device is present."""
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    except Exception:
        pass
    return 0


from . import dense  # noqa: E402,F401
