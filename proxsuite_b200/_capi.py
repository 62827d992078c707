"""ctypes loader of the C-ABI shared library (include/pqp.h).

The library is built in-tree (proxsuite_b200/libpqp_b200.so) by
`__graft_entry__.build()` / `make -C proxsuite_b200/csrc`.  There is no Python
or CPU fallback: if the library is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PQP_B200_LIB") or os.path.join(_HERE, "libpqp_b200.so")

PQP_OK, PQP_EINVAL, PQP_ECUDA, PQP_ESTATE = 0, -1, -2, -3


class pqp_settings(C.Structure):
    """Field-for-field mirror of `pqp_settings` (Settings<double>, settings.hpp:88-210)."""

    _fields_ = [(k, C.c_double) for k in (
        "default_rho", "default_mu_eq", "default_mu_in", "alpha_bcl", "beta_bcl",
        "refactor_dual_feasibility_threshold", "refactor_rho_threshold", "mu_min_eq", "mu_min_in",
        "mu_max_eq_inv", "mu_max_in_inv", "mu_update_factor", "mu_update_inv_factor",
        "cold_reset_mu_eq", "cold_reset_mu_in", "cold_reset_mu_eq_inv", "cold_reset_mu_in_inv",
        "eps_abs", "eps_rel", "eps_refact", "eps_duality_gap_abs", "eps_duality_gap_rel",
        "preconditioner_accuracy", "eps_primal_inf", "eps_dual_inf", "alpha_gpdal",
        "default_H_eigenvalue_estimate")] + [(k, C.c_int64) for k in (
        "max_iter", "max_iter_in", "safe_guard", "nb_iterative_refinement", "preconditioner_max_iter",
        "frequence_infeasibility_check")] + [(k, C.c_int32) for k in (
        "verbose", "initial_guess", "update_preconditioner", "compute_preconditioner", "compute_timings",
        "check_duality_gap", "bcl_update", "merit_function_type", "primal_infeasibility_solving", "reserved_")]


class pqp_info(C.Structure):
    """Mirror of `pqp_info` (Info<double>, results.hpp:28-58)."""

    _fields_ = [("mu_eq", C.c_double), ("mu_eq_inv", C.c_double), ("mu_in", C.c_double), ("mu_in_inv", C.c_double),
                ("rho", C.c_double), ("nu", C.c_double), ("iter", C.c_int64), ("iter_ext", C.c_int64),
                ("mu_updates", C.c_int64), ("rho_updates", C.c_int64), ("status", C.c_int64),
                ("setup_time", C.c_double), ("solve_time", C.c_double), ("run_time", C.c_double),
                ("objValue", C.c_double), ("pri_res", C.c_double), ("dua_res", C.c_double),
                ("duality_gap", C.c_double), ("iterative_residual", C.c_double),
                ("minimal_H_eigenvalue_estimate", C.c_double)]


import numpy as _np  # noqa: E402

# numpy view of an array of pqp_info (all members are 8 bytes wide, no padding)
INFO_DTYPE = _np.dtype([(k, _np.int64 if t is C.c_int64 else _np.float64) for k, t in pqp_info._fields_])
assert INFO_DTYPE.itemsize == C.sizeof(pqp_info)

# every symbol include/pqp.h declares
EXPORTED_SYMBOLS = [
    "pqp_settings_default", "pqp_dense_backend_choice", "pqp_batch_create", "pqp_batch_destroy", "pqp_batch_size",
    "pqp_batch_dims", "pqp_batch_settings_get", "pqp_batch_settings_set", "pqp_batch_init", "pqp_batch_init_device",
    "pqp_batch_update", "pqp_batch_warm_start", "pqp_batch_solve", "pqp_batch_solve_async", "pqp_batch_sync", "pqp_batch_select",
    "pqp_batch_results", "pqp_batch_results_device", "pqp_batch_scaled", "pqp_batch_backward", "pqp_batch_backward_device", "pqp_batch_results_copy_device", "pqp_batch_cleanup", "pqp_batch_timings",
    "pqp_random_qp", "pqp_last_error", "pqp_version",
    "pqp_sharded_create", "pqp_sharded_destroy", "pqp_sharded_count", "pqp_sharded_shard", "pqp_sharded_settings_set",
    "pqp_sharded_init", "pqp_sharded_update", "pqp_sharded_solve", "pqp_sharded_results",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C proxsuite_b200/csrc). "
            "proxsuite_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
    L.pqp_last_error.restype = C.c_char_p
    L.pqp_version.restype = C.c_char_p
    L.pqp_settings_default.argtypes = [C.POINTER(pqp_settings), C.c_int]
    L.pqp_settings_default.restype = None
    L.pqp_dense_backend_choice.argtypes = [C.c_int, i64, i64, i64, C.c_int]
    L.pqp_batch_create.restype = vp
    L.pqp_batch_create.argtypes = [i64, i64, i64, i64, C.c_int, C.c_int, C.c_int, C.c_int]
    L.pqp_batch_destroy.argtypes = [vp]
    L.pqp_batch_destroy.restype = None
    L.pqp_batch_size.restype = i64
    L.pqp_batch_size.argtypes = [vp]
    L.pqp_batch_dims.argtypes = [vp] + [vp] * 6
    L.pqp_batch_settings_get.argtypes = [vp, i64, C.POINTER(pqp_settings)]
    L.pqp_batch_settings_set.argtypes = [vp, i64, C.POINTER(pqp_settings)]
    data = [vp] * 9
    L.pqp_batch_init.argtypes = [vp, i64, i64] + data + [C.c_int] + [vp] * 4
    L.pqp_batch_init_device.argtypes = [vp, i64, i64] + data + [C.c_int] + [vp] * 4
    L.pqp_batch_update.argtypes = [vp, i64, i64] + data + [C.c_int] + [vp] * 4
    L.pqp_batch_warm_start.argtypes = [vp, i64, i64, vp, vp, vp]
    L.pqp_batch_solve.argtypes = [vp]
    L.pqp_batch_solve_async.argtypes = [vp, vp]
    L.pqp_batch_sync.argtypes = [vp]
    L.pqp_batch_select.argtypes = [vp, vp, i64]
    L.pqp_batch_results.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp]
    L.pqp_batch_results_device.argtypes = [vp, vp, vp, vp, vp]
    L.pqp_batch_scaled.argtypes = [vp, i64] + [vp] * 9
    L.pqp_batch_cleanup.argtypes = [vp, i64, i64]
    L.pqp_batch_backward.argtypes = [vp, i64, i64, vp, dbl, dbl, dbl] + [vp] * 7
    L.pqp_batch_backward_device.argtypes = [vp, i64, i64, vp, dbl, dbl, dbl] + [vp] * 7
    L.pqp_batch_results_copy_device.argtypes = [vp, i64, i64, vp, vp, vp]
    L.pqp_batch_timings.argtypes = [vp, vp, vp, vp]
    L.pqp_batch_debug_trace.argtypes = [vp, vp, i64]
    L.pqp_batch_launch_config.argtypes = [vp, vp, vp, vp, vp]
    L.pqp_batch_profile.argtypes = [vp, vp, C.c_int]
    L.pqp_batch_occupancy.argtypes = [vp, C.c_int]
    L.pqp_random_qp.argtypes = [C.c_int, C.c_uint64, i64, i64, i64, dbl, dbl] + [vp] * 9
    L.pqp_sharded_create.restype = vp
    L.pqp_sharded_create.argtypes = [i64, i64, i64, i64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
    L.pqp_sharded_destroy.argtypes = [vp]
    L.pqp_sharded_destroy.restype = None
    L.pqp_sharded_count.argtypes = [vp]
    L.pqp_sharded_shard.restype = vp
    L.pqp_sharded_shard.argtypes = [vp, C.c_int, vp, vp]
    L.pqp_sharded_settings_set.argtypes = [vp, C.POINTER(pqp_settings)]
    L.pqp_sharded_init.argtypes = [vp] + data + [C.c_int] + [vp] * 4
    L.pqp_sharded_update.argtypes = [vp] + data + [C.c_int] + [vp] * 4
    L.pqp_sharded_solve.argtypes = [vp]
    L.pqp_sharded_results.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    _lib = L
    return L


def last_error() -> str:
    return lib().pqp_last_error().decode()


def check(rc: int):
    if rc == PQP_OK:
        return
    msg = last_error()
    if rc == PQP_EINVAL:
        raise ValueError(msg)  # nanobind surfaces std::invalid_argument as ValueError
    raise RuntimeError(f"proxsuite_b200 error {rc}: {msg}")
