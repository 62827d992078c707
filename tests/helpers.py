"""Shared helpers for the test-suite: optimality residuals recomputed from the
ORIGINAL data, exactly as the reference's tests do
(test/src/dense_qp_with_eq_and_in.cpp:46-56, dense_qp_wrapper.cpp:6803-6900)."""
import numpy as np


def kkt_residuals(d, x, y, z):
    """pri_res = max(|Ax-b|_inf, |[Cx-u]_+ + [Cx-l]_-|_inf [, box]),
    dua_res = |Hx+g+A'y+C'z_C (+ z_box)|_inf."""
    n = d["H"].shape[0]
    n_in = d["C"].shape[0] if d.get("C") is not None else 0
    pri = 0.0
    dual = d["H"] @ x + d["g"]
    if d.get("A") is not None and d["A"].shape[0] > 0:
        pri = max(pri, np.abs(d["A"] @ x - d["b"]).max())
        dual = dual + d["A"].T @ y
    if n_in > 0:
        cx = d["C"] @ x
        pri = max(pri, np.abs(np.maximum(cx - d["u"], 0) + np.minimum(cx - d["l"], 0)).max())
        dual = dual + d["C"].T @ z[:n_in]
    if d.get("u_box") is not None:
        pri = max(pri, np.abs(np.maximum(x - d["u_box"], 0) + np.minimum(x - d["l_box"], 0)).max())
        dual = dual + z[n_in:n_in + n]
    return float(pri), float(np.abs(dual).max())
