"""Builds tests/golden/maros_meszaros_small.npz from the reference's Maros-Meszaros data
(/root/reference/test/data/maros_meszaros_data/*.mat; run in the build container only -
the GPU box has no /root/reference). Follows test/include/maros_meszaros.hpp:20-98 (load)
and :121-160 (preprocess_qp: rows with l == u become equalities, the rest inequalities),
and keeps the problems of test/src/dense_maros_meszaros.cpp:13-85 with n <= 120 and at most
260 constraint rows. Matrices are stored dense (they are tiny), compressed."""
import glob
import os
import sys
import warnings

import numpy as np
import scipy.io

SRC = "/root/reference/test/data/maros_meszaros_data/"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "maros_meszaros_small.npz")


def load(path):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = scipy.io.loadmat(path)
    P = np.asarray(m["P"].todense(), dtype=np.float64)
    A = np.asarray(m["A"].todense(), dtype=np.float64)
    q = np.asarray(m["q"], dtype=np.float64).ravel()
    l = np.asarray(m["l"], dtype=np.float64).ravel()
    u = np.asarray(m["u"], dtype=np.float64).ravel()
    eq = l == u
    return dict(H=P, g=q, A=A[eq], b=l[eq], C=A[~eq], l=l[~eq], u=u[~eq])


def main_large():
    """The REST of the problems the reference's dense test actually runs (test/src/dense_maros_meszaros.cpp:95-99 skips
    n > 1000 or more than 1000 constraint rows): 34 problems with n up to 760, stored sparse (COO) in
    maros_meszaros_large.npz; tests/test_oracle_maros.py rebuilds the dense arrays."""
    import re
    src = open("/root/reference/test/src/dense_maros_meszaros.cpp").read()
    listed = re.findall(r'MAROS_MESZAROS_DIR "([^"]+)\.mat"', src)
    out, names = {}, []
    for name in listed:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = scipy.io.loadmat(SRC + name + ".mat")
        n, rows = m["P"].shape[0], m["A"].shape[0]
        if n > 1000 or rows > 1000 or (n <= 120 and rows <= 260):
            continue
        d = load(SRC + name + ".mat")
        names.append(name)
        for k in ("H", "A", "C"):
            r, c = np.nonzero(d[k])
            out[f"{name}/{k}_shape"] = np.array(d[k].shape, dtype=np.int32)
            out[f"{name}/{k}_rc"] = np.stack([r, c]).astype(np.int32)
            out[f"{name}/{k}_v"] = d[k][r, c]
        for k in ("g", "b", "l", "u"):
            out[f"{name}/{k}"] = d[k]
    out["names"] = np.array(names)
    path = os.path.join(os.path.dirname(OUT), "maros_meszaros_large.npz")
    np.savez_compressed(path, **out)
    print(len(names), "problems ->", path, os.path.getsize(path), "bytes")


def main():
    out = {}
    names = []
    for path in sorted(glob.glob(SRC + "*.mat")):
        name = os.path.basename(path)[:-4]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = scipy.io.loadmat(path)
        n, rows = m["P"].shape[0], m["A"].shape[0]
        if n > 120 or rows > 260:
            continue
        d = load(path)
        names.append(name)
        for k, v in d.items():
            out[f"{name}/{k}"] = v
    out["names"] = np.array(names)
    np.savez_compressed(OUT, **out)
    print(len(names), "problems ->", OUT, os.path.getsize(OUT), "bytes")
    print(" ".join(names))


if __name__ == "__main__":
    main()
    main_large()
