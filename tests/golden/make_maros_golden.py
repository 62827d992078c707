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
    sys.exit(main())
