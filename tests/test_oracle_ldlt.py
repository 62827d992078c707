"""Oracle LDLT (oracle/ldlt.hpp) against dense numpy algebra; mirrors what
test/src/dense_ldlt*.cpp check in the reference: factorise / solve / insert /
delete / diagonal update / rank-r update reproduce the modified matrix."""
import numpy as np
import pytest


def quasi_definite(rng, n, m):
    B = rng.standard_normal((n, n))
    H = B @ B.T + n * np.eye(n) * 0.1
    A = rng.standard_normal((m, n))
    K = np.block([[H, A.T], [A, -1e-2 * np.eye(m)]])
    return K


@pytest.mark.parametrize("n,m", [(5, 2), (20, 10), (40, 33)])
def test_factorize_solve(oracle, n, m):
    rng = np.random.default_rng(n)
    K = quasi_definite(rng, n, m)
    L = oracle.OracleLdlt(K, cap=n + m + 8)
    assert np.allclose(L.reconstruct(), K, atol=1e-9)
    rhs = rng.standard_normal(n + m)
    assert np.allclose(K @ L.solve(rhs), rhs, atol=1e-8)


@pytest.mark.parametrize("r", [1, 2, 3, 5])
def test_insert_then_delete(oracle, r):
    rng = np.random.default_rng(r)
    n, m = 12, 5
    K = quasi_definite(rng, n, m)
    L = oracle.OracleLdlt(K, cap=n + m + 2 * r + 2)
    dim = n + m
    cols = np.zeros((dim + r, r))
    cols[:n, :] = rng.standard_normal((n, r))
    for k in range(r):
        cols[dim + k, k] = -0.1
    L.insert_block_at(dim, cols)
    Kn = np.zeros((dim + r, dim + r))
    Kn[:dim, :dim] = K
    Kn[:, dim:] = cols
    Kn[dim:, :] = cols.T
    assert L.dim() == dim + r
    assert np.allclose(L.reconstruct(), Kn, atol=1e-8)
    rhs = rng.standard_normal(dim + r)
    assert np.allclose(Kn @ L.solve(rhs), rhs, atol=1e-7)
    # delete a subset (sorted indices), including one in the middle of the y block
    dele = sorted({dim + r - 1, n + 1} | ({dim} if r > 1 else set()))
    L.delete_at(dele)
    keep = [i for i in range(dim + r) if i not in dele]
    Kd = Kn[np.ix_(keep, keep)]
    assert L.dim() == len(keep)
    assert np.allclose(L.reconstruct(), Kd, atol=1e-8)


def test_insert_in_the_middle(oracle):
    rng = np.random.default_rng(7)
    n, m = 9, 4
    K = quasi_definite(rng, n, m)
    L = oracle.OracleLdlt(K, cap=n + m + 4)
    dim = n + m
    r = 2
    i = 3
    cols = rng.standard_normal((dim + r, r))
    D = rng.standard_normal((r, r))
    cols[i:i + r, :] = D + D.T + 5 * np.eye(r)
    L.insert_block_at(i, cols)
    idx_old = [k for k in range(dim + r) if not (i <= k < i + r)]
    Kn = np.zeros((dim + r, dim + r))
    Kn[np.ix_(idx_old, idx_old)] = K
    Kn[:, i:i + r] = cols
    Kn[i:i + r, :] = cols.T
    assert np.allclose(L.reconstruct(), Kn, atol=1e-8)


def test_diagonal_and_rank_updates(oracle):
    rng = np.random.default_rng(3)
    n, m = 15, 6
    K = quasi_definite(rng, n, m)
    L = oracle.OracleLdlt(K)
    idx = list(range(n, n + m))
    alpha = np.full(m, 1e-2 - 1e-3)
    L.diagonal_update(idx, alpha)
    K2 = K.copy()
    K2[idx, idx] += alpha
    assert np.allclose(L.reconstruct(), K2, atol=1e-9)
    for r in (1, 4, 6):
        W = rng.standard_normal((n + m, r)) * 0.3
        a = rng.uniform(0.1, 1.0, r)
        L.rank_r_update(W, a)
        K2 = K2 + W @ np.diag(a) @ W.T
        assert np.allclose(L.reconstruct(), K2, atol=1e-8)
