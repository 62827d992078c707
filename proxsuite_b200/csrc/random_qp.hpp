// Deterministic synthetic dense QP generators (host code, no GPU).
//
// Mirrors the input specification every reference test/benchmark uses
// (/root/reference/include/proxsuite/proxqp/utils/random_qp_problems.hpp):
//   :104-147  Lehmer-64 generator, set_seed, uniform_rand, normal_rand
//   :308-334  sparse_positive_definite_rand_not_compressed
//   :354-368  sparse_matrix_rand_not_compressed
//   :463-502  dense_strongly_convex_qp
//   :505-543  dense_not_strongly_convex_qp
//   :546-590  dense_degenerate_qp
//   :592-628  dense_box_constrained_qp
// and the extra draws of benchmark/timings-box-constraints.cpp:33-51 and
// benchmark/timings-diagonal-hessian.cpp:33-56.
//
// All matrices are produced ROW-MAJOR (the solver's layout, dense/fwd.hpp:16-33).
// The only step that is not bit-reproducible against the reference is
// lambda_min(H) (Eigen's SelfAdjointEigenSolver there, Householder + implicit
// QL here); both agree to rounding and every consumer (oracle, GPU path, CPU
// baseline) is fed the same generated arrays.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace pqp {
namespace randqp {

using u64 = std::uint64_t;
using u128 = unsigned __int128;

struct Lehmer
{
  u128 state;
  static constexpr u64 kMul = 0xda942042e4dd58b5ULL;
  Lehmer()
    : state(u128(kMul) * u128(kMul))
  {
  }
  u64 next()
  {
    state *= u128(kMul);
    return u64(state >> 64);
  }
  void set_seed(u64 seed)
  {
    state = u128(seed) + 1;
    next();
    next();
  }
  double uniform()
  {
    u64 a = next() / (u64(1) << 11);
    return double(a) / double(u64(1) << 53);
  }
  double normal()
  {
    static const double pi2 = std::atan(1.0) * 8;
    double u1 = uniform();
    double u2 = uniform();
    double ln = std::log(u1);
    double sq = std::sqrt(-2 * ln);
    return sq * std::cos(pi2 * u2);
  }
};

// Smallest eigenvalue of a symmetric matrix (row-major n x n, destroyed):
// Householder tridiagonalisation followed by implicit-shift QL.
inline double
sym_min_eigenvalue(std::vector<double> a, int n)
{
  std::vector<double> d(static_cast<std::size_t>(n)), e(static_cast<std::size_t>(n));
  auto A = [&](int i, int j) -> double& { return a[std::size_t(i) * std::size_t(n) + std::size_t(j)]; };
  for (int i = n - 1; i > 0; --i) {
    int l = i - 1;
    double h = 0, scale = 0;
    if (l > 0) {
      for (int k = 0; k <= l; ++k) {
        scale += std::fabs(A(i, k));
      }
      if (scale == 0.0) {
        e[std::size_t(i)] = A(i, l);
      } else {
        for (int k = 0; k <= l; ++k) {
          A(i, k) /= scale;
          h += A(i, k) * A(i, k);
        }
        double f = A(i, l);
        double g = f >= 0 ? -std::sqrt(h) : std::sqrt(h);
        e[std::size_t(i)] = scale * g;
        h -= f * g;
        A(i, l) = f - g;
        f = 0;
        for (int j = 0; j <= l; ++j) {
          g = 0;
          for (int k = 0; k <= j; ++k) {
            g += A(j, k) * A(i, k);
          }
          for (int k = j + 1; k <= l; ++k) {
            g += A(k, j) * A(i, k);
          }
          e[std::size_t(j)] = g / h;
          f += e[std::size_t(j)] * A(i, j);
        }
        double hh = f / (h + h);
        for (int j = 0; j <= l; ++j) {
          f = A(i, j);
          e[std::size_t(j)] = g = e[std::size_t(j)] - hh * f;
          for (int k = 0; k <= j; ++k) {
            A(j, k) -= (f * e[std::size_t(k)] + g * A(i, k));
          }
        }
      }
    } else {
      e[std::size_t(i)] = A(i, l);
    }
    d[std::size_t(i)] = h;
  }
  e[0] = 0;
  for (int i = 0; i < n; ++i) {
    d[std::size_t(i)] = A(i, i);
  }
  // QL with implicit shifts (eigenvalues only)
  for (int i = 1; i < n; ++i) {
    e[std::size_t(i - 1)] = e[std::size_t(i)];
  }
  e[std::size_t(n - 1)] = 0;
  for (int l = 0; l < n; ++l) {
    int iter = 0;
    int m;
    do {
      for (m = l; m < n - 1; ++m) {
        double dd = std::fabs(d[std::size_t(m)]) + std::fabs(d[std::size_t(m + 1)]);
        if (std::fabs(e[std::size_t(m)]) <= 2.3e-16 * dd) {
          break;
        }
      }
      if (m != l) {
        if (iter++ == 200) {
          break;
        }
        double g = (d[std::size_t(l + 1)] - d[std::size_t(l)]) / (2.0 * e[std::size_t(l)]);
        double r = std::hypot(g, 1.0);
        g = d[std::size_t(m)] - d[std::size_t(l)] + e[std::size_t(l)] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1, c = 1, p = 0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[std::size_t(i)];
          double b = c * e[std::size_t(i)];
          e[std::size_t(i + 1)] = (r = std::hypot(f, g));
          if (r == 0.0) {
            d[std::size_t(i + 1)] -= p;
            e[std::size_t(m)] = 0;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[std::size_t(i + 1)] - p;
          r = (d[std::size_t(i)] - g) * s + 2.0 * c * b;
          d[std::size_t(i + 1)] = g + (p = s * r);
          g = c * r - b;
        }
        if (r == 0.0 && i >= l) {
          continue;
        }
        d[std::size_t(l)] -= p;
        e[std::size_t(l)] = g;
        e[std::size_t(m)] = 0;
      }
    } while (m != l);
  }
  double mn = d[0];
  for (int i = 1; i < n; ++i) {
    mn = std::fmin(mn, d[std::size_t(i)]);
  }
  return mn;
}

// random_qp_problems.hpp:308-334 (row-major n x n output)
inline void
sparse_positive_definite_rand(Lehmer& rng, int n, double rho, double p, double* H)
{
  std::vector<double> T(std::size_t(n) * std::size_t(n), 0.0);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) {
      double urandom = rng.uniform();
      if (urandom < p / 2) {
        T[std::size_t(i) * std::size_t(n) + std::size_t(j)] = rng.normal();
      }
    }
  }
  std::vector<double> S(std::size_t(n) * std::size_t(n));
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) {
      S[std::size_t(i) * std::size_t(n) + std::size_t(j)] = (T[std::size_t(i) * std::size_t(n) + std::size_t(j)] + T[std::size_t(j) * std::size_t(n) + std::size_t(i)]) * 0.5;
    }
  }
  double mn = sym_min_eigenvalue(S, n);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) {
      H[std::size_t(i) * std::size_t(n) + std::size_t(j)] = S[std::size_t(i) * std::size_t(n) + std::size_t(j)];
    }
    H[std::size_t(i) * std::size_t(n) + std::size_t(i)] += rho + std::fabs(mn);
  }
}

// random_qp_problems.hpp:354-368 (row-major rows x cols output)
inline void
sparse_matrix_rand(Lehmer& rng, int rows, int cols, double p, double* A)
{
  for (int i = 0; i < rows; ++i) {
    for (int j = 0; j < cols; ++j) {
      double v = 0;
      if (rng.uniform() < p) {
        v = rng.normal();
      }
      A[std::size_t(i) * std::size_t(cols) + std::size_t(j)] = v;
    }
  }
}

inline void
matvec(const double* M, int rows, int cols, const double* x, double* y)
{
  for (int i = 0; i < rows; ++i) {
    double s = 0;
    for (int j = 0; j < cols; ++j) {
      s += M[std::size_t(i) * std::size_t(cols) + std::size_t(j)] * x[j];
    }
    y[i] = s;
  }
}

// random_qp_problems.hpp:463-502. Draw order: H, g, A, C, x_sol, delta.
inline void
dense_strongly_convex_qp(Lehmer& rng, int n, int n_eq, int n_in, double sparsity, double strong_convexity, double* H, double* g, double* A, double* b, double* C, double* u, double* l)
{
  sparse_positive_definite_rand(rng, n, strong_convexity, sparsity, H);
  for (int i = 0; i < n; ++i) {
    g[i] = rng.normal();
  }
  sparse_matrix_rand(rng, n_eq, n, sparsity, A);
  sparse_matrix_rand(rng, n_in, n, sparsity, C);
  std::vector<double> x_sol(static_cast<std::size_t>(n)), delta(static_cast<std::size_t>(n_in));
  for (int i = 0; i < n; ++i) {
    x_sol[std::size_t(i)] = rng.normal();
  }
  for (int i = 0; i < n_in; ++i) {
    delta[std::size_t(i)] = rng.uniform();
  }
  matvec(C, n_in, n, x_sol.data(), u);
  for (int i = 0; i < n_in; ++i) {
    u[i] += delta[std::size_t(i)];
    l[i] = -1.e20;
  }
  matvec(A, n_eq, n, x_sol.data(), b);
}

// random_qp_problems.hpp:505-543. Draw order: H, A, C, x_sol, y_sol, z_sol, delta.
inline void
dense_not_strongly_convex_qp(Lehmer& rng, int n, int n_eq, int n_in, double sparsity, double* H, double* g, double* A, double* b, double* C, double* u, double* l)
{
  sparse_positive_definite_rand(rng, n, 0.0, sparsity, H);
  sparse_matrix_rand(rng, n_eq, n, sparsity, A);
  sparse_matrix_rand(rng, n_in, n, sparsity, C);
  std::vector<double> x_sol(static_cast<std::size_t>(n)), y_sol(static_cast<std::size_t>(n_eq)), z_sol(static_cast<std::size_t>(n_in)), delta(static_cast<std::size_t>(n_in));
  for (auto& v : x_sol) v = rng.normal();
  for (auto& v : y_sol) v = rng.normal();
  for (auto& v : z_sol) v = rng.normal();
  for (auto& v : delta) v = rng.uniform();
  matvec(C, n_in, n, x_sol.data(), u);
  for (int i = 0; i < n_in; ++i) {
    double cx = u[i];
    u[i] = cx + delta[std::size_t(i)];
    l[i] = cx - delta[std::size_t(i)];
  }
  matvec(A, n_eq, n, x_sol.data(), b);
  matvec(H, n, n, x_sol.data(), g);
  for (int j = 0; j < n; ++j) {
    double s = g[j];
    for (int i = 0; i < n_in; ++i) {
      s += C[std::size_t(i) * std::size_t(n) + std::size_t(j)] * z_sol[std::size_t(i)];
    }
    for (int i = 0; i < n_eq; ++i) {
      s += A[std::size_t(i) * std::size_t(n) + std::size_t(j)] * y_sol[std::size_t(i)];
    }
    g[j] = -s;
  }
}

// random_qp_problems.hpp:546-590. C has 2*n_in rows (the same block twice).
inline void
dense_degenerate_qp(Lehmer& rng, int n, int n_eq, int n_in, double sparsity, double strong_convexity, double* H, double* g, double* A, double* b, double* C /*2 n_in x n*/, double* u /*2 n_in*/, double* l /*2 n_in*/)
{
  sparse_positive_definite_rand(rng, n, strong_convexity, sparsity, H);
  for (int i = 0; i < n; ++i) {
    g[i] = rng.normal();
  }
  sparse_matrix_rand(rng, n_eq, n, sparsity, A);
  std::vector<double> x_sol(static_cast<std::size_t>(n)), delta(std::size_t(2 * n_in));
  for (auto& v : x_sol) v = rng.normal();
  for (auto& v : delta) v = rng.uniform();
  matvec(A, n_eq, n, x_sol.data(), b);
  sparse_matrix_rand(rng, n_in, n, sparsity, C);
  for (int i = 0; i < n_in; ++i) {
    for (int j = 0; j < n; ++j) {
      C[std::size_t(n_in + i) * std::size_t(n) + std::size_t(j)] = C[std::size_t(i) * std::size_t(n) + std::size_t(j)];
    }
  }
  matvec(C, 2 * n_in, n, x_sol.data(), u);
  for (int i = 0; i < 2 * n_in; ++i) {
    u[i] += delta[std::size_t(i)];
    l[i] = -1.e20;
  }
}

// random_qp_problems.hpp:592-628 (requires n_in == n: C = I). Draw order: H, g, A, x_sol, delta.
inline void
dense_box_constrained_qp(Lehmer& rng, int n, int n_eq, int n_in, double sparsity, double strong_convexity, double* H, double* g, double* A, double* b, double* C, double* u, double* l)
{
  sparse_positive_definite_rand(rng, n, strong_convexity, sparsity, H);
  for (int i = 0; i < n; ++i) {
    g[i] = rng.normal();
  }
  sparse_matrix_rand(rng, n_eq, n, sparsity, A);
  std::vector<double> x_sol(static_cast<std::size_t>(n)), delta(static_cast<std::size_t>(n_in));
  for (auto& v : x_sol) v = rng.normal();
  for (auto& v : delta) v = rng.uniform();
  matvec(A, n_eq, n, x_sol.data(), b);
  for (int i = 0; i < n_in; ++i) {
    for (int j = 0; j < n; ++j) {
      C[std::size_t(i) * std::size_t(n) + std::size_t(j)] = (i == j) ? 1.0 : 0.0;
    }
    u[i] = x_sol[std::size_t(i)] + delta[std::size_t(i)];
    l[i] = x_sol[std::size_t(i)] - delta[std::size_t(i)];
  }
}

// benchmark/timings-box-constraints.cpp:30-51: dense_strongly_convex_qp, then a
// second x_sol / delta draw that replaces u and b, then per-coordinate box
// bounds x_sol[i] +- U(0,1). With diagonal_hessian != 0 the Hessian is
// replaced by diag(0, 1, ..., n-1) (benchmark/timings-diagonal-hessian.cpp:52-56).
inline void
dense_box_benchmark_qp(Lehmer& rng, int n, int n_eq, int n_in, double sparsity, double strong_convexity, int diagonal_hessian, double* H, double* g, double* A, double* b, double* C, double* u, double* l, double* u_box, double* l_box)
{
  dense_strongly_convex_qp(rng, n, n_eq, n_in, sparsity, strong_convexity, H, g, A, b, C, u, l);
  std::vector<double> x_sol(static_cast<std::size_t>(n)), delta(static_cast<std::size_t>(n_in));
  for (auto& v : x_sol) v = rng.normal();
  for (auto& v : delta) v = rng.uniform();
  matvec(C, n_in, n, x_sol.data(), u);
  for (int i = 0; i < n_in; ++i) {
    u[i] += delta[std::size_t(i)];
  }
  matvec(A, n_eq, n, x_sol.data(), b);
  for (int i = 0; i < n; ++i) {
    double shift = rng.uniform();
    u_box[i] = x_sol[std::size_t(i)] + shift;
    l_box[i] = x_sol[std::size_t(i)] - shift;
  }
  if (diagonal_hessian) {
    for (int i = 0; i < n; ++i) {
      for (int j = 0; j < n; ++j) {
        H[std::size_t(i) * std::size_t(n) + std::size_t(j)] = (i == j) ? double(i) : 0.0;
      }
    }
  }
}

} // namespace randqp
} // namespace pqp
