// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (plain C++17 + OpenMP, no Eigen) of the reference's dense
// ProxQP path: proxsuite::proxqp::dense::QP / BatchQP + solve_in_parallel.
// The reference itself cannot be compiled in this image (Eigen 3 is an
// un-vendored dependency, CMakeLists.txt:180, and is not installed), so this
// restatement is the checker. It is pinned by the reference's own known-answer
// tests and acceptance criteria (tests/test_oracle_*.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use anything under oracle/.
//
// Follows (file:line under /root/reference/include/proxsuite/proxqp):
//   settings.hpp:88-316          Settings<T> and defaults
//   results.hpp:28-203           Info<T>, Results<T>, cleanup/cold_start
//   status.hpp:17-43             enums
//   dense/model.hpp:23-149       Model<T>
//   dense/workspace.hpp:25-378   Workspace<T> (state + cleanup semantics)
//   dense/preconditioner/ruiz.hpp:31-311, 403-694   Ruiz equilibration
//   dense/helpers.hpp:176-285, 300-329, 374-763     setup / update / factorization
//   dense/utils.hpp:166-587      global residuals and infeasibility tests
//   dense/linesearch.hpp:51-786  exact line search, active_set_change
//   dense/solver.hpp:40-1843     refactorize, mu_update, iterative refinement,
//                                Newton loops, BCL, qp_solve
//   dense/wrapper.hpp:82-113, 354-962, 1253-1311    QP, BatchQP
//   parallel/qp_solve.hpp:17-60  solve_in_parallel
#pragma once
#include "ldlt.hpp"
#include <cfloat>
#include <cstdio>
#include <limits>
#include <stdexcept>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace oracle {

using Vec = std::vector<double>;

// status.hpp:17-26
enum QPSolverOutput
{
  PROXQP_SOLVED = 0,
  PROXQP_MAX_ITER_REACHED = 1,
  PROXQP_PRIMAL_INFEASIBLE = 2,
  PROXQP_SOLVED_CLOSEST_PRIMAL_FEASIBLE = 3,
  PROXQP_DUAL_INFEASIBLE = 4,
  PROXQP_NOT_RUN = 5
};
// status.hpp:28-35
enum InitialGuessStatus
{
  NO_INITIAL_GUESS = 0,
  EQUALITY_CONSTRAINED_INITIAL_GUESS = 1,
  WARM_START_WITH_PREVIOUS_RESULT = 2,
  WARM_START = 3,
  COLD_START_WITH_PREVIOUS_RESULT = 4
};
enum PreconditionerStatus
{
  PRECOND_EXECUTE = 0,
  PRECOND_KEEP = 1,
  PRECOND_IDENTITY = 2
};
// settings.hpp:26-45
enum DenseBackend
{
  BACKEND_AUTOMATIC = 0,
  BACKEND_PRIMAL_DUAL_LDLT = 1,
  BACKEND_PRIMAL_LDLT = 2
};
enum MeritFunctionType
{
  MERIT_GPDAL = 0,
  MERIT_PDAL = 1
};
enum HessianType
{
  HESSIAN_ZERO = 0,
  HESSIAN_DENSE = 1,
  HESSIAN_DIAGONAL = 2
};

// helpers/common.hpp:17-24
inline double
infinite_bound()
{
  return std::sqrt(std::numeric_limits<double>::max());
}

// settings.hpp:88-316
struct Settings
{
  double default_rho = 1e-6;
  double default_mu_eq = 1e-3;
  double default_mu_in = 1e-1;
  double alpha_bcl = 0.1;
  double beta_bcl = 0.9;
  double refactor_dual_feasibility_threshold = 1e-2;
  double refactor_rho_threshold = 1e-7;
  double mu_min_eq = 1e-9;
  double mu_min_in = 1e-8;
  double mu_max_eq_inv = 1e9;
  double mu_max_in_inv = 1e8;
  double mu_update_factor = 0.1;
  double mu_update_inv_factor = 10;
  double cold_reset_mu_eq = 1. / 1.1;
  double cold_reset_mu_in = 1. / 1.1;
  double cold_reset_mu_eq_inv = 1.1;
  double cold_reset_mu_in_inv = 1.1;
  double eps_abs = 1e-5;
  double eps_rel = 0;
  isize max_iter = 10000;
  isize max_iter_in = 1500;
  isize safe_guard = 10000;
  isize nb_iterative_refinement = 10;
  double eps_refact = 1e-6;
  bool verbose = false;
  int initial_guess = EQUALITY_CONSTRAINED_INITIAL_GUESS;
  bool update_preconditioner = false;
  bool compute_preconditioner = true;
  bool compute_timings = false;
  bool check_duality_gap = false;
  double eps_duality_gap_abs = 1e-4;
  double eps_duality_gap_rel = 0;
  isize preconditioner_max_iter = 10;
  double preconditioner_accuracy = 1e-3;
  double eps_primal_inf = 1e-4;
  double eps_dual_inf = 1e-4;
  bool bcl_update = true;
  int merit_function_type = MERIT_GPDAL;
  double alpha_gpdal = 0.95;
  bool primal_infeasibility_solving = false;
  isize frequence_infeasibility_check = 1;
  double default_H_eigenvalue_estimate = 0.;
  explicit Settings(int dense_backend = BACKEND_PRIMAL_DUAL_LDLT)
  {
    default_rho = dense_backend == BACKEND_PRIMAL_LDLT ? 1e-5 : 1e-6; // settings.hpp:302-313
  }
};

// results.hpp:28-58
struct Info
{
  double mu_eq = 1e-3, mu_eq_inv = 1e3, mu_in = 1e-1, mu_in_inv = 1e1, rho = 1e-6, nu = 1.;
  isize iter = 0, iter_ext = 0, mu_updates = 0, rho_updates = 0;
  int status = PROXQP_NOT_RUN;
  double setup_time = 0, solve_time = 0, run_time = 0;
  double objValue = 0, pri_res = 0, dua_res = 0, duality_gap = 0, iterative_residual = 0;
  double minimal_H_eigenvalue_estimate = 0;
};

// results.hpp:67-203
struct Results
{
  Vec x, y, z, se, si;
  Info info;
  Results() = default;
  Results(isize dim, isize n_eq, isize n_in, bool box, int backend)
  {
    x.assign(std::size_t(dim), 0);
    y.assign(std::size_t(n_eq), 0);
    isize nc = n_in + (box ? dim : 0);
    z.assign(std::size_t(nc), 0);
    se.assign(std::size_t(n_eq), 0);
    si.assign(std::size_t(nc), 0);
    info.rho = backend == BACKEND_PRIMAL_LDLT ? 1e-5 : 1e-6;
  }
  void zero_vars()
  {
    std::fill(x.begin(), x.end(), 0.);
    std::fill(y.begin(), y.end(), 0.);
    std::fill(z.begin(), z.end(), 0.);
    std::fill(se.begin(), se.end(), 0.);
    std::fill(si.begin(), si.end(), 0.);
  }
  void cleanup_statistics()
  {
    info.run_time = 0;
    info.setup_time = 0;
    info.solve_time = 0;
    info.objValue = 0;
    info.iter = 0;
    info.iter_ext = 0;
    info.mu_updates = 0;
    info.rho_updates = 0;
    info.pri_res = 0;
    info.dua_res = 0;
    info.duality_gap = 0;
    info.iterative_residual = 0;
    info.status = PROXQP_MAX_ITER_REACHED;
  }
  void cold_start(const Settings* s)
  {
    info.rho = 1e-6;
    info.mu_eq_inv = 1e3;
    info.mu_eq = 1e-3;
    info.mu_in_inv = 1e1;
    info.mu_in = 1e-1;
    info.nu = 1.;
    info.minimal_H_eigenvalue_estimate = 0.;
    if (s) {
      info.rho = s->default_rho;
      info.mu_eq = s->default_mu_eq;
      info.mu_eq_inv = 1. / info.mu_eq;
      info.mu_in = s->default_mu_in;
      info.mu_in_inv = 1. / info.mu_in;
      info.minimal_H_eigenvalue_estimate = s->default_H_eigenvalue_estimate;
    }
    cleanup_statistics();
  }
  void cleanup(const Settings* s)
  {
    zero_vars();
    cold_start(s);
  }
  void cleanup_all_except_prox_parameters()
  {
    zero_vars();
    cleanup_statistics();
  }
};

struct Mat
{
  isize rows = 0, cols = 0;
  Vec a; // row-major
  Mat() = default;
  Mat(isize r, isize c)
    : rows(r)
    , cols(c)
    , a(std::size_t(r * c), 0.)
  {
  }
  double& operator()(isize i, isize j) { return a[std::size_t(i * cols + j)]; }
  double operator()(isize i, isize j) const { return a[std::size_t(i * cols + j)]; }
  const double* row(isize i) const { return a.data() + i * cols; }
  double* row(isize i) { return a.data() + i * cols; }
  void set_zero() { std::fill(a.begin(), a.end(), 0.); }
};

// dense/model.hpp:23-149
struct Model
{
  isize dim = 0, n_eq = 0, n_in = 0;
  Mat H, A, C;
  Vec g, b, u, l, u_box, l_box;
  Model() = default;
  Model(isize dim_, isize n_eq_, isize n_in_, bool box)
    : dim(dim_)
    , n_eq(n_eq_)
    , n_in(n_in_)
    , H(dim_, dim_)
    , A(n_eq_, dim_)
    , C(n_in_, dim_)
  {
    if (dim == 0) {
      throw std::invalid_argument("wrong argument size: the dimension wrt the primal variable x should be strictly positive.");
    }
    g.assign(std::size_t(dim), 0.);
    b.assign(std::size_t(n_eq), 0.);
    u.assign(std::size_t(n_in), infinite_bound());
    l.assign(std::size_t(n_in), -infinite_bound());
    if (box) {
      u_box.assign(std::size_t(dim), infinite_bound());
      l_box.assign(std::size_t(dim), -infinite_bound());
    }
  }
};

inline double
infty_norm(const double* v, isize n)
{
  double m = 0;
  for (isize i = 0; i < n; ++i) {
    double a = std::fabs(v[i]);
    if (a > m || a != a) {
      m = a;
    }
  }
  return m;
}
inline double
infty_norm(const Vec& v)
{
  return infty_norm(v.data(), isize(v.size()));
}
inline double
dot(const double* a, const double* b, isize n)
{
  double s = 0;
  for (isize i = 0; i < n; ++i) {
    s += a[i] * b[i];
  }
  return s;
}
inline double
pos_part(double v)
{
  return v >= 0 ? v : 0.;
}
inline double
neg_part(double v)
{
  return v <= 0 ? v : 0.;
}
// y = M x, row-major
inline void
gemv(const Mat& M, const double* x, double* y)
{
  for (isize i = 0; i < M.rows; ++i) {
    y[i] = dot(M.row(i), x, M.cols);
  }
}
// y (+)= M^T x
inline void
gemv_t(const Mat& M, const double* x, double* y, bool accumulate)
{
  if (!accumulate) {
    for (isize j = 0; j < M.cols; ++j) {
      y[j] = 0;
    }
  }
  for (isize i = 0; i < M.rows; ++i) {
    const double xi = x[i];
    const double* __restrict r = M.row(i);
    for (isize j = 0; j < M.cols; ++j) {
      y[j] += r[j] * xi;
    }
  }
}
// y = sym(lower(H)) x   (selfadjointView<Lower>)
inline void
symv_lower(const Mat& H, const double* x, double* y)
{
  isize n = H.rows;
  for (isize i = 0; i < n; ++i) {
    y[i] = 0;
  }
  for (isize i = 0; i < n; ++i) {
    const double* r = H.row(i);
    double acc = 0;
    const double xi = x[i];
    for (isize j = 0; j < i; ++j) {
      acc += r[j] * x[j];
      y[j] += r[j] * xi;
    }
    y[i] += acc + r[i] * xi;
  }
}

// dense/preconditioner/ruiz.hpp:316-695
struct Ruiz
{
  Vec delta;
  double c = 1;
  isize dim = 0, n_eq = 0, n_in = 0;
  Ruiz() = default;
  Ruiz(isize dim_, isize n_eq_, isize n_in_, bool box)
    : delta(std::size_t(dim_ + n_eq_ + n_in_ + (box ? dim_ : 0)), 1.)
    , c(1)
    , dim(dim_)
    , n_eq(n_eq_)
    , n_in(n_in_)
  {
  }
  const double* dx() const { return delta.data(); }
  const double* deq() const { return delta.data() + dim; }
  const double* din() const { return delta.data() + dim + n_eq; }
  const double* dbox() const { return delta.data() + (isize(delta.size()) - dim); }
  void scale_primal(double* v) const { for (isize i = 0; i < dim; ++i) v[i] /= dx()[i]; }
  void unscale_primal(double* v) const { for (isize i = 0; i < dim; ++i) v[i] *= dx()[i]; }
  void scale_dual_eq(double* v) const { for (isize i = 0; i < n_eq; ++i) v[i] = v[i] / deq()[i] * c; }
  void unscale_dual_eq(double* v) const { for (isize i = 0; i < n_eq; ++i) v[i] = v[i] * deq()[i] / c; }
  void scale_dual_in(double* v) const { for (isize i = 0; i < n_in; ++i) v[i] = v[i] / din()[i] * c; }
  void unscale_dual_in(double* v) const { for (isize i = 0; i < n_in; ++i) v[i] = v[i] * din()[i] / c; }
  void scale_box_dual_in(double* v) const { for (isize i = 0; i < dim; ++i) v[i] = v[i] / dbox()[i] * c; }
  void unscale_box_dual_in(double* v) const { for (isize i = 0; i < dim; ++i) v[i] = dbox()[i] * v[i] / c; }
  void scale_primal_residual_eq(double* v) const { for (isize i = 0; i < n_eq; ++i) v[i] *= deq()[i]; }
  void unscale_primal_residual_eq(double* v) const { for (isize i = 0; i < n_eq; ++i) v[i] /= deq()[i]; }
  void scale_primal_residual_in(double* v) const { for (isize i = 0; i < n_in; ++i) v[i] *= din()[i]; }
  void unscale_primal_residual_in(double* v) const { for (isize i = 0; i < n_in; ++i) v[i] /= din()[i]; }
  void scale_box_primal_residual_in(double* v) const { for (isize i = 0; i < dim; ++i) v[i] *= dbox()[i]; }
  void unscale_box_primal_residual_in(double* v) const { for (isize i = 0; i < dim; ++i) v[i] /= dbox()[i]; }
  void scale_dual_residual(double* v) const { for (isize i = 0; i < dim; ++i) v[i] *= dx()[i] * c; }
  void unscale_dual_residual(double* v) const { for (isize i = 0; i < dim; ++i) v[i] /= dx()[i] * c; }
};

// dense/workspace.hpp:25-378
struct Workspace
{
  Ldlt ldl;
  Mat H_scaled, A_scaled, C_scaled;
  Vec g_scaled, b_scaled, u_scaled, l_scaled, u_box_scaled, l_box_scaled, i_scaled;
  Vec x_prev, y_prev, z_prev;
  Mat kkt;
  std::vector<isize> current_bijection_map, new_bijection_map;
  std::vector<unsigned char> active_set_up, active_set_low, active_inequalities;
  Vec Hdx, Cdx, Adx, active_part_z, alphas;
  Vec dw_aug, rhs, err;
  double dual_feasibility_rhs_2 = 0, correction_guess_rhs_g = 0, alpha = 1;
  Vec dual_residual_scaled, primal_residual_in_scaled_up;
  Vec up_plus_alphaCdx, low_plus_alphaCdx, CTz;
  Vec tmp_n, new_cols;
  std::vector<isize> ibuf;
  bool constraints_changed = false, dirty = false, refactorize = false, proximal_parameter_update = false, is_initialized = false;
  isize n_c = 0;
  isize max_nc = 0; // largest active-set size seen (statistics only)
  Counters cnt;

  Workspace() = default;
  Workspace(isize dim, isize n_eq, isize n_in, bool box, int backend)
    : H_scaled(dim, dim)
    , A_scaled(n_eq, dim)
    , C_scaled(n_in, dim)
  {
    isize ncons = n_in + (box ? dim : 0);
    g_scaled.assign(std::size_t(dim), 0);
    b_scaled.assign(std::size_t(n_eq), 0);
    u_scaled.assign(std::size_t(n_in), 0);
    l_scaled.assign(std::size_t(n_in), 0);
    if (box) {
      u_box_scaled.assign(std::size_t(dim), 0);
      l_box_scaled.assign(std::size_t(dim), 0);
      i_scaled.assign(std::size_t(dim), 1.);
    }
    x_prev.assign(std::size_t(dim), 0);
    y_prev.assign(std::size_t(n_eq), 0);
    z_prev.assign(std::size_t(ncons), 0);
    if (backend == BACKEND_PRIMAL_LDLT) {
      kkt = Mat(dim, dim);
      ldl.reserve(dim);
    } else {
      kkt = Mat(dim + n_eq, dim + n_eq);
      ldl.reserve(dim + n_eq + ncons);
    }
    current_bijection_map.resize(std::size_t(ncons));
    new_bijection_map.resize(std::size_t(ncons));
    for (isize i = 0; i < ncons; ++i) {
      current_bijection_map[std::size_t(i)] = i;
      new_bijection_map[std::size_t(i)] = i;
    }
    active_set_up.assign(std::size_t(ncons), 0);
    active_set_low.assign(std::size_t(ncons), 0);
    active_inequalities.assign(std::size_t(ncons), 0);
    Hdx.assign(std::size_t(dim), 0);
    Cdx.assign(std::size_t(ncons), 0);
    Adx.assign(std::size_t(n_eq), 0);
    active_part_z.assign(std::size_t(ncons), 0);
    alphas.reserve(std::size_t(2 * ncons));
    dw_aug.assign(std::size_t(dim + n_eq + ncons), 0);
    rhs.assign(std::size_t(dim + n_eq + ncons), 0);
    err.assign(std::size_t(dim + n_eq + ncons), 0);
    dual_residual_scaled.assign(std::size_t(dim), 0);
    primal_residual_in_scaled_up.assign(std::size_t(ncons), 0);
    up_plus_alphaCdx.assign(std::size_t(ncons), 0);
    low_plus_alphaCdx.assign(std::size_t(ncons), 0);
    CTz.assign(std::size_t(dim), 0);
    tmp_n.assign(std::size_t(dim), 0);
    // ldl.cnt is (re)bound to this->cnt at the start of every qp_solve, so
    // that copies/moves of a Workspace never keep a dangling pointer.
  }
  // workspace.hpp:330-377
  void cleanup(bool box)
  {
    (void)box;
    H_scaled.set_zero();
    A_scaled.set_zero();
    C_scaled.set_zero();
    auto z = [](Vec& v) { std::fill(v.begin(), v.end(), 0.); };
    z(g_scaled); z(b_scaled); z(u_scaled); z(l_scaled);
    z(Hdx); z(Cdx); z(Adx); z(active_part_z); z(dw_aug); z(rhs); z(err);
    alpha = 1.;
    z(dual_residual_scaled); z(primal_residual_in_scaled_up);
    z(up_plus_alphaCdx); z(low_plus_alphaCdx); z(CTz);
    z(x_prev); z(y_prev); z(z_prev);
    for (std::size_t i = 0; i < current_bijection_map.size(); ++i) {
      current_bijection_map[i] = isize(i);
      new_bijection_map[i] = isize(i);
      active_inequalities[i] = 0;
    }
    constraints_changed = false;
    dirty = false;
    refactorize = false;
    proximal_parameter_update = false;
    is_initialized = false;
    n_c = 0;
  }
};

// ---------------------------------------------------------------------------
// Ruiz equilibration, ruiz.hpp:31-311 (Symmetry::general) and :403-512.
// ---------------------------------------------------------------------------
inline double
ruiz_scale_qp_in_place(Ruiz& ruiz, Workspace& w, double epsilon, isize max_iter, bool for_infeasible, int hessian_type, bool box)
{
  const double machine_eps = std::numeric_limits<double>::epsilon();
  double c = 1;
  Mat& H = w.H_scaled;
  Mat& A = w.A_scaled;
  Mat& C = w.C_scaled;
  isize n = H.rows, n_eq = A.rows, n_in = C.rows;
  isize ncons = n_in + (box ? n : 0);
  if (box) {
    std::fill(w.i_scaled.begin(), w.i_scaled.end(), 1.);
  }
  double gamma = 1;
  Vec delta(std::size_t(n + n_eq + ncons), 0.);
  Vec colmax(static_cast<std::size_t>(n));
  isize iter = 1;
  auto err_delta = [&]() {
    double m = 0;
    for (double d : delta) {
      m = std::max(m, std::fabs(1 - d));
    }
    return m;
  };
  while (err_delta() > epsilon) {
    if (iter == max_iter) {
      break;
    } else {
      ++iter;
    }
    // column infinity norms
    for (isize k = 0; k < n; ++k) {
      colmax[std::size_t(k)] = 0;
    }
    if (hessian_type == HESSIAN_DENSE) {
      for (isize i = 0; i < n; ++i) {
        const double* r = H.row(i);
        for (isize k = 0; k < n; ++k) {
          colmax[std::size_t(k)] = std::max(colmax[std::size_t(k)], std::fabs(r[k]));
        }
      }
    } else if (hessian_type == HESSIAN_DIAGONAL) {
      for (isize k = 0; k < n; ++k) {
        colmax[std::size_t(k)] = std::fabs(H(k, k));
      }
    }
    for (isize i = 0; i < n_eq; ++i) {
      const double* r = A.row(i);
      for (isize k = 0; k < n; ++k) {
        colmax[std::size_t(k)] = std::max(colmax[std::size_t(k)], std::fabs(r[k]));
      }
    }
    for (isize i = 0; i < n_in; ++i) {
      const double* r = C.row(i);
      for (isize k = 0; k < n; ++k) {
        colmax[std::size_t(k)] = std::max(colmax[std::size_t(k)], std::fabs(r[k]));
      }
    }
    for (isize k = 0; k < n; ++k) {
      double m = colmax[std::size_t(k)];
      if (box) {
        m = std::max(m, w.i_scaled[std::size_t(k)]);
      }
      double aux = std::sqrt(m);
      delta[std::size_t(k)] = aux == 0 ? 1. : 1. / (aux + machine_eps);
    }
    if (for_infeasible) {
      for (isize k = n; k < n + n_eq + ncons; ++k) {
        delta[std::size_t(k)] = 1.;
      }
    } else {
      for (isize k = 0; k < n_eq; ++k) {
        double aux = std::sqrt(infty_norm(A.row(k), n));
        delta[std::size_t(n + k)] = aux == 0 ? 1. : 1. / (aux + machine_eps);
      }
      for (isize k = 0; k < n_in; ++k) {
        double aux = std::sqrt(infty_norm(C.row(k), n));
        delta[std::size_t(n + n_eq + k)] = aux == 0 ? 1. : 1. / (aux + machine_eps);
      }
      if (box) {
        for (isize k = 0; k < n; ++k) {
          delta[std::size_t(n + n_eq + n_in + k)] = 1. / std::sqrt(w.i_scaled[std::size_t(k)] + machine_eps);
        }
      }
    }
    // normalise A, C
    for (isize i = 0; i < n_eq; ++i) {
      double* r = A.row(i);
      double di = delta[std::size_t(n + i)];
      for (isize k = 0; k < n; ++k) {
        r[k] = di * r[k] * delta[std::size_t(k)];
      }
    }
    for (isize i = 0; i < n_in; ++i) {
      double* r = C.row(i);
      double di = delta[std::size_t(n + n_eq + i)];
      for (isize k = 0; k < n; ++k) {
        r[k] = di * r[k] * delta[std::size_t(k)];
      }
    }
    if (box) {
      for (isize k = 0; k < n; ++k) {
        double dt = delta[std::size_t(n + n_eq + n_in + k)];
        w.i_scaled[std::size_t(k)] *= delta[std::size_t(k)];
        w.i_scaled[std::size_t(k)] *= dt;
        w.u_box_scaled[std::size_t(k)] *= dt;
        w.l_box_scaled[std::size_t(k)] *= dt;
      }
    }
    for (isize k = 0; k < n; ++k) {
      w.g_scaled[std::size_t(k)] *= delta[std::size_t(k)];
    }
    for (isize k = 0; k < n_eq; ++k) {
      w.b_scaled[std::size_t(k)] *= delta[std::size_t(n + k)];
    }
    for (isize k = 0; k < n_in; ++k) {
      w.u_scaled[std::size_t(k)] *= delta[std::size_t(n + n_eq + k)];
      w.l_scaled[std::size_t(k)] *= delta[std::size_t(n + n_eq + k)];
    }
    // normalise H
    if (hessian_type == HESSIAN_DENSE) {
      double colsum = 0;
      for (isize k = 0; k < n; ++k) {
        colmax[std::size_t(k)] = 0;
      }
      for (isize i = 0; i < n; ++i) {
        double* r = H.row(i);
        double di = delta[std::size_t(i)];
        for (isize k = 0; k < n; ++k) {
          r[k] = di * r[k] * delta[std::size_t(k)];
          colmax[std::size_t(k)] = std::max(colmax[std::size_t(k)], std::fabs(r[k]));
        }
      }
      for (isize k = 0; k < n; ++k) {
        colsum += colmax[std::size_t(k)];
      }
      gamma = 1 / std::max(1., colsum / double(n));
      // quirk: H is NOT multiplied by gamma here.
    } else if (hessian_type == HESSIAN_DIAGONAL) {
      double dmax = 0;
      for (isize k = 0; k < n; ++k) {
        H(k, k) *= delta[std::size_t(k)];
        H(k, k) *= delta[std::size_t(k)];
        dmax = std::max(dmax, std::fabs(H(k, k)));
      }
      gamma = 1 / std::max(1., dmax / double(n));
      for (double& v : H.a) {
        v *= gamma;
      }
    }
    for (isize k = 0; k < n; ++k) {
      w.g_scaled[std::size_t(k)] *= gamma;
    }
    for (std::size_t k = 0; k < delta.size(); ++k) {
      ruiz.delta[k] *= delta[k];
    }
    c *= gamma;
  }
  return c;
}

// ruiz.hpp:403-512
inline void
ruiz_scale_qp(Ruiz& ruiz, Workspace& w, bool execute, bool for_infeasible, isize max_iter, double epsilon, int hessian_type, bool box)
{
  if (execute) {
    std::fill(ruiz.delta.begin(), ruiz.delta.end(), 1.);
    ruiz.c = ruiz_scale_qp_in_place(ruiz, w, epsilon, max_iter, for_infeasible, hessian_type, box);
    return;
  }
  Mat& H = w.H_scaled;
  Mat& A = w.A_scaled;
  Mat& C = w.C_scaled;
  isize n = H.rows, n_eq = A.rows, n_in = C.rows;
  const Vec& delta = ruiz.delta;
  for (isize i = 0; i < n_eq; ++i) {
    double* r = A.row(i);
    double di = delta[std::size_t(n + i)];
    for (isize k = 0; k < n; ++k) {
      r[k] = di * r[k] * delta[std::size_t(k)];
    }
  }
  for (isize i = 0; i < n_in; ++i) {
    double* r = C.row(i);
    double di = delta[std::size_t(n + n_eq + i)];
    for (isize k = 0; k < n; ++k) {
      r[k] = di * r[k] * delta[std::size_t(k)];
    }
  }
  if (hessian_type == HESSIAN_DENSE) {
    for (isize i = 0; i < n; ++i) {
      double* r = H.row(i);
      double di = delta[std::size_t(i)];
      for (isize k = 0; k < n; ++k) {
        r[k] = di * r[k] * delta[std::size_t(k)];
      }
    }
  } else if (hessian_type == HESSIAN_DIAGONAL) {
    for (isize k = 0; k < n; ++k) {
      H(k, k) *= delta[std::size_t(k)];
      H(k, k) *= delta[std::size_t(k)];
    }
  }
  for (isize k = 0; k < n; ++k) {
    w.g_scaled[std::size_t(k)] *= delta[std::size_t(k)];
  }
  for (isize k = 0; k < n_eq; ++k) {
    w.b_scaled[std::size_t(k)] *= delta[std::size_t(n + k)];
  }
  for (isize k = 0; k < n_in; ++k) {
    w.l_scaled[std::size_t(k)] *= delta[std::size_t(n + n_eq + k)];
    w.u_scaled[std::size_t(k)] *= delta[std::size_t(n + n_eq + k)];
  }
  if (box) {
    isize off = isize(delta.size()) - n;
    for (isize k = 0; k < n; ++k) {
      w.u_box_scaled[std::size_t(k)] *= delta[std::size_t(off + k)];
      w.l_box_scaled[std::size_t(k)] *= delta[std::size_t(off + k)];
      w.i_scaled[std::size_t(k)] *= delta[std::size_t(off + k)];
      w.i_scaled[std::size_t(k)] *= delta[std::size_t(k)];
    }
  }
  for (isize k = 0; k < n; ++k) {
    w.g_scaled[std::size_t(k)] *= ruiz.c;
  }
  for (double& v : H.a) {
    v *= ruiz.c;
  }
}

struct QP; // fwd
void qp_solve(QP& qp);

// wrapper.hpp:82-113
inline int
dense_backend_choice(int backend, isize dim, isize n_eq, isize n_in, bool box)
{
  if (backend != BACKEND_AUTOMATIC) {
    return backend;
  }
  isize ncons = n_in + (box ? dim : 0);
  double threshold = 1.5, frequence = 0.2;
  double d = double(dim);
  double pd = 0.5 * std::pow(double(n_eq) / d, 2) + 0.17 * (std::pow(double(n_eq) / d, 3) + std::pow(double(ncons) / d, 3)) +
              frequence * std::pow(double(n_eq + ncons) / d, 2) / d;
  double p = threshold * ((0.5 * double(n_eq) + double(ncons)) / d + frequence / d);
  return pd > p ? BACKEND_PRIMAL_LDLT : BACKEND_PRIMAL_DUAL_LDLT;
}

// Optional inputs of init/update: nullptr == nullopt.
struct QPData
{
  const double* H = nullptr; // row-major dim x dim
  const double* g = nullptr;
  const double* A = nullptr; // row-major n_eq x dim
  const double* b = nullptr;
  const double* C = nullptr; // row-major n_in x dim
  const double* l = nullptr;
  const double* u = nullptr;
  const double* l_box = nullptr;
  const double* u_box = nullptr;
};

// wrapper.hpp:115-963
struct QP
{
  int dense_backend;
  bool box_constraints;
  int hessian_type;
  Results results;
  Settings settings;
  Model model;
  Workspace work;
  Ruiz ruiz;

  QP(isize dim, isize n_eq, isize n_in, bool box = false, int hessian = HESSIAN_DENSE, int backend = BACKEND_PRIMAL_DUAL_LDLT)
    : dense_backend(dense_backend_choice(backend, dim, n_eq, n_in, box))
    , box_constraints(box)
    , hessian_type(hessian)
    , results(dim, n_eq, n_in, box, dense_backend)
    , settings(dense_backend)
    , model(dim, n_eq, n_in, box)
    , work(dim, n_eq, n_in, box, dense_backend)
    , ruiz(dim, n_eq, n_in, box)
  {
  }
  QP(const QP& o) = default;
  QP(QP&& o) = default;

  isize n_constraints() const { return model.n_in + (box_constraints ? model.dim : 0); }

  // helpers.hpp:678-705
  void update_proximal_parameters(const double* rho, const double* mu_eq, const double* mu_in)
  {
    if (rho) {
      settings.default_rho = *rho;
      results.info.rho = *rho;
      work.proximal_parameter_update = true;
    }
    if (mu_eq) {
      settings.default_mu_eq = *mu_eq;
      results.info.mu_eq = *mu_eq;
      results.info.mu_eq_inv = 1. / results.info.mu_eq;
      work.proximal_parameter_update = true;
    }
    if (mu_in) {
      settings.default_mu_in = *mu_in;
      results.info.mu_in = *mu_in;
      results.info.mu_in_inv = 1. / results.info.mu_in;
      work.proximal_parameter_update = true;
    }
  }
  // helpers.hpp:174-189
  void update_default_rho_with_minimal_Hessian_eigen_value(const double* manual)
  {
    if (manual) {
      settings.default_H_eigenvalue_estimate = *manual;
      results.info.minimal_H_eigenvalue_estimate = settings.default_H_eigenvalue_estimate;
    }
    settings.default_rho += std::fabs(results.info.minimal_H_eigenvalue_estimate);
    results.info.rho = settings.default_rho;
  }

  // helpers.hpp:500-667
  void setup(const QPData& d, int preconditioner_status)
  {
    bool box = box_constraints;
    switch (settings.initial_guess) {
      case EQUALITY_CONSTRAINED_INITIAL_GUESS:
      case NO_INITIAL_GUESS:
      case WARM_START:
        if (work.proximal_parameter_update) {
          results.cleanup_all_except_prox_parameters();
        } else {
          results.cleanup(&settings);
        }
        work.cleanup(box);
        break;
      case COLD_START_WITH_PREVIOUS_RESULT:
        if (work.proximal_parameter_update) {
          results.cleanup_statistics();
        } else {
          results.cold_start(&settings);
        }
        work.cleanup(box);
        break;
      case WARM_START_WITH_PREVIOUS_RESULT:
        if (work.refactorize || work.proximal_parameter_update) {
          work.cleanup(box);
          work.refactorize = true;
        }
        results.cleanup_statistics();
        break;
    }
    isize n = model.dim, n_eq = model.n_eq, n_in = model.n_in;
    if (d.H) std::copy(d.H, d.H + n * n, model.H.a.begin());
    if (d.g) std::copy(d.g, d.g + n, model.g.begin());
    if (d.A) std::copy(d.A, d.A + n_eq * n, model.A.a.begin());
    if (d.b) std::copy(d.b, d.b + n_eq, model.b.begin());
    if (d.C) std::copy(d.C, d.C + n_in * n, model.C.a.begin());
    if (d.u) std::copy(d.u, d.u + n_in, model.u.begin());
    if (d.l) std::copy(d.l, d.l + n_in, model.l.begin());
    if (d.u_box) std::copy(d.u_box, d.u_box + n, model.u_box.begin());
    if (d.l_box) std::copy(d.l_box, d.l_box + n, model.l_box.begin());
    copy_model_to_scaled();
    for (isize i = 0; i < n_in; ++i) {
      work.u_scaled[std::size_t(i)] = model.u[std::size_t(i)] <= 1e20 ? model.u[std::size_t(i)] : 1e20;
      work.l_scaled[std::size_t(i)] = model.l[std::size_t(i)] >= -1e20 ? model.l[std::size_t(i)] : -1e20;
    }
    if (box) {
      for (isize i = 0; i < n; ++i) {
        work.u_box_scaled[std::size_t(i)] = model.u_box[std::size_t(i)] <= 1e20 ? model.u_box[std::size_t(i)] : 1e20;
        work.l_box_scaled[std::size_t(i)] = model.l_box[std::size_t(i)] >= -1e20 ? model.l_box[std::size_t(i)] : -1e20;
      }
    }
    work.dual_feasibility_rhs_2 = infty_norm(model.g);
    setup_equilibration(preconditioner_status == PRECOND_EXECUTE);
  }
  void copy_model_to_scaled()
  {
    if (hessian_type != HESSIAN_ZERO) {
      work.H_scaled.a = model.H.a;
    }
    work.g_scaled = model.g;
    work.A_scaled.a = model.A.a;
    work.b_scaled = model.b;
    work.C_scaled.a = model.C.a;
  }
  // helpers.hpp:298-329
  void setup_equilibration(bool execute)
  {
    ruiz_scale_qp(ruiz, work, execute, settings.primal_infeasibility_solving, settings.preconditioner_max_iter, settings.preconditioner_accuracy, hessian_type, box_constraints);
    work.correction_guess_rhs_g = infty_norm(work.g_scaled);
  }

  void check_sizes(const QPData&) {}

  // wrapper.hpp:354-498 and :520-703. Sizes are validated by the C API layer
  // (raw pointers carry no size), which raises what the reference throws.
  void init(const QPData& d, bool compute_preconditioner = true, const double* rho = nullptr, const double* mu_eq = nullptr, const double* mu_in = nullptr, const double* manual_minimal_H_eigenvalue = nullptr)
  {
    if (!box_constraints && (d.l_box || d.u_box)) {
      throw std::invalid_argument("wrong model setup: the QP object is designed without box constraints, but is initialized with lower or upper box inequalities.");
    }
    settings.compute_preconditioner = compute_preconditioner;
    work.refactorize = settings.initial_guess == WARM_START_WITH_PREVIOUS_RESULT;
    work.proximal_parameter_update = false;
    update_proximal_parameters(rho, mu_eq, mu_in);
    update_default_rho_with_minimal_Hessian_eigen_value(manual_minimal_H_eigenvalue);
    setup(d, compute_preconditioner ? PRECOND_EXECUTE : PRECOND_IDENTITY);
    work.is_initialized = true;
  }

  // wrapper.hpp:723-918 + helpers.hpp:372-480
  void update(const QPData& d, bool update_preconditioner = false, const double* rho = nullptr, const double* mu_eq = nullptr, const double* mu_in = nullptr, const double* manual_minimal_H_eigenvalue = nullptr)
  {
    if (!box_constraints && (d.l_box || d.u_box)) {
      throw std::invalid_argument("wrong model setup: the QP object is designed without box constraints, but the update includes lower or upper box inequalities.");
    }
    settings.update_preconditioner = update_preconditioner;
    if (!work.is_initialized) {
      // wrapper.hpp:743-746 (manual eigenvalue is not forwarded by the reference)
      init(d, update_preconditioner, rho, mu_eq, mu_in, nullptr);
      return;
    }
    work.refactorize = false;
    work.proximal_parameter_update = false;
    int status = update_preconditioner ? PRECOND_EXECUTE : PRECOND_KEEP;
    isize n = model.dim, n_eq = model.n_eq, n_in = model.n_in;
    if (d.g) std::copy(d.g, d.g + n, model.g.begin());
    if (d.b) std::copy(d.b, d.b + n_eq, model.b.begin());
    if (d.u) std::copy(d.u, d.u + n_in, model.u.begin());
    if (d.l) std::copy(d.l, d.l + n_in, model.l.begin());
    if (d.u_box && box_constraints) std::copy(d.u_box, d.u_box + n, model.u_box.begin());
    if (d.l_box && box_constraints) std::copy(d.l_box, d.l_box + n, model.l_box.begin());
    if (d.H || d.A || d.C) {
      work.refactorize = true;
    }
    if (d.H) std::copy(d.H, d.H + n * n, model.H.a.begin());
    if (d.A) std::copy(d.A, d.A + n_eq * n, model.A.a.begin());
    if (d.C) std::copy(d.C, d.C + n_in * n, model.C.a.begin());
    update_proximal_parameters(rho, mu_eq, mu_in);
    update_default_rho_with_minimal_Hessian_eigen_value(manual_minimal_H_eigenvalue);
    QPData none;
    setup(none, status);
  }

  // helpers.hpp:715-763
  void warm_start(const double* x, const double* y, const double* z)
  {
    if (!x && !y && !z) {
      return;
    }
    settings.initial_guess = WARM_START;
    if (x) std::copy(x, x + model.dim, results.x.begin());
    if (y) std::copy(y, y + model.n_eq, results.y.begin());
    // the reference checks z against n_in only (helpers.hpp:744-750) but
    // assigns the whole vector; with box constraints callers pass n_in + dim.
    if (z) std::copy(z, z + n_constraints(), results.z.begin());
  }
  void solve() { qp_solve(*this); }
  void solve(const double* x, const double* y, const double* z)
  {
    warm_start(x, y, z);
    qp_solve(*this);
  }
  void cleanup()
  {
    results.cleanup(&settings);
    work.cleanup(box_constraints);
  }
};

} // namespace oracle

#include "proxqp_solver.hpp"
#include "proxqp_backward.hpp"
