// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE. See proxqp_oracle.hpp.
//
// Restatement of dense/solver.hpp, dense/linesearch.hpp, dense/utils.hpp and
// the factorisation helpers of dense/helpers.hpp (file:line cited per
// function, under /root/reference/include/proxsuite/proxqp).
#pragma once
#include <cstdio>
#include <cstdlib>

namespace oracle {

// helpers.hpp:239-285
inline void
setup_factorization(QP& qp)
{
  Workspace& w = qp.work;
  const Model& m = qp.model;
  isize n = m.dim, n_eq = m.n_eq;
  Mat& kkt = w.kkt;
  isize ld = kkt.cols;
  for (isize i = 0; i < n; ++i) {
    for (isize j = 0; j < n; ++j) {
      kkt(i, j) = qp.hessian_type == HESSIAN_ZERO ? 0. : w.H_scaled(i, j);
    }
    kkt(i, i) += qp.results.info.rho;
  }
  if (qp.dense_backend == BACKEND_PRIMAL_DUAL_LDLT) {
    for (isize i = 0; i < n_eq; ++i) {
      for (isize j = 0; j < n; ++j) {
        kkt(n + i, j) = w.A_scaled(i, j);
        kkt(j, n + i) = w.A_scaled(i, j);
      }
      for (isize j = 0; j < n_eq; ++j) {
        kkt(n + i, n + j) = 0.;
      }
      kkt(n + i, n + i) = -qp.results.info.mu_eq;
    }
    // ldl.factorize(kkt.transpose()): the lower triangle of kkt^T (col-major
    // view of the row-major storage) is the upper triangle of kkt; kkt is
    // symmetric so reading the lower triangle of kkt is equivalent.
    w.ldl.factorize(kkt.a.data(), ld, n + n_eq);
  } else {
    // PrimalLDLT: kkt += mu_eq_inv * A^T A
    for (isize k = 0; k < n_eq; ++k) {
      const double* r = w.A_scaled.row(k);
      for (isize i = 0; i < n; ++i) {
        double f = qp.results.info.mu_eq_inv * r[i];
        if (f == 0) {
          continue;
        }
        for (isize j = 0; j < n; ++j) {
          kkt(i, j) += f * r[j];
        }
      }
    }
    w.ldl.factorize(kkt.a.data(), ld, n);
  }
}

// Fill column `col` (length rows) of the new-columns buffer for constraint i.
inline void
fill_constraint_col(QP& qp, double* col, isize i)
{
  isize n = qp.model.dim, n_in = qp.model.n_in;
  if (i >= n_in) {
    for (isize k = 0; k < n; ++k) {
      col[k] = 0;
    }
    col[i - n_in] = qp.work.i_scaled[std::size_t(i - n_in)];
  } else {
    const double* r = qp.work.C_scaled.row(i);
    for (isize k = 0; k < n; ++k) {
      col[k] = r[k];
    }
  }
}

// solver.hpp:38-115
inline void
refactorize(QP& qp, double rho_new)
{
  Workspace& w = qp.work;
  Results& res = qp.results;
  const Model& m = qp.model;
  if (!w.constraints_changed && rho_new == res.info.rho) {
    return;
  }
  isize n = m.dim, n_eq = m.n_eq, n_in = m.n_in;
  isize ncons = qp.n_constraints();
  if (qp.dense_backend == BACKEND_PRIMAL_DUAL_LDLT) {
    for (isize i = 0; i < n; ++i) {
      w.kkt(i, i) += rho_new - res.info.rho;
    }
    for (isize i = 0; i < n_eq; ++i) {
      w.kkt(n + i, n + i) = -res.info.mu_eq;
    }
    w.ldl.factorize(w.kkt.a.data(), w.kkt.cols, n + n_eq);
    isize n_c = w.n_c;
    isize rows = n + n_eq + n_c;
    w.new_cols.assign(std::size_t(rows * std::max<isize>(1, n_c)), 0.);
    for (isize i = 0; i < ncons; ++i) {
      isize j = w.current_bijection_map[std::size_t(i)];
      if (j < n_c) {
        double* col = &w.new_cols[std::size_t(j * rows)];
        fill_constraint_col(qp, col, i);
        for (isize k = n; k < rows; ++k) {
          col[k] = 0;
        }
        col[n + n_eq + j] = -res.info.mu_in;
      }
    }
    (void)n_in;
    w.ldl.insert_block_at(n + n_eq, w.new_cols.data(), rows, n_c);
  } else {
    for (isize i = 0; i < n; ++i) {
      for (isize j = 0; j < n; ++j) {
        w.kkt(i, j) = qp.hessian_type == HESSIAN_ZERO ? 0. : w.H_scaled(i, j);
      }
    }
    for (isize k = 0; k < n_eq; ++k) {
      const double* r = w.A_scaled.row(k);
      for (isize i = 0; i < n; ++i) {
        double f = res.info.mu_eq_inv * r[i];
        for (isize j = 0; j < n; ++j) {
          w.kkt(i, j) += f * r[j];
        }
      }
    }
    for (isize i = 0; i < n; ++i) {
      w.kkt(i, i) += res.info.rho;
    }
    for (isize i = 0; i < ncons; ++i) {
      if (w.active_inequalities[std::size_t(i)]) {
        if (i >= n_in) {
          double s = w.i_scaled[std::size_t(i - n_in)];
          w.kkt(i - n_in, i - n_in) += s * s * res.info.mu_in_inv;
        } else {
          const double* r = w.C_scaled.row(i);
          for (isize a = 0; a < n; ++a) {
            double f = r[a] * res.info.mu_in_inv;
            for (isize b = 0; b < n; ++b) {
              w.kkt(a, b) += f * r[b];
            }
          }
        }
      }
    }
    w.ldl.factorize(w.kkt.a.data(), w.kkt.cols, n);
  }
  w.constraints_changed = false;
}

// solver.hpp:128-232
inline void
mu_update(QP& qp, double mu_eq_new, double mu_in_new)
{
  Workspace& w = qp.work;
  Results& res = qp.results;
  isize n = qp.model.dim, n_eq = qp.model.n_eq, n_c = w.n_c;
  isize ncons = qp.n_constraints();
  if (n_eq + n_c == 0) {
    return;
  }
  if (qp.dense_backend == BACKEND_PRIMAL_DUAL_LDLT) {
    Vec alpha(std::size_t(n_eq + n_c));
    std::vector<isize> indices(std::size_t(n_eq + n_c));
    for (isize k = 0; k < n_eq; ++k) {
      alpha[std::size_t(k)] = res.info.mu_eq - mu_eq_new;
      indices[std::size_t(k)] = n + k;
    }
    for (isize k = 0; k < n_c; ++k) {
      alpha[std::size_t(n_eq + k)] = res.info.mu_in - mu_in_new;
      indices[std::size_t(n_eq + k)] = n + n_eq + k;
    }
    w.ldl.diagonal_update_clobber_indices(indices.data(), n_eq + n_c, alpha.data());
  } else {
    {
      w.new_cols.assign(std::size_t(n * std::max<isize>(1, n_c)), 0.);
      Vec alpha(std::size_t(std::max<isize>(1, n_c)), 1. / mu_in_new - res.info.mu_in_inv);
      for (isize i = 0; i < ncons; ++i) {
        isize j = w.current_bijection_map[std::size_t(i)];
        if (j < n_c) {
          fill_constraint_col(qp, &w.new_cols[std::size_t(j * n)], i);
        }
      }
      w.ldl.rank_r_update(w.new_cols.data(), n, n_c, alpha.data());
    }
    {
      w.new_cols.assign(std::size_t(n * std::max<isize>(1, n_eq)), 0.);
      Vec alpha(std::size_t(std::max<isize>(1, n_eq)), 1. / mu_eq_new - res.info.mu_eq_inv);
      for (isize k = 0; k < n_eq; ++k) {
        const double* r = w.A_scaled.row(k);
        for (isize a = 0; a < n; ++a) {
          w.new_cols[std::size_t(k * n + a)] = r[a];
        }
      }
      w.ldl.rank_r_update(w.new_cols.data(), n, n_eq, alpha.data());
    }
  }
  w.constraints_changed = true;
}

// solver.hpp:243-318
inline void
iterative_residual(QP& qp, isize inner_pb_dim)
{
  Workspace& w = qp.work;
  const Results& res = qp.results;
  isize n = qp.model.dim, n_eq = qp.model.n_eq, n_in = qp.model.n_in;
  isize ncons = qp.n_constraints();
  double* err = w.err.data();
  const double* dw = w.dw_aug.data();
  Vec& Hdx = w.Hdx;
  Vec& Adx = w.Adx;
  Vec& ATdy = w.CTz;
  for (isize i = 0; i < inner_pb_dim; ++i) {
    err[i] = w.rhs[std::size_t(i)];
  }
  switch (qp.hessian_type) {
    case HESSIAN_ZERO:
      break;
    case HESSIAN_DENSE:
      symv_lower(w.H_scaled, dw, Hdx.data());
      for (isize i = 0; i < n; ++i) {
        err[i] -= Hdx[std::size_t(i)];
      }
      break;
    case HESSIAN_DIAGONAL:
      for (isize i = 0; i < n; ++i) {
        Hdx[std::size_t(i)] = w.H_scaled(i, i) * dw[i];
        err[i] -= Hdx[std::size_t(i)];
      }
      break;
  }
  for (isize i = 0; i < n; ++i) {
    err[i] -= res.info.rho * dw[i];
  }
  gemv_t(w.A_scaled, dw + n, ATdy.data(), false);
  for (isize i = 0; i < n; ++i) {
    err[i] -= ATdy[std::size_t(i)];
  }
  if (ncons > n_in) {
    for (isize i = 0; i < n; ++i) {
      w.active_part_z[std::size_t(n_in + i)] = dw[i] * w.i_scaled[std::size_t(i)];
    }
  }
  for (isize i = 0; i < ncons; ++i) {
    isize j = w.current_bijection_map[std::size_t(i)];
    if (j < w.n_c) {
      double dzj = dw[n + n_eq + j];
      if (i >= n_in) {
        err[i - n_in] -= dzj * w.i_scaled[std::size_t(i - n_in)];
        err[n + n_eq + j] -= (w.active_part_z[std::size_t(i)] - dzj * res.info.mu_in);
      } else {
        const double* r = w.C_scaled.row(i);
        for (isize k = 0; k < n; ++k) {
          err[k] -= dzj * r[k];
        }
        err[n + n_eq + j] -= (dot(r, dw, n) - dzj * res.info.mu_in);
      }
    }
  }
  gemv(w.A_scaled, dw, Adx.data());
  for (isize i = 0; i < n_eq; ++i) {
    err[n + i] -= Adx[std::size_t(i)];
    err[n + i] += dw[n + i] * res.info.mu_eq;
  }
  w.cnt.n_resid += 1;
  w.cnt.resid_nc += double(w.n_c);
}

// solver.hpp:320-392
inline void
solve_linear_system(QP& qp, Vec& dwv, isize inner_pb_dim)
{
  Workspace& w = qp.work;
  const Results& res = qp.results;
  double* dw = dwv.data();
  isize n = qp.model.dim, n_eq = qp.model.n_eq, n_in = qp.model.n_in;
  isize ncons = qp.n_constraints();
  if (qp.dense_backend == BACKEND_PRIMAL_DUAL_LDLT) {
    w.ldl.solve_in_place(dw, inner_pb_dim);
    return;
  }
  // PrimalLDLT
  {
    // dx rhs += mu_eq_inv * A^T dw_y
    for (isize k = 0; k < n_eq; ++k) {
      double f = res.info.mu_eq_inv * dw[n + k];
      const double* r = w.A_scaled.row(k);
      for (isize a = 0; a < n; ++a) {
        dw[a] += f * r[a];
      }
    }
    for (isize i = 0; i < ncons; ++i) {
      isize j = w.current_bijection_map[std::size_t(i)];
      if (j < w.n_c) {
        if (i >= n_in) {
          dw[i - n_in] += dw[j + n + n_eq] * w.i_scaled[std::size_t(i - n_in)];
        } else {
          const double* r = w.C_scaled.row(i);
          double f = dw[j + n + n_eq];
          for (isize a = 0; a < n; ++a) {
            dw[a] += f * r[a];
          }
        }
      }
    }
    w.ldl.solve_in_place(dw, n);
    for (isize k = 0; k < n_eq; ++k) {
      dw[n + k] -= res.info.mu_eq_inv * dw[n + k];
      dw[n + k] += res.info.mu_eq_inv * dot(w.A_scaled.row(k), dw, n);
    }
    for (isize i = 0; i < ncons; ++i) {
      isize j = w.current_bijection_map[std::size_t(i)];
      if (j < w.n_c) {
        if (i >= n_in) {
          dw[j + n + n_eq] -= res.info.mu_in_inv * dw[j + n + n_eq];
          dw[j + n + n_eq] += res.info.mu_in_inv * dw[i - n_in];
        } else {
          dw[j + n + n_eq] -= res.info.mu_in_inv * dw[j + n + n_eq];
          dw[j + n + n_eq] += res.info.mu_in_inv * dot(w.C_scaled.row(i), dw, n);
        }
      }
    }
  }
}

// solver.hpp:406-541
inline void
iterative_solve_with_permut_fact(QP& qp, double eps, isize inner_pb_dim)
{
  Workspace& w = qp.work;
  const Settings& s = qp.settings;
  std::fill(w.err.begin(), w.err.end(), 0.);
  isize it = 0, it_stability = 0;
  auto err_norm = [&]() { return infty_norm(w.err.data(), inner_pb_dim); };
  for (isize i = 0; i < inner_pb_dim; ++i) {
    w.dw_aug[std::size_t(i)] = w.rhs[std::size_t(i)];
  }
  solve_linear_system(qp, w.dw_aug, inner_pb_dim);
  iterative_residual(qp, inner_pb_dim);
  ++it;
  double preverr = err_norm();
  while (err_norm() >= eps) {
    if (it >= s.nb_iterative_refinement) {
      break;
    }
    ++it;
    solve_linear_system(qp, w.err, inner_pb_dim);
    for (isize i = 0; i < inner_pb_dim; ++i) {
      w.dw_aug[std::size_t(i)] += w.err[std::size_t(i)];
      w.err[std::size_t(i)] = 0;
    }
    iterative_residual(qp, inner_pb_dim);
    if (err_norm() > preverr) {
      it_stability += 1;
    } else {
      it_stability = 0;
    }
    if (it_stability == 2) {
      break;
    }
    preverr = err_norm();
  }
  if (err_norm() >= std::max(eps, s.eps_refact)) {
    refactorize(qp, qp.results.info.rho);
    it = 0;
    it_stability = 0;
    for (isize i = 0; i < inner_pb_dim; ++i) {
      w.dw_aug[std::size_t(i)] = w.rhs[std::size_t(i)];
    }
    solve_linear_system(qp, w.dw_aug, inner_pb_dim);
    iterative_residual(qp, inner_pb_dim);
    preverr = err_norm();
    ++it;
    while (err_norm() >= eps) {
      if (it >= s.nb_iterative_refinement) {
        break;
      }
      ++it;
      solve_linear_system(qp, w.err, inner_pb_dim);
      for (isize i = 0; i < inner_pb_dim; ++i) {
        w.dw_aug[std::size_t(i)] += w.err[std::size_t(i)];
        w.err[std::size_t(i)] = 0;
      }
      iterative_residual(qp, inner_pb_dim);
      if (err_norm() > preverr) {
        it_stability += 1;
      } else {
        it_stability = 0;
      }
      if (it_stability == 2) {
        break;
      }
      preverr = err_norm();
    }
  }
  qp.results.info.iterative_residual = err_norm();
  for (isize i = 0; i < inner_pb_dim; ++i) {
    w.rhs[std::size_t(i)] = 0;
  }
}

// helpers.hpp:199-228
inline void
compute_equality_constrained_initial_guess(QP& qp)
{
  Workspace& w = qp.work;
  isize n = qp.model.dim, n_eq = qp.model.n_eq;
  std::fill(w.rhs.begin(), w.rhs.end(), 0.);
  for (isize i = 0; i < n; ++i) {
    w.rhs[std::size_t(i)] = -w.g_scaled[std::size_t(i)];
  }
  for (isize i = 0; i < n_eq; ++i) {
    w.rhs[std::size_t(n + i)] = w.b_scaled[std::size_t(i)];
  }
  iterative_solve_with_permut_fact(qp, 1., n + n_eq);
  for (isize i = 0; i < n; ++i) {
    qp.results.x[std::size_t(i)] = w.dw_aug[std::size_t(i)];
  }
  for (isize i = 0; i < n_eq; ++i) {
    qp.results.y[std::size_t(i)] = w.dw_aug[std::size_t(n + i)];
  }
  std::fill(w.dw_aug.begin(), w.dw_aug.end(), 0.);
  std::fill(w.rhs.begin(), w.rhs.end(), 0.);
}

// utils.hpp:164-252
inline void
global_primal_residual(QP& qp, double& primal_feasibility_lhs, double& primal_feasibility_eq_rhs_0, double& primal_feasibility_in_rhs_0, double& primal_feasibility_eq_lhs, double& primal_feasibility_in_lhs)
{
  Workspace& w = qp.work;
  Results& res = qp.results;
  const Model& m = qp.model;
  const Ruiz& ruiz = qp.ruiz;
  isize n = m.dim, n_eq = m.n_eq, n_in = m.n_in;
  bool box = qp.box_constraints;
  gemv(w.A_scaled, res.x.data(), res.se.data());
  gemv(w.C_scaled, res.x.data(), w.primal_residual_in_scaled_up.data());
  if (box) {
    for (isize i = 0; i < n; ++i) {
      w.primal_residual_in_scaled_up[std::size_t(n_in + i)] = res.x[std::size_t(i)];
    }
    ruiz.unscale_primal(w.primal_residual_in_scaled_up.data() + n_in);
  }
  ruiz.unscale_primal_residual_eq(res.se.data());
  primal_feasibility_eq_rhs_0 = infty_norm(res.se.data(), n_eq);
  ruiz.unscale_primal_residual_in(w.primal_residual_in_scaled_up.data());
  primal_feasibility_in_rhs_0 = infty_norm(w.primal_residual_in_scaled_up.data(), n_in);
  for (isize i = 0; i < n_in; ++i) {
    double v = w.primal_residual_in_scaled_up[std::size_t(i)];
    res.si[std::size_t(i)] = pos_part(v - m.u[std::size_t(i)]) + neg_part(v - m.l[std::size_t(i)]);
  }
  if (box) {
    for (isize i = 0; i < n; ++i) {
      double v = w.primal_residual_in_scaled_up[std::size_t(n_in + i)];
      res.si[std::size_t(n_in + i)] = pos_part(v - m.u_box[std::size_t(i)]) + neg_part(v - m.l_box[std::size_t(i)]);
      w.active_part_z[std::size_t(n_in + i)] = res.x[std::size_t(i)] - res.si[std::size_t(n_in + i)];
    }
    primal_feasibility_in_rhs_0 = std::max(primal_feasibility_in_rhs_0, infty_norm(w.active_part_z.data() + n_in, n));
    primal_feasibility_in_rhs_0 = std::max(primal_feasibility_in_rhs_0, infty_norm(res.x.data(), n));
  }
  for (isize i = 0; i < n_eq; ++i) {
    res.se[std::size_t(i)] -= m.b[std::size_t(i)];
  }
  primal_feasibility_in_lhs = infty_norm(res.si);
  primal_feasibility_eq_lhs = infty_norm(res.se);
  primal_feasibility_lhs = std::max(primal_feasibility_eq_lhs, primal_feasibility_in_lhs);
  if (qp.settings.primal_infeasibility_solving && res.info.status == PROXQP_PRIMAL_INFEASIBLE) {
    gemv_t(m.A, res.se.data(), w.rhs.data(), false);
    gemv_t(m.C, res.si.data(), w.rhs.data(), true);
    primal_feasibility_lhs = infty_norm(w.rhs.data(), n);
  }
  ruiz.scale_primal_residual_eq(res.se.data());
  w.cnt.n_global_res += 0.5;
}

// utils.hpp:269-324 (arguments are modified in place)
inline bool
global_primal_residual_infeasibility(QP& qp, double* ATdy, double* CTdz, double* dy, double* dz)
{
  Workspace& w = qp.work;
  const Model& m = qp.model;
  const Ruiz& ruiz = qp.ruiz;
  isize n = m.dim, n_eq = m.n_eq, n_in = m.n_in;
  isize ncons = qp.n_constraints();
  bool res = infty_norm(dy, n_eq) != 0 || infty_norm(dz, ncons) != 0;
  if (!res) {
    return res;
  }
  ruiz.unscale_dual_residual(ATdy);
  ruiz.unscale_dual_residual(CTdz);
  double lower_bound_1 = dot(dy, w.b_scaled.data(), n_eq);
  for (isize i = 0; i < n_in; ++i) {
    lower_bound_1 += pos_part(dz[i]) * w.u_scaled[std::size_t(i)] - neg_part(dz[i]) * w.l_scaled[std::size_t(i)];
  }
  ruiz.unscale_dual_eq(dy);
  ruiz.unscale_dual_in(dz);
  if (qp.box_constraints) {
    for (isize i = 0; i < n; ++i) {
      lower_bound_1 += pos_part(dz[n_in + i]) * w.u_box_scaled[std::size_t(i)] - neg_part(dz[n_in + i]) * w.l_box_scaled[std::size_t(i)];
    }
    ruiz.unscale_box_dual_in(dz + n_in);
  }
  double upper_bound = qp.settings.eps_primal_inf * std::max(infty_norm(dy, n_eq), infty_norm(dz, ncons));
  double lower_bound_2 = 0;
  for (isize i = 0; i < n; ++i) {
    lower_bound_2 = std::max(lower_bound_2, std::fabs(ATdy[i] + CTdz[i]));
  }
  res = lower_bound_2 <= upper_bound && lower_bound_1 <= -upper_bound;
  return res;
}

// utils.hpp:343-419 (arguments are modified in place)
inline bool
global_dual_residual_infeasibility(QP& qp, double* Adx, double* Cdx, double* Hdx, double* dx)
{
  Workspace& w = qp.work;
  const Model& m = qp.model;
  const Ruiz& ruiz = qp.ruiz;
  isize n = m.dim, n_eq = m.n_eq, n_in = m.n_in;
  ruiz.unscale_dual_residual(Hdx);
  ruiz.unscale_primal_residual_eq(Adx);
  ruiz.unscale_primal_residual_in(Cdx);
  if (qp.box_constraints) {
    ruiz.unscale_box_primal_residual_in(Cdx + n_in);
  }
  double gdx = dot(dx, w.g_scaled.data(), n);
  ruiz.unscale_primal(dx);
  double bound = infty_norm(dx, n) * qp.settings.eps_dual_inf;
  double bound_neg = -bound;
  bool first_cond = infty_norm(Adx, n_eq) <= bound;
  for (isize i = 0; i < n_in; ++i) {
    double Cdx_i = Cdx[i];
    if (w.u_scaled[std::size_t(i)] <= 1e20 && w.l_scaled[std::size_t(i)] >= -1e20) {
      first_cond = first_cond && Cdx_i <= bound && Cdx_i >= bound_neg;
    } else if (w.u_scaled[std::size_t(i)] > 1e20) {
      first_cond = first_cond && Cdx_i >= bound_neg;
    } else if (w.l_scaled[std::size_t(i)] < -1e20) {
      first_cond = first_cond && Cdx_i <= bound;
    }
  }
  if (qp.box_constraints) {
    for (isize i = 0; i < n; ++i) {
      double dx_i = dx[i];
      if (w.u_box_scaled[std::size_t(i)] <= 1e20 && w.l_box_scaled[std::size_t(i)] >= -1e20) {
        first_cond = first_cond && dx_i <= bound && dx_i >= bound_neg;
      } else if (w.u_box_scaled[std::size_t(i)] > 1e20) {
        first_cond = first_cond && dx_i >= bound_neg;
      } else if (w.l_box_scaled[std::size_t(i)] < -1e20) {
        first_cond = first_cond && dx_i <= bound;
      }
    }
  }
  bound *= ruiz.c;
  bound_neg *= ruiz.c;
  bool second_cond_alt1 = infty_norm(Hdx, n) <= bound && gdx <= bound_neg;
  return first_cond && second_cond_alt1 && infty_norm(dx, n) != 0;
}

// utils.hpp:437-587
inline void
global_dual_residual(QP& qp, double& dual_feasibility_lhs, double& dual_feasibility_rhs_0, double& dual_feasibility_rhs_1, double& dual_feasibility_rhs_3, double& rhs_duality_gap, double& duality_gap)
{
  Workspace& w = qp.work;
  Results& res = qp.results;
  const Model& m = qp.model;
  const Ruiz& ruiz = qp.ruiz;
  isize n = m.dim, n_eq = m.n_eq, n_in = m.n_in;
  bool box = qp.box_constraints;
  const double inf_b = infinite_bound();
  w.dual_residual_scaled = w.g_scaled;
  switch (qp.hessian_type) {
    case HESSIAN_ZERO:
      dual_feasibility_rhs_0 = 0;
      break;
    case HESSIAN_DENSE:
      symv_lower(w.H_scaled, res.x.data(), w.CTz.data());
      for (isize i = 0; i < n; ++i) {
        w.dual_residual_scaled[std::size_t(i)] += w.CTz[std::size_t(i)];
      }
      ruiz.unscale_dual_residual(w.CTz.data());
      dual_feasibility_rhs_0 = infty_norm(w.CTz);
      break;
    case HESSIAN_DIAGONAL:
      for (isize i = 0; i < n; ++i) {
        w.CTz[std::size_t(i)] = w.H_scaled(i, i) * res.x[std::size_t(i)];
        w.dual_residual_scaled[std::size_t(i)] += w.CTz[std::size_t(i)];
      }
      ruiz.unscale_dual_residual(w.CTz.data());
      dual_feasibility_rhs_0 = infty_norm(w.CTz);
      break;
  }
  ruiz.unscale_primal(res.x.data());
  duality_gap = dot(m.g.data(), res.x.data(), n);
  rhs_duality_gap = std::fabs(duality_gap);
  if (qp.hessian_type != HESSIAN_ZERO) {
    double xHx = dot(w.CTz.data(), res.x.data(), n);
    duality_gap += xHx;
    rhs_duality_gap = std::max(rhs_duality_gap, std::fabs(xHx));
  }
  ruiz.scale_primal(res.x.data());

  gemv_t(w.A_scaled, res.y.data(), w.CTz.data(), false);
  for (isize i = 0; i < n; ++i) {
    w.dual_residual_scaled[std::size_t(i)] += w.CTz[std::size_t(i)];
  }
  ruiz.unscale_dual_residual(w.CTz.data());
  dual_feasibility_rhs_1 = infty_norm(w.CTz);

  gemv_t(w.C_scaled, res.z.data(), w.CTz.data(), false);
  for (isize i = 0; i < n; ++i) {
    w.dual_residual_scaled[std::size_t(i)] += w.CTz[std::size_t(i)];
  }
  ruiz.unscale_dual_residual(w.CTz.data());
  dual_feasibility_rhs_3 = infty_norm(w.CTz);
  if (box) {
    for (isize i = 0; i < n; ++i) {
      w.CTz[std::size_t(i)] = res.z[std::size_t(n_in + i)] * w.i_scaled[std::size_t(i)];
      w.dual_residual_scaled[std::size_t(i)] += w.CTz[std::size_t(i)];
    }
    ruiz.unscale_dual_residual(w.CTz.data());
    dual_feasibility_rhs_3 = std::max(infty_norm(w.CTz), dual_feasibility_rhs_3);
  }
  ruiz.unscale_dual_residual(w.dual_residual_scaled.data());
  dual_feasibility_lhs = infty_norm(w.dual_residual_scaled);
  ruiz.scale_dual_residual(w.dual_residual_scaled.data());

  ruiz.unscale_dual_eq(res.y.data());
  const double by = dot(m.b.data(), res.y.data(), n_eq);
  rhs_duality_gap = std::max(rhs_duality_gap, std::fabs(by));
  duality_gap += by;
  ruiz.scale_dual_eq(res.y.data());

  ruiz.unscale_dual_in(res.z.data());
  double zu = 0, zl = 0;
  for (isize i = 0; i < n_in; ++i) {
    if (w.active_set_up[std::size_t(i)]) {
      zu += res.z[std::size_t(i)] * std::min(m.u[std::size_t(i)], inf_b);
    }
    if (w.active_set_low[std::size_t(i)]) {
      zl += res.z[std::size_t(i)] * std::max(m.l[std::size_t(i)], -inf_b);
    }
  }
  rhs_duality_gap = std::max(rhs_duality_gap, std::fabs(zu));
  duality_gap += zu;
  rhs_duality_gap = std::max(rhs_duality_gap, std::fabs(zl));
  duality_gap += zl;
  ruiz.scale_dual_in(res.z.data());
  if (box) {
    ruiz.unscale_box_dual_in(res.z.data() + n_in);
    zu = 0;
    zl = 0;
    for (isize i = 0; i < n; ++i) {
      if (w.active_set_up[std::size_t(n_in + i)]) {
        zu += res.z[std::size_t(n_in + i)] * std::min(m.u_box[std::size_t(i)], inf_b);
      }
      if (w.active_set_low[std::size_t(n_in + i)]) {
        zl += res.z[std::size_t(n_in + i)] * std::max(m.l_box[std::size_t(i)], -inf_b);
      }
    }
    rhs_duality_gap = std::max(rhs_duality_gap, std::fabs(zu));
    duality_gap += zu;
    rhs_duality_gap = std::max(rhs_duality_gap, std::fabs(zl));
    duality_gap += zl;
    ruiz.scale_box_dual_in(res.z.data() + n_in);
  }
  w.cnt.n_global_res += 0.5;
}

namespace linesearch {

struct DerivResult
{
  double a, b, grad;
};

// linesearch.hpp:49-167 (GPDAL) and :178-311 (PDAL)
inline DerivResult
derivative_results(QP& qp, double alpha)
{
  Workspace& w = qp.work;
  Results& res = qp.results;
  const Settings& s = qp.settings;
  isize n = qp.model.dim, n_eq = qp.model.n_eq;
  isize ncons = qp.n_constraints();
  const double* dx = w.dw_aug.data();
  const double* dy = w.dw_aug.data() + n;
  const double* dz = w.dw_aug.data() + (isize(w.dw_aug.size()) - ncons);
  bool gpdal = s.merit_function_type == MERIT_GPDAL;
  w.cnt.ls_evals += 1;

  for (isize i = 0; i < ncons; ++i) {
    w.up_plus_alphaCdx[std::size_t(i)] = w.primal_residual_in_scaled_up[std::size_t(i)] + w.Cdx[std::size_t(i)] * alpha;
    w.low_plus_alphaCdx[std::size_t(i)] = res.si[std::size_t(i)] + w.Cdx[std::size_t(i)] * alpha;
  }
  double adx2 = dot(w.Adx.data(), w.Adx.data(), n_eq);
  double a = dot(dx, w.Hdx.data(), n) + res.info.mu_eq_inv * adx2 + res.info.rho * dot(dx, dx, n);
  double sq = 0;
  for (isize i = 0; i < n_eq; ++i) {
    double e = w.Adx[std::size_t(i)] - dy[i] * res.info.mu_eq;
    w.err[std::size_t(n + i)] = e;
    sq += e * e;
  }
  if (gpdal) {
    a += sq * res.info.mu_eq_inv;
  } else {
    a += sq * res.info.mu_eq_inv * res.info.nu;
  }
  for (isize i = 0; i < n; ++i) {
    w.err[std::size_t(i)] = res.info.rho * (res.x[std::size_t(i)] - w.x_prev[std::size_t(i)]) + w.g_scaled[std::size_t(i)];
  }
  double t1 = 0;
  for (isize i = 0; i < n_eq; ++i) {
    t1 += w.Adx[std::size_t(i)] * (res.se[std::size_t(i)] + res.y[std::size_t(i)] * res.info.mu_eq);
  }
  double b = dot(res.x.data(), w.Hdx.data(), n) + dot(w.err.data(), dx, n) + res.info.mu_eq_inv * t1;
  double t2 = 0;
  for (isize i = 0; i < n_eq; ++i) {
    w.rhs[std::size_t(n + i)] = res.se[std::size_t(i)];
    t2 += w.err[std::size_t(n + i)] * res.se[std::size_t(i)];
  }
  if (gpdal) {
    b += res.info.mu_eq_inv * t2;
  } else {
    b += res.info.nu * res.info.mu_eq_inv * t2;
  }
  // Cdx_act and [w-u]_+ + [w-l]_-
  double* errt = w.err.data() + (isize(w.err.size()) - ncons);
  double sq_act = 0, dot_act = 0;
  for (isize i = 0; i < ncons; ++i) {
    bool up = w.up_plus_alphaCdx[std::size_t(i)] > 0.;
    bool low = w.low_plus_alphaCdx[std::size_t(i)] < 0.;
    double cdx_act = (up || low) ? w.Cdx[std::size_t(i)] : 0.;
    errt[i] = cdx_act;
    double apz = (up ? w.primal_residual_in_scaled_up[std::size_t(i)] : 0.) + (low ? res.si[std::size_t(i)] : 0.);
    w.active_part_z[std::size_t(i)] = apz;
    sq_act += cdx_act * cdx_act;
    dot_act += apz * cdx_act;
  }
  if (gpdal) {
    a += res.info.mu_in_inv * sq_act / s.alpha_gpdal;
    a += res.info.mu_in * (1. - s.alpha_gpdal) * dot(dz, dz, ncons);
    b += res.info.mu_in_inv * dot_act / s.alpha_gpdal;
    b += res.info.mu_in * (1. - s.alpha_gpdal) * dot(dz, res.z.data(), ncons);
  } else {
    a += res.info.mu_in_inv * sq_act;
    b += res.info.mu_in_inv * dot_act;
    double sq2 = 0, dot2 = 0;
    for (isize i = 0; i < ncons; ++i) {
      errt[i] -= dz[i] * res.info.mu_in;
      w.active_part_z[std::size_t(i)] -= res.z[std::size_t(i)] * res.info.mu_in;
      sq2 += errt[i] * errt[i];
      dot2 += errt[i] * w.active_part_z[std::size_t(i)];
    }
    a += res.info.nu * res.info.mu_in_inv * sq2;
    b += res.info.nu * res.info.mu_in_inv * dot2;
  }
  return { a, b, a * alpha + b };
}

// linesearch.hpp:320-538
inline void
primal_dual_ls(QP& qp)
{
  Workspace& w = qp.work;
  Results& res = qp.results;
  isize ncons = qp.n_constraints();
  const double machine_eps = std::numeric_limits<double>::epsilon();
  w.alpha = 1.;
  double alpha_ = 1.;
  w.alphas.clear();
  for (isize i = 0; i < ncons; ++i) {
    if (w.Cdx[std::size_t(i)] != 0.) {
      alpha_ = -w.primal_residual_in_scaled_up[std::size_t(i)] / (w.Cdx[std::size_t(i)] + machine_eps);
      if (alpha_ > machine_eps) {
        w.alphas.push_back(alpha_);
      }
      alpha_ = -res.si[std::size_t(i)] / (w.Cdx[std::size_t(i)] + machine_eps);
      if (alpha_ > machine_eps) {
        w.alphas.push_back(alpha_);
      }
    }
  }
  std::sort(w.alphas.begin(), w.alphas.end());
  w.alphas.erase(std::unique(w.alphas.begin(), w.alphas.end()), w.alphas.end());
  isize n_alpha = isize(w.alphas.size());
  if (n_alpha == 0) {
    DerivResult r = derivative_results(qp, 0.);
    w.alpha = -r.b / r.a;
    return;
  }
  const double infty = std::numeric_limits<double>::infinity();
  double last_neg_grad = 0, alpha_last_neg = 0, first_pos_grad = 0, alpha_first_pos = infty;
  for (isize i = 0; i < n_alpha; ++i) {
    alpha_ = w.alphas[std::size_t(i)];
    double gr = derivative_results(qp, alpha_).grad;
    if (gr < 0.) {
      alpha_last_neg = alpha_;
      last_neg_grad = gr;
    } else {
      first_pos_grad = gr;
      alpha_first_pos = alpha_;
      break;
    }
  }
  if (alpha_last_neg == 0.) {
    last_neg_grad = derivative_results(qp, alpha_last_neg).grad;
  }
  if (alpha_first_pos == infty) {
    DerivResult r = derivative_results(qp, 2 * alpha_last_neg + 1);
    w.alpha = -r.b / r.a;
  } else {
    w.alpha = std::fabs(alpha_last_neg - last_neg_grad * (alpha_first_pos - alpha_last_neg) / (first_pos_grad - last_neg_grad));
  }
}

// linesearch.hpp:549-786
inline void
active_set_change(QP& qp)
{
  Workspace& w = qp.work;
  Results& res = qp.results;
  isize n = qp.model.dim, n_eq = qp.model.n_eq;
  isize ncons = qp.n_constraints();
  isize n_c_f = w.n_c;
  w.new_bijection_map = w.current_bijection_map;
  std::vector<isize>& planned = w.ibuf;
  planned.resize(std::size_t(std::max<isize>(1, ncons)));
  {
    isize count = 0;
    for (isize i = 0; i < ncons; ++i) {
      if (w.current_bijection_map[std::size_t(i)] < w.n_c) {
        if (!w.active_inequalities[std::size_t(i)]) {
          planned[std::size_t(count)] = w.current_bijection_map[std::size_t(i)] + n + n_eq;
          ++count;
          const isize bi = w.new_bijection_map[std::size_t(i)];
          for (isize j = 0; j < ncons; ++j) {
            if (w.new_bijection_map[std::size_t(j)] > bi) {
              w.new_bijection_map[std::size_t(j)] -= 1;
            }
          }
          n_c_f -= 1;
          w.new_bijection_map[std::size_t(i)] = ncons - 1;
        }
      }
    }
    std::sort(planned.begin(), planned.begin() + count);
    if (qp.dense_backend == BACKEND_PRIMAL_DUAL_LDLT) {
      w.ldl.delete_at(planned.data(), count);
    } else if (count > 0) {
      w.new_cols.assign(std::size_t(n * count), 0.);
      Vec alpha(std::size_t(count), -res.info.mu_in_inv);
      for (isize i = 0; i < count; ++i) {
        isize index = planned[std::size_t(i)] - (n + n_eq);
        fill_constraint_col(qp, &w.new_cols[std::size_t(i * n)], index);
      }
      w.ldl.rank_r_update(w.new_cols.data(), n, count, alpha.data());
    }
    if (count > 0) {
      w.constraints_changed = true;
    }
  }
  {
    isize count = 0;
    double mu_in_neg = -res.info.mu_in;
    isize n_c = n_c_f;
    for (isize i = 0; i < ncons; ++i) {
      if (w.active_inequalities[std::size_t(i)]) {
        if (w.new_bijection_map[std::size_t(i)] >= n_c_f) {
          planned[std::size_t(count)] = i;
          ++count;
          const isize bi = w.new_bijection_map[std::size_t(i)];
          for (isize j = 0; j < ncons; ++j) {
            if (w.new_bijection_map[std::size_t(j)] < bi && w.new_bijection_map[std::size_t(j)] >= n_c_f) {
              w.new_bijection_map[std::size_t(j)] += 1;
            }
          }
          w.new_bijection_map[std::size_t(i)] = n_c_f;
          n_c_f += 1;
        }
      }
    }
    if (qp.dense_backend == BACKEND_PRIMAL_DUAL_LDLT) {
      isize rows = n + n_eq + n_c_f;
      w.new_cols.assign(std::size_t(rows * std::max<isize>(1, count)), 0.);
      for (isize k = 0; k < count; ++k) {
        isize index = planned[std::size_t(k)];
        double* col = &w.new_cols[std::size_t(k * rows)];
        fill_constraint_col(qp, col, index);
        col[n + n_eq + n_c + k] = mu_in_neg;
      }
      w.ldl.insert_block_at(n + n_eq + n_c, w.new_cols.data(), rows, count);
    } else if (count > 0) {
      w.new_cols.assign(std::size_t(n * count), 0.);
      Vec alpha(std::size_t(count), res.info.mu_in_inv);
      for (isize k = 0; k < count; ++k) {
        fill_constraint_col(qp, &w.new_cols[std::size_t(k * n)], planned[std::size_t(k)]);
      }
      w.ldl.rank_r_update(w.new_cols.data(), n, count, alpha.data());
    }
    if (count > 0) {
      w.constraints_changed = true;
    }
  }
  w.n_c = n_c_f;
  if (n_c_f > w.max_nc) w.max_nc = n_c_f;
  w.current_bijection_map = w.new_bijection_map;
}

} // namespace linesearch

// solver.hpp:564-614
inline void
bcl_update(QP& qp, double& primal_feasibility_lhs_new, double& bcl_eta_ext, double& bcl_eta_in, double bcl_eta_ext_init, double eps_in_min, double& new_bcl_mu_in, double& new_bcl_mu_eq, double& new_bcl_mu_in_inv, double& new_bcl_mu_eq_inv)
{
  const Settings& s = qp.settings;
  Results& res = qp.results;
  if (primal_feasibility_lhs_new <= bcl_eta_ext || res.info.iter > s.safe_guard) {
    bcl_eta_ext *= std::pow(res.info.mu_in, s.beta_bcl);
    bcl_eta_in = std::max(bcl_eta_in * res.info.mu_in, eps_in_min);
  } else {
    res.y = qp.work.y_prev;
    res.z = qp.work.z_prev;
    new_bcl_mu_in = std::max(res.info.mu_in * s.mu_update_factor, s.mu_min_in);
    new_bcl_mu_eq = std::max(res.info.mu_eq * s.mu_update_factor, s.mu_min_eq);
    new_bcl_mu_in_inv = std::min(res.info.mu_in_inv * s.mu_update_inv_factor, s.mu_max_in_inv);
    new_bcl_mu_eq_inv = std::min(res.info.mu_eq_inv * s.mu_update_inv_factor, s.mu_max_eq_inv);
    bcl_eta_ext = bcl_eta_ext_init * std::pow(new_bcl_mu_in, s.alpha_bcl);
    bcl_eta_in = std::max(new_bcl_mu_in, eps_in_min);
  }
}

// solver.hpp:637-677
inline void
Martinez_update(QP& qp, double& primal_feasibility_lhs_new, double& primal_feasibility_lhs_old, double& bcl_eta_in, double eps_in_min, double& new_bcl_mu_in, double& new_bcl_mu_eq, double& new_bcl_mu_in_inv, double& new_bcl_mu_eq_inv)
{
  const Settings& s = qp.settings;
  Results& res = qp.results;
  bcl_eta_in = std::max(bcl_eta_in * 0.1, eps_in_min);
  if (primal_feasibility_lhs_new <= 0.95 * primal_feasibility_lhs_old) {
  } else {
    new_bcl_mu_in = std::max(res.info.mu_in * s.mu_update_factor, s.mu_min_in);
    new_bcl_mu_eq = std::max(res.info.mu_eq * s.mu_update_factor, s.mu_min_eq);
    new_bcl_mu_in_inv = std::min(res.info.mu_in_inv * s.mu_update_inv_factor, s.mu_max_in_inv);
    new_bcl_mu_eq_inv = std::min(res.info.mu_eq_inv * s.mu_update_inv_factor, s.mu_max_eq_inv);
  }
}

// solver.hpp:687-743
inline double
compute_inner_loop_saddle_point(QP& qp)
{
  Workspace& w = qp.work;
  Results& res = qp.results;
  const Settings& s = qp.settings;
  isize n = qp.model.dim, n_eq = qp.model.n_eq;
  isize ncons = qp.n_constraints();
  double f = s.merit_function_type == MERIT_GPDAL ? s.alpha_gpdal : 1.;
  for (isize i = 0; i < ncons; ++i) {
    w.active_part_z[std::size_t(i)] = pos_part(w.primal_residual_in_scaled_up[std::size_t(i)]) + neg_part(res.si[std::size_t(i)]);
    w.active_part_z[std::size_t(i)] -= f * res.z[std::size_t(i)] * res.info.mu_in;
  }
  double err = infty_norm(w.active_part_z);
  for (isize i = 0; i < n_eq; ++i) {
    w.err[std::size_t(n + i)] = res.se[std::size_t(i)];
  }
  err = std::max(err, infty_norm(res.se));
  err = std::max(err, infty_norm(w.dual_residual_scaled));
  return err;
}

// solver.hpp:754-869
inline void
primal_dual_semi_smooth_newton_step(QP& qp, double eps)
{
  Workspace& w = qp.work;
  Results& res = qp.results;
  const Settings& s = qp.settings;
  isize n = qp.model.dim, n_eq = qp.model.n_eq, n_in = qp.model.n_in;
  isize ncons = qp.n_constraints();
  isize numactive = 0;
  for (isize i = 0; i < ncons; ++i) {
    w.active_set_up[std::size_t(i)] = w.primal_residual_in_scaled_up[std::size_t(i)] >= 0;
    w.active_set_low[std::size_t(i)] = res.si[std::size_t(i)] <= 0;
    w.active_inequalities[std::size_t(i)] = w.active_set_up[std::size_t(i)] || w.active_set_low[std::size_t(i)];
    numactive += w.active_inequalities[std::size_t(i)] ? 1 : 0;
  }
  isize inner_pb_dim = n + n_eq + numactive;
  std::fill(w.rhs.begin(), w.rhs.end(), 0.);
  std::fill(w.dw_aug.begin(), w.dw_aug.end(), 0.);
  linesearch::active_set_change(qp);
  for (isize i = 0; i < n; ++i) {
    w.rhs[std::size_t(i)] = -w.dual_residual_scaled[std::size_t(i)];
  }
  if (qp.box_constraints) {
    for (isize i = 0; i < n; ++i) {
      w.active_part_z[std::size_t(n_in + i)] = res.z[std::size_t(n_in + i)] * w.i_scaled[std::size_t(i)];
    }
  }
  for (isize i = 0; i < n_eq; ++i) {
    w.rhs[std::size_t(n + i)] = -res.se[std::size_t(i)];
  }
  double f = s.merit_function_type == MERIT_GPDAL ? s.alpha_gpdal : 1.;
  for (isize i = 0; i < ncons; ++i) {
    isize j = w.current_bijection_map[std::size_t(i)];
    if (j < w.n_c) {
      if (w.active_set_up[std::size_t(i)]) {
        w.rhs[std::size_t(j + n + n_eq)] = -w.primal_residual_in_scaled_up[std::size_t(i)] + res.z[std::size_t(i)] * res.info.mu_in * f;
      } else if (w.active_set_low[std::size_t(i)]) {
        w.rhs[std::size_t(j + n + n_eq)] = -res.si[std::size_t(i)] + res.z[std::size_t(i)] * res.info.mu_in * f;
      }
    } else {
      if (i >= n_in) {
        w.rhs[std::size_t(i - n_in)] += w.active_part_z[std::size_t(i)];
      } else {
        const double zi = res.z[std::size_t(i)];
        const double* r = w.C_scaled.row(i);
        for (isize k = 0; k < n; ++k) {
          w.rhs[std::size_t(k)] += zi * r[k];
        }
      }
    }
  }
  iterative_solve_with_permut_fact(qp, eps, inner_pb_dim);
  for (isize j = 0; j < ncons; ++j) {
    isize i = w.current_bijection_map[std::size_t(j)];
    if (i < w.n_c) {
      w.active_part_z[std::size_t(j)] = w.dw_aug[std::size_t(n + n_eq + i)];
    } else {
      w.active_part_z[std::size_t(j)] = -res.z[std::size_t(j)];
    }
  }
  isize off = isize(w.dw_aug.size()) - ncons;
  for (isize j = 0; j < ncons; ++j) {
    w.dw_aug[std::size_t(off + j)] = w.active_part_z[std::size_t(j)];
  }
}

// solver.hpp:882-1077
inline void
primal_dual_newton_semi_smooth(QP& qp, double eps_int)
{
  Workspace& w = qp.work;
  Results& res = qp.results;
  const Settings& s = qp.settings;
  isize n = qp.model.dim, n_eq = qp.model.n_eq, n_in = qp.model.n_in;
  isize ncons = qp.n_constraints();
  bool box = qp.box_constraints;
  double err_in = 1e6;
  Vec& CTdz = w.tmp_n;
  for (isize iter = 0; iter <= s.max_iter_in; ++iter) {
    if (iter == s.max_iter_in) {
      res.info.iter += s.max_iter_in + 1;
      break;
    }
    primal_dual_semi_smooth_newton_step(qp, eps_int);
    w.cnt.n_newton += 1;
    Vec& Hdx = w.Hdx;
    Vec& Adx = w.Adx;
    Vec& Cdx = w.Cdx;
    Vec& ATdy = w.CTz;
    double* dx = w.dw_aug.data();
    double* dy = w.dw_aug.data() + n;
    double* dz = w.dw_aug.data() + (isize(w.dw_aug.size()) - ncons);
    std::fill(CTdz.begin(), CTdz.end(), 0.);
    if (n_in > 0) {
      gemv(w.C_scaled, dx, Cdx.data());
      gemv_t(w.C_scaled, dz, CTdz.data(), false);
      w.cnt.n_cdx += 1;
    }
    if (box) {
      for (isize i = 0; i < n; ++i) {
        w.active_part_z[std::size_t(n_in + i)] = dz[n_in + i] * w.i_scaled[std::size_t(i)];
        CTdz[std::size_t(i)] += w.active_part_z[std::size_t(n_in + i)];
        Cdx[std::size_t(n_in + i)] = dx[i] * w.i_scaled[std::size_t(i)];
      }
    }
    if (s.merit_function_type == MERIT_GPDAL) {
      for (isize i = 0; i < ncons; ++i) {
        Cdx[std::size_t(i)] += (s.alpha_gpdal - 1.) * res.info.mu_in * dz[i];
      }
    }
    if (n_in > 0 || box) {
      linesearch::primal_dual_ls(qp);
    }
    double alpha = w.alpha;
    if (infty_norm(w.dw_aug) * std::fabs(alpha) < 1e-11 && iter > 0) {
      res.info.iter += iter + 1;
      break;
    }
    for (isize i = 0; i < n; ++i) {
      res.x[std::size_t(i)] += alpha * dx[i];
    }
    for (isize i = 0; i < ncons; ++i) {
      w.primal_residual_in_scaled_up[std::size_t(i)] += alpha * Cdx[std::size_t(i)];
      res.si[std::size_t(i)] += alpha * Cdx[std::size_t(i)];
      res.z[std::size_t(i)] += alpha * dz[i];
    }
    for (isize i = 0; i < n_eq; ++i) {
      res.se[std::size_t(i)] += alpha * (Adx[std::size_t(i)] - res.info.mu_eq * dy[i]);
      res.y[std::size_t(i)] += alpha * dy[i];
    }
    if (qp.hessian_type == HESSIAN_ZERO) {
      for (isize i = 0; i < n; ++i) {
        w.dual_residual_scaled[std::size_t(i)] += alpha * (res.info.rho * dx[i] + ATdy[std::size_t(i)] + CTdz[std::size_t(i)]);
      }
    } else {
      for (isize i = 0; i < n; ++i) {
        w.dual_residual_scaled[std::size_t(i)] += alpha * (res.info.rho * dx[i] + Hdx[std::size_t(i)] + ATdy[std::size_t(i)] + CTdz[std::size_t(i)]);
      }
    }
    err_in = compute_inner_loop_saddle_point(qp);
    if (iter % s.frequence_infeasibility_check == 0 || s.primal_infeasibility_solving) {
      bool is_primal_infeasible = global_primal_residual_infeasibility(qp, ATdy.data(), CTdz.data(), dy, dz);
      bool is_dual_infeasible = global_dual_residual_infeasibility(qp, Adx.data(), Cdx.data(), Hdx.data(), dx);
      if (is_primal_infeasible) {
        res.info.status = PROXQP_PRIMAL_INFEASIBLE;
        if (!s.primal_infeasibility_solving) {
          res.info.iter += iter + 1;
          break;
        }
      } else if (is_dual_infeasible) {
        res.info.status = PROXQP_DUAL_INFEASIBLE;
        res.info.iter += iter + 1;
        break;
      }
    }
    if (err_in <= eps_int) {
      res.info.iter += iter + 1;
      break;
    }
  }
}

inline void
scale_warm_start(QP& qp)
{
  qp.ruiz.scale_primal(qp.results.x.data());
  qp.ruiz.scale_dual_eq(qp.results.y.data());
  qp.ruiz.scale_dual_in(qp.results.z.data());
  if (qp.box_constraints) {
    qp.ruiz.scale_box_dual_in(qp.results.z.data() + qp.model.n_in);
  }
}
inline void
active_set_from_z(QP& qp)
{
  Workspace& w = qp.work;
  isize ncons = qp.n_constraints();
  w.n_c = 0;
  w.max_nc = 0;
  for (isize i = 0; i < ncons; ++i) {
    w.active_inequalities[std::size_t(i)] = qp.results.z[std::size_t(i)] != 0 ? 1 : 0;
  }
  linesearch::active_set_change(qp);
}

// solver.hpp:1088-1843
inline void
qp_solve(QP& qp)
{
  Workspace& w = qp.work;
  Results& res = qp.results;
  const Settings& s = qp.settings;
  const Model& m = qp.model;
  const Ruiz& ruiz = qp.ruiz;
  bool box = qp.box_constraints;
  isize n = m.dim, n_eq = m.n_eq, n_in = m.n_in;
  isize ncons = qp.n_constraints();
  w.ldl.cnt = &w.cnt;

  if (w.dirty) {
    switch (s.initial_guess) {
      case EQUALITY_CONSTRAINED_INITIAL_GUESS:
      case NO_INITIAL_GUESS:
        w.cleanup(box);
        res.cleanup(&s);
        break;
      case COLD_START_WITH_PREVIOUS_RESULT:
      case WARM_START:
        w.cleanup(box);
        res.cold_start(&s);
        scale_warm_start(qp);
        break;
      case WARM_START_WITH_PREVIOUS_RESULT:
        res.cleanup_statistics();
        scale_warm_start(qp);
        break;
    }
    if (s.initial_guess != WARM_START_WITH_PREVIOUS_RESULT) {
      qp.copy_model_to_scaled();
      w.u_scaled = m.u;
      w.l_scaled = m.l;
      qp.setup_equilibration(false);
      setup_factorization(qp);
    }
    switch (s.initial_guess) {
      case EQUALITY_CONSTRAINED_INITIAL_GUESS:
        compute_equality_constrained_initial_guess(qp);
        break;
      case COLD_START_WITH_PREVIOUS_RESULT:
      case WARM_START:
        active_set_from_z(qp);
        break;
      default:
        break;
    }
  } else {
    switch (s.initial_guess) {
      case EQUALITY_CONSTRAINED_INITIAL_GUESS:
        setup_factorization(qp);
        compute_equality_constrained_initial_guess(qp);
        break;
      case COLD_START_WITH_PREVIOUS_RESULT:
      case WARM_START:
        scale_warm_start(qp);
        setup_factorization(qp);
        active_set_from_z(qp);
        break;
      case NO_INITIAL_GUESS:
        setup_factorization(qp);
        break;
      case WARM_START_WITH_PREVIOUS_RESULT:
        scale_warm_start(qp);
        if (w.refactorize) {
          setup_factorization(qp);
          active_set_from_z(qp);
        }
        break;
    }
  }
  double bcl_eta_ext_init = std::pow(0.1, s.alpha_bcl);
  double bcl_eta_ext = bcl_eta_ext_init;
  double bcl_eta_in = 1;
  double eps_in_min = std::min(s.eps_abs, 1e-9);
  double primal_feasibility_eq_rhs_0 = 0, primal_feasibility_in_rhs_0 = 0;
  double dual_feasibility_rhs_0 = 0, dual_feasibility_rhs_1 = 0, dual_feasibility_rhs_3 = 0;
  double primal_feasibility_lhs = 0, primal_feasibility_eq_lhs = 0, primal_feasibility_in_lhs = 0, dual_feasibility_lhs = 0;
  double duality_gap = 0, rhs_duality_gap = 0;
  double scaled_eps = s.eps_abs;

  for (isize iter = 0; iter < s.max_iter; ++iter) {
    global_primal_residual(qp, primal_feasibility_lhs, primal_feasibility_eq_rhs_0, primal_feasibility_in_rhs_0, primal_feasibility_eq_lhs, primal_feasibility_in_lhs);
    global_dual_residual(qp, dual_feasibility_lhs, dual_feasibility_rhs_0, dual_feasibility_rhs_1, dual_feasibility_rhs_3, rhs_duality_gap, duality_gap);
    res.info.pri_res = primal_feasibility_lhs;
    res.info.dua_res = dual_feasibility_lhs;
    res.info.duality_gap = duality_gap;
    if (std::getenv("ORC_TRACE")) std::fprintf(stderr, "orc iter %lld pri %.4e dua %.4e mu_in %.1e status %d iter %lld\n", (long long)iter, primal_feasibility_lhs, dual_feasibility_lhs, res.info.mu_in, int(res.info.status), (long long)res.info.iter);
    double new_bcl_mu_in = res.info.mu_in, new_bcl_mu_eq = res.info.mu_eq;
    double new_bcl_mu_in_inv = res.info.mu_in_inv, new_bcl_mu_eq_inv = res.info.mu_eq_inv;
    double rhs_pri = scaled_eps;
    if (s.eps_rel != 0) {
      rhs_pri += s.eps_rel * std::max(primal_feasibility_eq_rhs_0, primal_feasibility_in_rhs_0);
    }
    bool is_primal_feasible = primal_feasibility_lhs <= rhs_pri;
    double rhs_dua = s.eps_abs;
    if (s.eps_rel != 0) {
      rhs_dua += s.eps_rel * std::max(std::max(dual_feasibility_rhs_3, dual_feasibility_rhs_0), std::max(dual_feasibility_rhs_1, w.dual_feasibility_rhs_2));
    }
    bool is_dual_feasible = dual_feasibility_lhs <= rhs_dua;
    if (is_primal_feasible && is_dual_feasible) {
      if (s.check_duality_gap) {
        if (std::fabs(res.info.duality_gap) <= s.eps_duality_gap_abs + s.eps_duality_gap_rel * rhs_duality_gap) {
          if (s.primal_infeasibility_solving && res.info.status == PROXQP_PRIMAL_INFEASIBLE) {
            res.info.status = PROXQP_SOLVED_CLOSEST_PRIMAL_FEASIBLE;
          } else {
            res.info.status = PROXQP_SOLVED;
          }
          break;
        }
      } else {
        res.info.status = PROXQP_SOLVED;
        break;
      }
    }
    res.info.iter_ext += 1;
    w.x_prev = res.x;
    w.y_prev = res.y;
    w.z_prev = res.z;
    ruiz.scale_primal_residual_in(w.primal_residual_in_scaled_up.data());
    if (box) {
      ruiz.scale_box_primal_residual_in(w.primal_residual_in_scaled_up.data() + n_in);
    }
    for (isize i = 0; i < ncons; ++i) {
      w.primal_residual_in_scaled_up[std::size_t(i)] += w.z_prev[std::size_t(i)] * res.info.mu_in;
    }
    if (s.merit_function_type == MERIT_GPDAL) {
      for (isize i = 0; i < ncons; ++i) {
        w.primal_residual_in_scaled_up[std::size_t(i)] += (s.alpha_gpdal - 1.) * res.info.mu_in * res.z[std::size_t(i)];
      }
    }
    res.si = w.primal_residual_in_scaled_up;
    for (isize i = 0; i < n_in; ++i) {
      w.primal_residual_in_scaled_up[std::size_t(i)] -= w.u_scaled[std::size_t(i)];
      res.si[std::size_t(i)] -= w.l_scaled[std::size_t(i)];
    }
    if (box) {
      for (isize i = 0; i < n; ++i) {
        w.primal_residual_in_scaled_up[std::size_t(n_in + i)] -= w.u_box_scaled[std::size_t(i)];
        res.si[std::size_t(n_in + i)] -= w.l_box_scaled[std::size_t(i)];
      }
    }
    primal_dual_newton_semi_smooth(qp, bcl_eta_in);
    if ((res.info.status == PROXQP_PRIMAL_INFEASIBLE && !s.primal_infeasibility_solving) || res.info.status == PROXQP_DUAL_INFEASIBLE) {
      for (isize i = 0; i < n; ++i) {
        res.x[std::size_t(i)] = w.dw_aug[std::size_t(i)];
      }
      for (isize i = 0; i < n_eq; ++i) {
        res.y[std::size_t(i)] = w.dw_aug[std::size_t(n + i)];
      }
      isize off = isize(w.dw_aug.size()) - ncons;
      for (isize i = 0; i < ncons; ++i) {
        res.z[std::size_t(i)] = w.dw_aug[std::size_t(off + i)];
      }
      break;
    }
    if (scaled_eps == s.eps_abs && s.primal_infeasibility_solving && res.info.status == PROXQP_PRIMAL_INFEASIBLE) {
      Vec ones_eq(std::size_t(n_eq), 1.), ones_in(std::size_t(n_in), 1.);
      gemv_t(m.A, ones_eq.data(), w.rhs.data(), false);
      gemv_t(m.C, ones_in.data(), w.rhs.data(), true);
      if (box) {
        for (isize i = 0; i < n; ++i) {
          w.rhs[std::size_t(i)] += w.i_scaled[std::size_t(i)];
        }
      }
      scaled_eps = infty_norm(w.rhs.data(), n) * s.eps_abs;
      for (isize i = n; i < n + n_eq + n_in; ++i) {
        w.rhs[std::size_t(i)] = 1.;
      }
    }
    double primal_feasibility_lhs_new = primal_feasibility_lhs;
    global_primal_residual(qp, primal_feasibility_lhs_new, primal_feasibility_eq_rhs_0, primal_feasibility_in_rhs_0, primal_feasibility_eq_lhs, primal_feasibility_in_lhs);
    is_primal_feasible = primal_feasibility_lhs_new <= (scaled_eps + s.eps_rel * std::max(primal_feasibility_eq_rhs_0, primal_feasibility_in_rhs_0));
    res.info.pri_res = primal_feasibility_lhs_new;
    if (is_primal_feasible) {
      double dual_feasibility_lhs_new = dual_feasibility_lhs;
      global_dual_residual(qp, dual_feasibility_lhs_new, dual_feasibility_rhs_0, dual_feasibility_rhs_1, dual_feasibility_rhs_3, rhs_duality_gap, duality_gap);
      res.info.dua_res = dual_feasibility_lhs_new;
      res.info.duality_gap = duality_gap;
      is_dual_feasible = dual_feasibility_lhs_new <= (s.eps_abs + s.eps_rel * std::max(std::max(dual_feasibility_rhs_3, dual_feasibility_rhs_0), std::max(dual_feasibility_rhs_1, w.dual_feasibility_rhs_2)));
      if (is_dual_feasible) {
        bool gap_ok = !s.check_duality_gap || std::fabs(res.info.duality_gap) <= s.eps_duality_gap_abs + s.eps_duality_gap_rel * rhs_duality_gap;
        if (gap_ok) {
          if (s.primal_infeasibility_solving && res.info.status == PROXQP_PRIMAL_INFEASIBLE) {
            res.info.status = PROXQP_SOLVED_CLOSEST_PRIMAL_FEASIBLE;
          } else {
            res.info.status = PROXQP_SOLVED;
          }
        }
      }
    }
    if (s.bcl_update) {
      bcl_update(qp, primal_feasibility_lhs_new, bcl_eta_ext, bcl_eta_in, bcl_eta_ext_init, eps_in_min, new_bcl_mu_in, new_bcl_mu_eq, new_bcl_mu_in_inv, new_bcl_mu_eq_inv);
    } else {
      Martinez_update(qp, primal_feasibility_lhs_new, primal_feasibility_lhs, bcl_eta_in, eps_in_min, new_bcl_mu_in, new_bcl_mu_eq, new_bcl_mu_in_inv, new_bcl_mu_eq_inv);
    }
    double dual_feasibility_lhs_new = dual_feasibility_lhs;
    global_dual_residual(qp, dual_feasibility_lhs_new, dual_feasibility_rhs_0, dual_feasibility_rhs_1, dual_feasibility_rhs_3, rhs_duality_gap, duality_gap);
    res.info.dua_res = dual_feasibility_lhs_new;
    res.info.duality_gap = duality_gap;
    if (primal_feasibility_lhs_new >= primal_feasibility_lhs && dual_feasibility_lhs_new >= dual_feasibility_lhs && res.info.mu_in <= 1e-5) {
      new_bcl_mu_in = s.cold_reset_mu_in;
      new_bcl_mu_eq = s.cold_reset_mu_eq;
      new_bcl_mu_in_inv = s.cold_reset_mu_in_inv;
      new_bcl_mu_eq_inv = s.cold_reset_mu_eq_inv;
    }
    if (res.info.mu_in != new_bcl_mu_in || res.info.mu_eq != new_bcl_mu_eq) {
      ++res.info.mu_updates;
      mu_update(qp, new_bcl_mu_eq, new_bcl_mu_in);
    }
    res.info.mu_eq = new_bcl_mu_eq;
    res.info.mu_in = new_bcl_mu_in;
    res.info.mu_eq_inv = new_bcl_mu_eq_inv;
    res.info.mu_in_inv = new_bcl_mu_in_inv;
  }
  ruiz.unscale_primal(res.x.data());
  ruiz.unscale_dual_eq(res.y.data());
  ruiz.unscale_dual_in(res.z.data());
  if (box) {
    ruiz.unscale_box_dual_in(res.z.data() + n_in);
  }
  if (s.primal_infeasibility_solving && res.info.status == PROXQP_PRIMAL_INFEASIBLE) {
    ruiz.unscale_primal_residual_eq(res.se.data());
    ruiz.unscale_primal_residual_in(res.si.data());
    if (box) {
      ruiz.unscale_box_primal_residual_in(res.si.data() + n_in);
    }
  }
  {
    // solver.hpp:1769-1781: objective from the lower-triangle walk of model.H
    double obj = 0;
    for (isize j = 0; j < n; ++j) {
      obj += 0.5 * (res.x[std::size_t(j)] * res.x[std::size_t(j)]) * m.H(j, j);
      double acc = 0;
      for (isize i = j + 1; i < n; ++i) {
        acc += m.H(i, j) * res.x[std::size_t(i)];
      }
      obj += res.x[std::size_t(j)] * acc;
    }
    obj += dot(m.g.data(), res.x.data(), n);
    res.info.objValue = obj;
  }
  w.dirty = true;
  w.is_initialized = true;
}

// parallel/qp_solve.hpp:17-60
inline void
solve_in_parallel(std::vector<QP>& qps, isize num_threads = 0)
{
#ifdef _OPENMP
  if (num_threads <= 0) {
    num_threads = std::max<isize>(omp_get_max_threads() / 2, 1);
  }
  omp_set_num_threads(int(num_threads));
  omp_set_dynamic(0);
#endif
  isize batch = isize(qps.size());
#pragma omp parallel for schedule(dynamic)
  for (isize i = 0; i < batch; ++i) {
    qps[std::size_t(i)].solve();
  }
}

} // namespace oracle
