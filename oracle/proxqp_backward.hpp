// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE. See proxqp_oracle.hpp.
//
// Restatement of the QPLayer backward pass of the reference (SURVEY.md section 8,
// row f2): dense/compute_ECJ.hpp:29-190 (compute_backward,
// compute_backward_loss_ESG) and dense/backward_data.hpp:27-129, file:line under
// /root/reference/include/proxsuite/proxqp. One extra solve of the regularised
// KKT system with the active set at the solution, then outer products.
// Pinned by the reference's own acceptance test test/src/dense_backward.cpp
// (backward Jacobians against central finite differences, |diff| < 1e-5),
// restated in tests/test_oracle_backward.py.
#pragma once

namespace oracle {

// dense/backward_data.hpp:27-129
struct BackwardData
{
  Mat dL_dH, dL_dA, dL_dC;
  Vec dL_dg, dL_db, dL_du, dL_dl;
  void initialize(isize dim, isize n_eq, isize n_in)
  {
    if (dL_dH.rows != dim || dL_dA.rows != n_eq || dL_dC.rows != n_in || dL_dH.cols != dim) {
      dL_dH = Mat(dim, dim);
      dL_dA = Mat(n_eq, dim);
      dL_dC = Mat(n_in, dim);
    }
    dL_dH.set_zero();
    dL_dA.set_zero();
    dL_dC.set_zero();
    dL_dg.assign(std::size_t(dim), 0.);
    dL_db.assign(std::size_t(n_eq), 0.);
    dL_du.assign(std::size_t(n_in), 0.);
    dL_dl.assign(std::size_t(n_in), 0.);
  }
};

// dense/compute_ECJ.hpp:127-188
inline void
compute_backward_loss_ESG(QP& qp, const double* loss_derivative, BackwardData& bd)
{
  Workspace& w = qp.work;
  const Model& m = qp.model;
  const Results& res = qp.results;
  const isize n = m.dim, n_eq = m.n_eq, n_in = m.n_in;
  // unpermuted dz step (compute_ECJ.hpp:131-143; the reference indexes loss_derivative with the permuted
  // position i of an inactive constraint, restated as is)
  for (isize j = 0; j < n_in; ++j) {
    const isize i = w.current_bijection_map[std::size_t(j)];
    if (i < w.n_c) {
      w.active_part_z[std::size_t(j)] = w.dw_aug[std::size_t(n + n_eq + i)];
    } else {
      w.active_part_z[std::size_t(j)] = loss_derivative[n + n_eq + i];
    }
  }
  for (isize j = 0; j < n_in; ++j) {
    w.dw_aug[std::size_t(n + n_eq + j)] = w.active_part_z[std::size_t(j)];
  }
  double* dx = w.dw_aug.data();
  double* dy = dx + n;
  double* dz = dy + n_eq;
  qp.ruiz.unscale_primal(dx);
  qp.ruiz.unscale_dual_eq(dy);
  qp.ruiz.unscale_dual_in(dz);
  // compute_ECJ.hpp:156-187
  for (isize i = 0; i < n_in; ++i) {
    for (isize j = 0; j < n; ++j) {
      bd.dL_dC(i, j) = dz[i] * res.x[std::size_t(j)] + res.z[std::size_t(i)] * dx[j];
    }
    bd.dL_du[std::size_t(i)] = w.active_set_up[std::size_t(i)] ? -dz[i] : 0.;
    bd.dL_dl[std::size_t(i)] = w.active_set_low[std::size_t(i)] ? -dz[i] : 0.;
  }
  for (isize i = 0; i < n_eq; ++i) {
    for (isize j = 0; j < n; ++j) {
      bd.dL_dA(i, j) = dy[i] * res.x[std::size_t(j)] + res.y[std::size_t(i)] * dx[j];
    }
    bd.dL_db[std::size_t(i)] = -dy[i];
  }
  for (isize i = 0; i < n; ++i) {
    for (isize j = 0; j < n; ++j) {
      bd.dL_dH(i, j) = 0.5 * (dx[i] * res.x[std::size_t(j)] + res.x[std::size_t(i)] * dx[j]);
    }
    bd.dL_dg[std::size_t(i)] = dx[i];
  }
}

// dense/compute_ECJ.hpp:29-125. `loss_derivative` has dim + n_eq + n_in entries (dL/dx, dL/dy, dL/dz).
inline void
compute_backward(QP& qp, const double* loss_derivative, BackwardData& bd, double eps = 1.E-4, double rho_new = 1.E-6, double mu_new = 1.E-6)
{
  if (qp.results.info.status == PROXQP_DUAL_INFEASIBLE) {
    throw std::invalid_argument("the QP problem is not feasible, so computing the derivatives is not valid in this setting. Try enabling infeasible solving if the problem is only primally infeasible.");
  }
  if (qp.box_constraints) {
    throw std::invalid_argument("compute_backward: box constraints are not handled by the reference's backward pass");
  }
  Workspace& w = qp.work;
  const Model& m = qp.model;
  Results& res = qp.results;
  const isize n = m.dim, n_eq = m.n_eq, n_in = m.n_in;
  w.ldl.cnt = &w.cnt;
  bd.initialize(n, n_eq, n_in);
  // active set at the solution, in the model's units (:52-61)
  isize numactive = 0;
  for (isize i = 0; i < n_in; ++i) {
    const double ctz = dot(m.C.row(i), res.x.data(), n) + res.z[std::size_t(i)];
    w.active_set_up[std::size_t(i)] = (ctz - m.u[std::size_t(i)]) >= 0. ? 1 : 0;
    w.active_set_low[std::size_t(i)] = (ctz - m.l[std::size_t(i)]) <= 0. ? 1 : 0;
    w.active_inequalities[std::size_t(i)] = (w.active_set_up[std::size_t(i)] || w.active_set_low[std::size_t(i)]) ? 1 : 0;
    numactive += w.active_inequalities[std::size_t(i)];
  }
  const isize inner_pb_dim = n + n_eq + numactive;
  std::fill(w.rhs.begin(), w.rhs.end(), 0.);
  // new proximal parameters (:66-68; the inverses are left as they are, like the reference)
  res.info.rho = rho_new;
  res.info.mu_eq = mu_new;
  res.info.mu_in = mu_new;
  // factorisation from scratch + the whole active set in one block (:74-90)
  setup_factorization(qp);
  w.n_c = 0;
  for (isize i = 0; i < n_in; ++i) {
    w.current_bijection_map[std::size_t(i)] = i;
    w.new_bijection_map[std::size_t(i)] = i;
  }
  linesearch::active_set_change(qp);
  w.constraints_changed = false; // no refactorisation afterwards (:91)
  // rhs = -loss_derivative, scaled block by block (:93-118)
  for (isize i = 0; i < n + n_eq + n_in; ++i) {
    w.rhs[std::size_t(i)] = -loss_derivative[i];
  }
  qp.ruiz.scale_dual_residual(w.rhs.data());
  bool eq_zero = true, in_zero = true;
  for (isize i = 0; i < n_eq; ++i) eq_zero = eq_zero && w.rhs[std::size_t(n + i)] == 0.;
  for (isize i = 0; i < n_in; ++i) in_zero = in_zero && w.rhs[std::size_t(n + n_eq + i)] == 0.;
  if (!eq_zero) {
    for (isize i = 0; i < n_eq; ++i) w.rhs[std::size_t(n + i)] = -loss_derivative[n + i];
    qp.ruiz.scale_primal_residual_eq(w.rhs.data() + n);
  }
  if (!in_zero) {
    // restated as written (:105-117): the in-place scaling sits inside the loop over i
    for (isize i = 0; i < n_in; ++i) {
      const isize j = w.current_bijection_map[std::size_t(i)];
      if (j < w.n_c) {
        w.rhs[std::size_t(j + n + n_eq)] = -loss_derivative[i + n + n_eq];
      }
      qp.ruiz.scale_primal_residual_in(w.rhs.data() + n + n_eq);
    }
  }
  iterative_solve_with_permut_fact(qp, eps, inner_pb_dim); // the full rhs is zeroed inside (:119-128)
  compute_backward_loss_ESG(qp, loss_derivative, bd);
}

} // namespace oracle
