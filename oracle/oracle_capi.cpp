// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE. See proxqp_oracle.hpp.
//
// Plain C entry points over the C++ restatement so that tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs can drive it through ctypes.
#include "proxqp_oracle.hpp"
#include "../proxsuite_b200/csrc/random_qp.hpp"
#include <chrono>
#include <cstring>
#include <memory>

using namespace oracle;

namespace {
thread_local std::string g_err;

struct Batch
{
  std::vector<QP> qps;
};

double*
setting_ptr_d(Settings& s, const char* name)
{
#define F(x)                                                                                                                                                                                                                                                   \
  if (std::strcmp(name, #x) == 0)                                                                                                                                                                                                                              \
    return &s.x;
  F(default_rho) F(default_mu_eq) F(default_mu_in) F(alpha_bcl) F(beta_bcl) F(refactor_dual_feasibility_threshold) F(refactor_rho_threshold) F(mu_min_eq) F(mu_min_in) F(mu_max_eq_inv) F(mu_max_in_inv) F(mu_update_factor) F(mu_update_inv_factor)
    F(cold_reset_mu_eq) F(cold_reset_mu_in) F(cold_reset_mu_eq_inv) F(cold_reset_mu_in_inv) F(eps_abs) F(eps_rel) F(eps_refact) F(eps_duality_gap_abs) F(eps_duality_gap_rel) F(preconditioner_accuracy) F(eps_primal_inf) F(eps_dual_inf) F(alpha_gpdal)
      F(default_H_eigenvalue_estimate)
#undef F
        return nullptr;
}
} // namespace

extern "C" {

const char*
orc_last_error()
{
  return g_err.c_str();
}

void*
orc_qp_create(long long n, long long n_eq, long long n_in, int box, int hessian, int backend)
{
  try {
    return new QP(n, n_eq, n_in, box != 0, hessian, backend);
  } catch (std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void
orc_qp_destroy(void* p)
{
  delete static_cast<QP*>(p);
}

int
orc_qp_set(void* p, const char* name, double v)
{
  QP& qp = *static_cast<QP*>(p);
  Settings& s = qp.settings;
  if (double* d = setting_ptr_d(s, name)) {
    *d = v;
    return 0;
  }
#define I(x)                                                                                                                                                                                                                                                   \
  if (std::strcmp(name, #x) == 0) {                                                                                                                                                                                                                            \
    s.x = decltype(s.x)(v);                                                                                                                                                                                                                                    \
    return 0;                                                                                                                                                                                                                                                  \
  }
  I(max_iter) I(max_iter_in) I(safe_guard) I(nb_iterative_refinement) I(verbose) I(initial_guess) I(update_preconditioner) I(compute_preconditioner) I(compute_timings) I(check_duality_gap) I(preconditioner_max_iter) I(bcl_update) I(merit_function_type)
    I(primal_infeasibility_solving) I(frequence_infeasibility_check)
#undef I
      g_err = std::string("unknown setting ") + name;
  return -1;
}
double
orc_qp_get(void* p, const char* name)
{
  QP& qp = *static_cast<QP*>(p);
  Settings& s = qp.settings;
  if (double* d = setting_ptr_d(s, name)) {
    return *d;
  }
#define I(x)                                                                                                                                                                                                                                                   \
  if (std::strcmp(name, #x) == 0)                                                                                                                                                                                                                              \
    return double(s.x);
  I(max_iter) I(max_iter_in) I(safe_guard) I(nb_iterative_refinement) I(verbose) I(initial_guess) I(update_preconditioner) I(compute_preconditioner) I(compute_timings) I(check_duality_gap) I(preconditioner_max_iter) I(bcl_update) I(merit_function_type)
    I(primal_infeasibility_solving) I(frequence_infeasibility_check)
#undef I
      if (std::strcmp(name, "dense_backend") == 0) return double(qp.dense_backend);
  if (std::strcmp(name, "max_nc") == 0) return double(qp.work.max_nc);
  return std::numeric_limits<double>::quiet_NaN();
}

// NULL pointers mean "absent" (nullopt in the reference).
int
orc_qp_init(void* p, const double* H, const double* g, const double* A, const double* b, const double* C, const double* l, const double* u, const double* l_box, const double* u_box, int compute_preconditioner, const double* rho, const double* mu_eq, const double* mu_in, const double* manual_eig)
{
  try {
    QPData d{ H, g, A, b, C, l, u, l_box, u_box };
    static_cast<QP*>(p)->init(d, compute_preconditioner != 0, rho, mu_eq, mu_in, manual_eig);
    return 0;
  } catch (std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
int
orc_qp_update(void* p, const double* H, const double* g, const double* A, const double* b, const double* C, const double* l, const double* u, const double* l_box, const double* u_box, int update_preconditioner, const double* rho, const double* mu_eq, const double* mu_in, const double* manual_eig)
{
  try {
    QPData d{ H, g, A, b, C, l, u, l_box, u_box };
    static_cast<QP*>(p)->update(d, update_preconditioner != 0, rho, mu_eq, mu_in, manual_eig);
    return 0;
  } catch (std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
int
orc_qp_solve(void* p, const double* x, const double* y, const double* z)
{
  try {
    static_cast<QP*>(p)->solve(x, y, z);
    return 0;
  } catch (std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
void
orc_qp_cleanup(void* p)
{
  static_cast<QP*>(p)->cleanup();
}

// info layout (doubles): 0 mu_eq, 1 mu_eq_inv, 2 mu_in, 3 mu_in_inv, 4 rho, 5 nu,
// 6 iter, 7 iter_ext, 8 mu_updates, 9 rho_updates, 10 status, 11 setup_time,
// 12 solve_time, 13 run_time, 14 objValue, 15 pri_res, 16 dua_res,
// 17 duality_gap, 18 iterative_residual, 19 minimal_H_eigenvalue_estimate
void
orc_qp_results(void* p, double* x, double* y, double* z, double* se, double* si, double* info)
{
  QP& qp = *static_cast<QP*>(p);
  const Results& r = qp.results;
  if (x) std::copy(r.x.begin(), r.x.end(), x);
  if (y) std::copy(r.y.begin(), r.y.end(), y);
  if (z) std::copy(r.z.begin(), r.z.end(), z);
  if (se) std::copy(r.se.begin(), r.se.end(), se);
  if (si) std::copy(r.si.begin(), r.si.end(), si);
  if (info) {
    const Info& i = r.info;
    double v[20] = { i.mu_eq, i.mu_eq_inv, i.mu_in, i.mu_in_inv, i.rho, i.nu, double(i.iter), double(i.iter_ext), double(i.mu_updates), double(i.rho_updates), double(i.status), i.setup_time, i.solve_time, i.run_time, i.objValue, i.pri_res,
                     i.dua_res, i.duality_gap, i.iterative_residual, i.minimal_H_eigenvalue_estimate };
    std::copy(v, v + 20, info);
  }
}

// scaled problem + Ruiz scaling (for the equilibration identity test,
// test/src/dense_ruiz_equilibration.cpp:63-71)
void
orc_qp_scaled(void* p, double* H, double* g, double* A, double* b, double* C, double* u, double* l, double* delta, double* c)
{
  QP& qp = *static_cast<QP*>(p);
  const Workspace& w = qp.work;
  if (H) std::copy(w.H_scaled.a.begin(), w.H_scaled.a.end(), H);
  if (g) std::copy(w.g_scaled.begin(), w.g_scaled.end(), g);
  if (A) std::copy(w.A_scaled.a.begin(), w.A_scaled.a.end(), A);
  if (b) std::copy(w.b_scaled.begin(), w.b_scaled.end(), b);
  if (C) std::copy(w.C_scaled.a.begin(), w.C_scaled.a.end(), C);
  if (u) std::copy(w.u_scaled.begin(), w.u_scaled.end(), u);
  if (l) std::copy(w.l_scaled.begin(), w.l_scaled.end(), l);
  if (delta) std::copy(qp.ruiz.delta.begin(), qp.ruiz.delta.end(), delta);
  if (c) *c = qp.ruiz.c;
}

// 19 counters in declaration order of oracle::Counters
void
orc_qp_counters(void* p, double* out, int reset)
{
  QP& qp = *static_cast<QP*>(p);
  Counters& c = qp.work.cnt;
  double v[19] = { c.n_factor, c.factor_m2, c.factor_m3, c.n_solve, c.solve_m2, c.solve_m, c.n_resid, c.resid_nc, c.rank_rt2, c.rank_chunk_t2, c.rank_rt, c.n_insert, c.insert_bytes, c.n_delete, c.delete_t2, c.ls_evals, c.n_cdx, c.n_global_res,
                   c.n_newton };
  if (out) std::copy(v, v + 19, out);
  if (reset) c = Counters();
}

// compute_backward (dense/compute_ECJ.hpp:29-125) on a solved QP; any output may be NULL
int
orc_qp_backward(void* p, const double* loss_derivative, double eps, double rho_new, double mu_new, double* dL_dH, double* dL_dg, double* dL_dA, double* dL_db, double* dL_dC, double* dL_du, double* dL_dl)
{
  try {
    QP& qp = *static_cast<QP*>(p);
    BackwardData bd;
    compute_backward(qp, loss_derivative, bd, eps, rho_new, mu_new);
    if (dL_dH) std::copy(bd.dL_dH.a.begin(), bd.dL_dH.a.end(), dL_dH);
    if (dL_dg) std::copy(bd.dL_dg.begin(), bd.dL_dg.end(), dL_dg);
    if (dL_dA) std::copy(bd.dL_dA.a.begin(), bd.dL_dA.a.end(), dL_dA);
    if (dL_db) std::copy(bd.dL_db.begin(), bd.dL_db.end(), dL_db);
    if (dL_dC) std::copy(bd.dL_dC.a.begin(), bd.dL_dC.a.end(), dL_dC);
    if (dL_du) std::copy(bd.dL_du.begin(), bd.dL_du.end(), dL_du);
    if (dL_dl) std::copy(bd.dL_dl.begin(), bd.dL_dl.end(), dL_dl);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

long long
orc_qp_max_nc(void* p)
{
  return static_cast<QP*>(p)->work.max_nc;
}

// ---- batch (std::vector<QP> / BatchQP, wrapper.hpp:1253-1311) --------------
void*
orc_batch_create(long long batch, long long n, long long n_eq, long long n_in, int box, int hessian, int backend)
{
  try {
    std::unique_ptr<Batch> b(new Batch);
    b->qps.reserve(std::size_t(batch));
    for (long long i = 0; i < batch; ++i) {
      b->qps.emplace_back(n, n_eq, n_in, box != 0, hessian, backend);
    }
    return b.release();
  } catch (std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void
orc_batch_destroy(void* p)
{
  delete static_cast<Batch*>(p);
}
void*
orc_batch_qp(void* p, long long i)
{
  return &static_cast<Batch*>(p)->qps[std::size_t(i)];
}
long long
orc_batch_size(void* p)
{
  return (long long)static_cast<Batch*>(p)->qps.size();
}
// solve_in_parallel(qps, num_threads); num_threads <= 0 -> reference default
// max(omp_get_max_threads()/2, 1); returns wall seconds.
double
orc_batch_solve(void* p, long long num_threads)
{
  Batch& b = *static_cast<Batch*>(p);
  auto t0 = std::chrono::steady_clock::now();
  solve_in_parallel(b.qps, num_threads);
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}
// serial loop, benchmark/timings-parallel.cpp:198-206
double
orc_batch_solve_serial(void* p)
{
  Batch& b = *static_cast<Batch*>(p);
  auto t0 = std::chrono::steady_clock::now();
  for (auto& q : b.qps) {
    q.solve();
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}
void
orc_batch_counters(void* p, double* out, int reset)
{
  Batch& b = *static_cast<Batch*>(p);
  Counters tot;
  for (auto& q : b.qps) {
    tot.add(q.work.cnt);
    if (reset) q.work.cnt = Counters();
  }
  double v[19] = { tot.n_factor, tot.factor_m2, tot.factor_m3, tot.n_solve, tot.solve_m2, tot.solve_m, tot.n_resid, tot.resid_nc, tot.rank_rt2, tot.rank_chunk_t2, tot.rank_rt, tot.n_insert, tot.insert_bytes, tot.n_delete, tot.delete_t2, tot.ls_evals,
                   tot.n_cdx, tot.n_global_res, tot.n_newton };
  std::copy(v, v + 19, out);
}
int
orc_omp_max_threads()
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// ---- LDLT unit access for tests -------------------------------------------
// Factorise a symmetric m x m matrix (row-major), then apply a script of
// operations; used by tests/test_oracle_ldlt.py.
void*
orc_ldlt_create(const double* mat, long long m, long long cap)
{
  Ldlt* l = new Ldlt;
  l->reserve(cap);
  l->factorize(mat, m, m);
  return l;
}
void
orc_ldlt_destroy(void* p)
{
  delete static_cast<Ldlt*>(p);
}
long long
orc_ldlt_dim(void* p)
{
  return static_cast<Ldlt*>(p)->dim();
}
void
orc_ldlt_solve(void* p, double* rhs)
{
  Ldlt* l = static_cast<Ldlt*>(p);
  l->solve_in_place(rhs, l->dim());
}
void
orc_ldlt_reconstruct(void* p, double* out)
{
  std::vector<double> o;
  static_cast<Ldlt*>(p)->reconstruct(o);
  std::copy(o.begin(), o.end(), out);
}
void
orc_ldlt_delete_at(void* p, const long long* idx, long long r)
{
  std::vector<isize> v(idx, idx + r);
  static_cast<Ldlt*>(p)->delete_at(v.data(), r);
}
// a: column-major (dim + r) x r
void
orc_ldlt_insert_block_at(void* p, long long i, const double* a, long long r)
{
  Ldlt* l = static_cast<Ldlt*>(p);
  l->insert_block_at(i, a, l->dim() + r, r);
}
void
orc_ldlt_diagonal_update(void* p, const long long* idx, long long r, const double* alpha)
{
  std::vector<isize> v(idx, idx + r);
  static_cast<Ldlt*>(p)->diagonal_update_clobber_indices(v.data(), r, alpha);
}
// w: column-major dim x r
void
orc_ldlt_rank_r_update(void* p, const double* w, long long r, const double* alpha)
{
  Ldlt* l = static_cast<Ldlt*>(p);
  l->rank_r_update(w, l->dim(), r, alpha);
}

// ---- generators (product header, reference-specified inputs) --------------
void
orc_lehmer_stream(unsigned long long seed, long long count, double* uniforms)
{
  pqp::randqp::Lehmer rng;
  rng.set_seed(seed);
  for (long long i = 0; i < count; ++i) {
    uniforms[i] = rng.uniform();
  }
}
// kind: 0 dense_strongly_convex_qp, 1 dense_not_strongly_convex_qp,
// 2 dense_degenerate_qp (C,u,l have 2*n_in rows), 3 dense_box_constrained_qp,
// 4 box benchmark (timings-box-constraints.cpp), 5 diagonal-Hessian benchmark.
// seed < 0: continue the caller-visible global stream (not supported) -> error.
int
orc_gen_qp(int kind, unsigned long long seed, int n, int n_eq, int n_in, double sparsity, double strong_convexity, double* H, double* g, double* A, double* b, double* C, double* u, double* l, double* u_box, double* l_box)
{
  pqp::randqp::Lehmer rng;
  rng.set_seed(seed);
  switch (kind) {
    case 0: pqp::randqp::dense_strongly_convex_qp(rng, n, n_eq, n_in, sparsity, strong_convexity, H, g, A, b, C, u, l); return 0;
    case 1: pqp::randqp::dense_not_strongly_convex_qp(rng, n, n_eq, n_in, sparsity, H, g, A, b, C, u, l); return 0;
    case 2: pqp::randqp::dense_degenerate_qp(rng, n, n_eq, n_in, sparsity, strong_convexity, H, g, A, b, C, u, l); return 0;
    case 3: pqp::randqp::dense_box_constrained_qp(rng, n, n_eq, n_in, sparsity, strong_convexity, H, g, A, b, C, u, l); return 0;
    case 4: pqp::randqp::dense_box_benchmark_qp(rng, n, n_eq, n_in, sparsity, strong_convexity, 0, H, g, A, b, C, u, l, u_box, l_box); return 0;
    case 5: pqp::randqp::dense_box_benchmark_qp(rng, n, n_eq, n_in, sparsity, strong_convexity, 1, H, g, A, b, C, u, l, u_box, l_box); return 0;
  }
  return -1;
}
} // extern "C"
