// proxsuite_b200.hpp — header-only C++17 host mirror of the reference's dense ProxQP interface
// on top of the C-ABI (pqp.h / libpqp_b200.so).
//
// Same names, argument order and error behaviour as
//   proxsuite::proxqp::dense::QP<T>            dense/wrapper.hpp:115-963
//   proxsuite::proxqp::dense::BatchQP<T>       dense/wrapper.hpp:1253-1311
//   proxsuite::proxqp::dense::solve_in_parallel parallel/qp_solve.hpp:17-60
//   proxsuite::proxqp::Settings<T> / Results<T> / Info<T>   settings.hpp:88-316, results.hpp:28-203
// with two differences forced by the boundary: matrices are passed as ROW-MAJOR `const double*`
// (what Eigen's RowMajor `.data()` yields, dense/fwd.hpp:16-33) instead of Eigen::Ref, and the
// numbers live on the GPU, so `results` is filled when a solve completes. There is no CPU
// fallback: without a usable CUDA device construction throws std::runtime_error.
//
// A QP created through BatchQP::init_qp_in_place shares ONE device batch with the other QPs of
// the same shape, and solve_in_parallel(BatchQP&) is ONE persistent kernel launch per shape —
// that is the fast path. A free-standing QP owns a batch of one.
#ifndef PROXSUITE_B200_HPP
#define PROXSUITE_B200_HPP

#include "pqp.h"

#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace proxsuite_b200 {
namespace proxqp {

using isize = std::int64_t;
template<typename T>
using optional = std::optional<T>;
constexpr std::nullopt_t nullopt = std::nullopt;

// status.hpp:17-43, settings.hpp:26-45
enum struct QPSolverOutput
{
  PROXQP_SOLVED = PQP_SOLVED,
  PROXQP_MAX_ITER_REACHED = PQP_MAX_ITER_REACHED,
  PROXQP_PRIMAL_INFEASIBLE = PQP_PRIMAL_INFEASIBLE,
  PROXQP_SOLVED_CLOSEST_PRIMAL_FEASIBLE = PQP_SOLVED_CLOSEST_PRIMAL_FEASIBLE,
  PROXQP_DUAL_INFEASIBLE = PQP_DUAL_INFEASIBLE,
  PROXQP_NOT_RUN = PQP_NOT_RUN
};
enum struct InitialGuessStatus
{
  NO_INITIAL_GUESS = PQP_NO_INITIAL_GUESS,
  EQUALITY_CONSTRAINED_INITIAL_GUESS = PQP_EQUALITY_CONSTRAINED_INITIAL_GUESS,
  WARM_START_WITH_PREVIOUS_RESULT = PQP_WARM_START_WITH_PREVIOUS_RESULT,
  WARM_START = PQP_WARM_START,
  COLD_START_WITH_PREVIOUS_RESULT = PQP_COLD_START_WITH_PREVIOUS_RESULT
};
enum struct DenseBackend
{
  Automatic = PQP_BACKEND_AUTOMATIC,
  PrimalDualLDLT = PQP_BACKEND_PRIMAL_DUAL_LDLT,
  PrimalLDLT = PQP_BACKEND_PRIMAL_LDLT
};
enum struct HessianType
{
  Zero = PQP_HESSIAN_ZERO,
  Dense = PQP_HESSIAN_DENSE,
  Diagonal = PQP_HESSIAN_DIAGONAL
};
enum struct MeritFunctionType
{
  GPDAL = PQP_MERIT_GPDAL,
  PDAL = PQP_MERIT_PDAL
};

// Settings<double>: the C struct IS the field-for-field mirror (pqp.h); C++ code uses the same
// member names as the reference (`qp.settings.eps_abs = 1e-9;`). Enumerations are int32 members.
using Settings = pqp_settings;
using Info = pqp_info;

namespace detail {
inline void
check(int rc)
{
  if (rc == PQP_OK) return;
  const std::string msg = pqp_last_error();
  if (rc == PQP_EINVAL) throw std::invalid_argument(msg); // what the reference throws (macros.hpp:18-35)
  throw std::runtime_error(msg);
}

// one device batch of same-shaped QPs
struct Group
{
  pqp_batch* handle = nullptr;
  isize capacity = 0, used = 0;
  isize dim = 0, n_eq = 0, n_in = 0;
  bool box = false;
  int backend = PQP_BACKEND_PRIMAL_DUAL_LDLT;
  Group(isize cap, isize dim_, isize n_eq_, isize n_in_, bool box_, HessianType h, DenseBackend be, int device)
    : capacity(cap)
    , dim(dim_)
    , n_eq(n_eq_)
    , n_in(n_in_)
    , box(box_)
  {
    handle = pqp_batch_create(cap, dim_, n_eq_, n_in_, box_ ? 1 : 0, int(h), int(be), device);
    if (!handle) {
      const std::string msg = pqp_last_error();
      if (dim_ == 0) throw std::invalid_argument(msg);
      throw std::runtime_error(msg);
    }
    int bx = 0, ht = 0;
    pqp_batch_dims(handle, nullptr, nullptr, nullptr, &bx, &ht, &backend);
  }
  Group(const Group&) = delete;
  Group& operator=(const Group&) = delete;
  ~Group()
  {
    if (handle) pqp_batch_destroy(handle);
  }
};
} // namespace detail

// Results<double> (results.hpp:67-203): x, y, z (z = [z_C ; z_box]), se, si, info
struct Results
{
  std::vector<double> x, y, z, se, si;
  Info info{};
};

namespace dense {

struct Model
{
  isize dim = 0, n_eq = 0, n_in = 0, n_total = 0;
};

class QP
{
public:
  Results results;
  Settings settings{};
  Model model;

  // QP<T>(dim, n_eq, n_in[, box_constraints][, HessianType][, DenseBackend]) — wrapper.hpp:140-333
  QP(isize dim, isize n_eq, isize n_in, bool box_constraints = false, HessianType hessian_type = HessianType::Dense, DenseBackend dense_backend = DenseBackend::PrimalDualLDLT, int device = -1)
    : QP(std::make_shared<detail::Group>(1, dim, n_eq, n_in, box_constraints, hessian_type, dense_backend, device))
  {
  }
  QP(isize dim, isize n_eq, isize n_in, bool box_constraints, DenseBackend dense_backend, HessianType hessian_type = HessianType::Dense)
    : QP(dim, n_eq, n_in, box_constraints, hessian_type, dense_backend)
  {
  }

  bool is_box_constrained() const { return group_->box; }

  // init(H, g, A, b, C, l, u, compute_preconditioner, rho, mu_eq, mu_in, manual_minimal_H_eigenvalue)
  // wrapper.hpp:354-498. nullptr = nullopt (absent).
  void init(const double* H, const double* g, const double* A, const double* b, const double* C, const double* l, const double* u, bool compute_preconditioner = true, optional<double> rho = nullopt, optional<double> mu_eq = nullopt,
            optional<double> mu_in = nullopt, optional<double> manual_minimal_H_eigenvalue = nullopt)
  {
    init(H, g, A, b, C, l, u, nullptr, nullptr, compute_preconditioner, rho, mu_eq, mu_in, manual_minimal_H_eigenvalue);
  }
  // box overload, wrapper.hpp:520-703
  void init(const double* H, const double* g, const double* A, const double* b, const double* C, const double* l, const double* u, const double* l_box, const double* u_box, bool compute_preconditioner = true, optional<double> rho = nullopt,
            optional<double> mu_eq = nullopt, optional<double> mu_in = nullopt, optional<double> manual_minimal_H_eigenvalue = nullopt)
  {
    push_settings();
    detail::check(pqp_batch_init(group_->handle, index_, 1, H, g, A, b, C, l, u, l_box, u_box, compute_preconditioner ? 1 : 0, ptr(rho), ptr(mu_eq), ptr(mu_in), ptr(manual_minimal_H_eigenvalue)));
    pull_settings();
  }
  // update(...) — wrapper.hpp:723-918; nullptr = unchanged
  void update(const double* H, const double* g, const double* A, const double* b, const double* C, const double* l, const double* u, bool update_preconditioner = false, optional<double> rho = nullopt, optional<double> mu_eq = nullopt,
              optional<double> mu_in = nullopt, optional<double> manual_minimal_H_eigenvalue = nullopt)
  {
    update(H, g, A, b, C, l, u, nullptr, nullptr, update_preconditioner, rho, mu_eq, mu_in, manual_minimal_H_eigenvalue);
  }
  void update(const double* H, const double* g, const double* A, const double* b, const double* C, const double* l, const double* u, const double* l_box, const double* u_box, bool update_preconditioner = false, optional<double> rho = nullopt,
              optional<double> mu_eq = nullopt, optional<double> mu_in = nullopt, optional<double> manual_minimal_H_eigenvalue = nullopt)
  {
    push_settings();
    detail::check(pqp_batch_update(group_->handle, index_, 1, H, g, A, b, C, l, u, l_box, u_box, update_preconditioner ? 1 : 0, ptr(rho), ptr(mu_eq), ptr(mu_in), ptr(manual_minimal_H_eigenvalue)));
    pull_settings();
  }
  // solve() — wrapper.hpp:922 (blocking; for several QPs use solve_in_parallel)
  void solve()
  {
    push_settings();
    const int64_t me = (int64_t)index_; // this QP only: the siblings of a shared device batch keep their results
    detail::check(pqp_batch_select(group_->handle, &me, 1));
    detail::check(pqp_batch_solve(group_->handle));
    fetch();
  }
  // solve(x, y, z) — wrapper.hpp:935-954: warm start from the given guess (nullptr = absent)
  void solve(const double* x, const double* y, const double* z)
  {
    detail::check(pqp_batch_warm_start(group_->handle, index_, 1, x, y, z));
    pull_settings();
    solve();
  }
  // cleanup() — wrapper.hpp:958-962
  void cleanup()
  {
    detail::check(pqp_batch_cleanup(group_->handle, index_, 1));
    fetch();
  }

  // -- used by BatchQP / solve_in_parallel ---------------------------------------------------
  void push_settings() { detail::check(pqp_batch_settings_set(group_->handle, index_, &settings)); }
  void pull_settings() { detail::check(pqp_batch_settings_get(group_->handle, index_, &settings)); }
  void fetch()
  {
    detail::check(pqp_batch_results(group_->handle, index_, 1, results.x.data(), results.y.data(), results.z.data(), results.se.data(), results.si.data(), &results.info));
    pull_settings();
  }
  detail::Group* group() const { return group_.get(); }
  isize index() const { return index_; }

private:
  friend class BatchQP;
  explicit QP(std::shared_ptr<detail::Group> g)
    : group_(std::move(g))
  {
    if (group_->used >= group_->capacity) throw std::invalid_argument("BatchQP: more QPs than the reserved batch size");
    index_ = group_->used++;
    const isize n = group_->dim, ne = group_->n_eq, ni = group_->n_in, nc = ni + (group_->box ? n : 0);
    model = Model{ n, ne, ni, n + ne + ni };
    results.x.assign(std::size_t(n), 0.0);
    results.y.assign(std::size_t(ne), 0.0);
    results.z.assign(std::size_t(nc), 0.0);
    results.se.assign(std::size_t(ne), 0.0);
    results.si.assign(std::size_t(nc), 0.0);
    pull_settings();
    detail::check(pqp_batch_results(group_->handle, index_, 1, nullptr, nullptr, nullptr, nullptr, nullptr, &results.info));
  }
  static const double* ptr(const optional<double>& v) { return v ? &*v : nullptr; }
  std::shared_ptr<detail::Group> group_;
  isize index_ = 0;
};

// BatchQP<T> — wrapper.hpp:1253-1311. QPs of the same shape share one device batch.
class BatchQP
{
public:
  explicit BatchQP(std::size_t batch_size)
    : capacity_(isize(batch_size))
  {
    qps_.reserve(batch_size);
  }
  // init_qp_in_place(dim, n_eq, n_in) — wrapper.hpp:1276-1283 (the reference has no box / Hessian
  // arguments here; they are accepted as an extension)
  QP& init_qp_in_place(isize dim, isize n_eq, isize n_in, bool box_constraints = false, HessianType hessian_type = HessianType::Dense, DenseBackend dense_backend = DenseBackend::PrimalDualLDLT)
  {
    if (dim == 0) throw std::invalid_argument("wrong argument size: the dimension wrt the primal variable x should be strictly positive.");
    const auto key = std::make_tuple(dim, n_eq, n_in, box_constraints, int(hessian_type), int(dense_backend));
    auto it = groups_.find(key);
    if (it == groups_.end() || it->second->used >= it->second->capacity)
      it = groups_.insert_or_assign(key, std::make_shared<detail::Group>(capacity_, dim, n_eq, n_in, box_constraints, hessian_type, dense_backend, -1)).first;
    qps_.push_back(std::unique_ptr<QP>(new QP(it->second)));
    return *qps_.back();
  }
  QP& get(isize i) { return *qps_.at(std::size_t(i)); }
  QP& operator[](isize i) { return *qps_[std::size_t(i)]; }
  isize size() const { return isize(qps_.size()); }

private:
  isize capacity_;
  std::vector<std::unique_ptr<QP>> qps_;
  std::map<std::tuple<isize, isize, isize, bool, int, int>, std::shared_ptr<detail::Group>> groups_;
};

namespace detail_parallel {
template<class It>
inline void
solve_groups(It first, It last)
{
  std::vector<proxsuite_b200::proxqp::detail::Group*> groups;
  for (It it = first; it != last; ++it) {
    QP& q = **it;
    q.push_settings();
    auto* g = q.group();
    bool seen = false;
    for (auto* h : groups) seen = seen || h == g;
    if (!seen) groups.push_back(g);
  }
  for (auto* g : groups) { // only the listed members of each device batch (parallel/qp_solve.hpp:17-38)
    std::vector<int64_t> idx;
    for (It it = first; it != last; ++it) {
      if ((**it).group() == g) idx.push_back((int64_t)(**it).index());
    }
    proxsuite_b200::proxqp::detail::check(pqp_batch_select(g->handle, idx.data(), (int64_t)idx.size()));
    proxsuite_b200::proxqp::detail::check(pqp_batch_solve_async(g->handle, nullptr));
  }
  for (auto* g : groups) proxsuite_b200::proxqp::detail::check(pqp_batch_sync(g->handle));
  for (It it = first; it != last; ++it) (**it).fetch();
}
} // namespace detail_parallel

// solve_in_parallel(BatchQP&, num_threads) — parallel/qp_solve.hpp:40-60. One persistent kernel
// launch per shape; `num_threads` is accepted for signature parity and ignored (the work
// distribution is the kernel's atomic queue).
inline void
solve_in_parallel(BatchQP& qps, const optional<std::size_t> num_threads = nullopt)
{
  (void)num_threads;
  std::vector<QP*> ptrs;
  for (isize i = 0; i < qps.size(); ++i) ptrs.push_back(&qps[i]);
  detail_parallel::solve_groups(ptrs.begin(), ptrs.end());
}
// solve_in_parallel(std::vector<QP>&, num_threads) — parallel/qp_solve.hpp:17-38. Free-standing QPs
// own a batch of one each (one launch per QP): correct, but BatchQP is the fast path.
inline void
solve_in_parallel(std::vector<QP>& qps, const optional<std::size_t> num_threads = nullopt)
{
  (void)num_threads;
  std::vector<QP*> ptrs;
  for (auto& q : qps) ptrs.push_back(&q);
  detail_parallel::solve_groups(ptrs.begin(), ptrs.end());
}

// One batch of same-shaped QPs over several GPUs of one node from ONE process (pqp_sharded_*): contiguous slices,
// no cross-device dependency inside the iteration (parallel/qp_solve.hpp:55-59). Stacked row-major host arrays,
// laid out like pqp_batch_init's.
class ShardedBatch
{
public:
  Settings settings{};
  ShardedBatch(isize batch, isize dim, isize n_eq, isize n_in, const std::vector<int>& devices, bool box_constraints = false, HessianType hessian_type = HessianType::Dense,
               DenseBackend dense_backend = DenseBackend::PrimalDualLDLT)
    : handle_(pqp_sharded_create(batch, dim, n_eq, n_in, box_constraints ? 1 : 0, int(hessian_type), int(dense_backend), devices.data(), int(devices.size())))
  {
    if (!handle_) throw std::runtime_error(std::string("proxsuite_b200: ") + pqp_last_error());
    pqp_settings_default(&settings, pqp_dense_backend_choice(int(dense_backend), dim, n_eq, n_in, box_constraints ? 1 : 0));
  }
  ~ShardedBatch() { pqp_sharded_destroy(handle_); }
  ShardedBatch(const ShardedBatch&) = delete;
  ShardedBatch& operator=(const ShardedBatch&) = delete;
  void init(const double* H, const double* g, const double* A, const double* b, const double* C, const double* l, const double* u, const double* l_box = nullptr, const double* u_box = nullptr, bool compute_preconditioner = true)
  {
    proxsuite_b200::proxqp::detail::check(pqp_sharded_settings_set(handle_, &settings));
    proxsuite_b200::proxqp::detail::check(pqp_sharded_init(handle_, H, g, A, b, C, l, u, l_box, u_box, compute_preconditioner ? 1 : 0, nullptr, nullptr, nullptr, nullptr));
  }
  void solve()
  {
    proxsuite_b200::proxqp::detail::check(pqp_sharded_settings_set(handle_, &settings));
    proxsuite_b200::proxqp::detail::check(pqp_sharded_solve(handle_));
  }
  void results(double* x, double* y, double* z, pqp_info* info = nullptr, double* se = nullptr, double* si = nullptr)
  {
    proxsuite_b200::proxqp::detail::check(pqp_sharded_results(handle_, x, y, z, se, si, info));
  }
  int shards() const { return pqp_sharded_count(handle_); }

private:
  pqp_sharded* handle_;
};

} // namespace dense
} // namespace proxqp
} // namespace proxsuite_b200

#endif // PROXSUITE_B200_HPP
