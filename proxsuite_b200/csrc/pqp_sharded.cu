// One batch sharded over several GPUs of one node, behind the C-ABI (include/pqp.h, pqp_sharded_*).
// solve_in_parallel has no cross-QP state (reference parallel/qp_solve.hpp:55-59), so a batch shards into contiguous
// slices, one per-device pqp_batch each. Host code only: everything goes through the public
// per-device entry points, so a C++ caller of pqp.h gets every GPU of the box without torch.distributed.
#include "../../include/pqp.h"

#include <cstdint>
#include <string>
#include <vector>

struct pqp_sharded
{
  int64_t B = 0, n = 0, ne = 0, ni = 0, nc = 0;
  int box = 0;
  std::vector<pqp_batch*> shard;
  std::vector<int64_t> first, count;
};

namespace {
// the slice of shard k, like proxsuite_b200/sharding.py:shard_bounds (the first B % G shards own one more QP)
void
bounds(int64_t B, int G, int k, int64_t& lo, int64_t& hi)
{
  const int64_t base = B / G, rem = B % G;
  lo = k * base + (k < rem ? k : rem);
  hi = lo + base + (k < rem ? 1 : 0);
}
const double*
at(const double* p, int64_t first, int64_t per_qp)
{
  return p ? p + first * per_qp : nullptr;
}
double*
at(double* p, int64_t first, int64_t per_qp)
{
  return p ? p + first * per_qp : nullptr;
}
} // namespace

extern "C" {

pqp_sharded*
pqp_sharded_create(int64_t batch, int64_t dim, int64_t n_eq, int64_t n_in, int box_constraints, int hessian_type, int dense_backend, const int* devices, int n_devices)
{
  if (batch <= 0 || n_devices <= 0 || !devices) return nullptr;
  pqp_sharded* s = new pqp_sharded();
  s->B = batch;
  s->n = dim;
  s->ne = n_eq;
  s->ni = n_in;
  s->box = box_constraints;
  s->nc = n_in + (box_constraints ? dim : 0);
  const int G = (int)(n_devices < batch ? n_devices : batch);
  for (int k = 0; k < G; ++k) {
    int64_t lo, hi;
    bounds(batch, G, k, lo, hi);
    pqp_batch* b = pqp_batch_create(hi - lo, dim, n_eq, n_in, box_constraints, hessian_type, dense_backend, devices[k]);
    if (!b) { // pqp_last_error() holds the reason
      pqp_sharded_destroy(s);
      return nullptr;
    }
    s->shard.push_back(b);
    s->first.push_back(lo);
    s->count.push_back(hi - lo);
  }
  return s;
}

void
pqp_sharded_destroy(pqp_sharded* s)
{
  if (!s) return;
  for (pqp_batch* b : s->shard) pqp_batch_destroy(b);
  delete s;
}

int
pqp_sharded_count(const pqp_sharded* s)
{
  return s ? (int)s->shard.size() : 0;
}

pqp_batch*
pqp_sharded_shard(pqp_sharded* s, int k, int64_t* first, int64_t* count)
{
  if (!s || k < 0 || k >= (int)s->shard.size()) return nullptr;
  if (first) *first = s->first[k];
  if (count) *count = s->count[k];
  return s->shard[k];
}

int
pqp_sharded_settings_set(pqp_sharded* s, const pqp_settings* in)
{
  if (!s || !in) return PQP_EINVAL;
  for (pqp_batch* b : s->shard) {
    if (int rc = pqp_batch_settings_set(b, -1, in)) return rc;
  }
  return 0;
}

static int
feed(pqp_sharded* s, bool update, const double* H, const double* g, const double* A, const double* b_, const double* C, const double* l, const double* u, const double* l_box, const double* u_box, int flag, const double* rho, const double* mu_eq,
     const double* mu_in, const double* eig)
{
  if (!s) return PQP_EINVAL;
  const int64_t n = s->n, ne = s->ne, ni = s->ni;
  // every shard's upload is enqueued before any is waited for: the copies of the devices overlap
  for (size_t k = 0; k < s->shard.size(); ++k) {
    const int64_t f = s->first[k], c = s->count[k];
    auto fn = update ? pqp_batch_update : pqp_batch_init;
    if (int rc = fn(s->shard[k], 0, c, at(H, f, n * n), at(g, f, n), at(A, f, ne * n), at(b_, f, ne), at(C, f, ni * n), at(l, f, ni), at(u, f, ni), at(l_box, f, n), at(u_box, f, n), flag, rho, mu_eq, mu_in, eig)) return rc;
  }
  return 0;
}

int
pqp_sharded_init(pqp_sharded* s, const double* H, const double* g, const double* A, const double* b_, const double* C, const double* l, const double* u, const double* l_box, const double* u_box, int compute_preconditioner, const double* rho,
                 const double* mu_eq, const double* mu_in, const double* manual_minimal_H_eigenvalue)
{
  return feed(s, false, H, g, A, b_, C, l, u, l_box, u_box, compute_preconditioner, rho, mu_eq, mu_in, manual_minimal_H_eigenvalue);
}

int
pqp_sharded_update(pqp_sharded* s, const double* H, const double* g, const double* A, const double* b_, const double* C, const double* l, const double* u, const double* l_box, const double* u_box, int update_preconditioner, const double* rho,
                   const double* mu_eq, const double* mu_in, const double* manual_minimal_H_eigenvalue)
{
  return feed(s, true, H, g, A, b_, C, l, u, l_box, u_box, update_preconditioner, rho, mu_eq, mu_in, manual_minimal_H_eigenvalue);
}

int
pqp_sharded_solve(pqp_sharded* s)
{
  if (!s) return PQP_EINVAL;
  for (pqp_batch* b : s->shard) { // enqueue on every device first ...
    if (int rc = pqp_batch_solve_async(b, nullptr)) return rc;
  }
  for (pqp_batch* b : s->shard) { // ... then wait for all of them
    if (int rc = pqp_batch_sync(b)) return rc;
  }
  return 0;
}

int
pqp_sharded_results(pqp_sharded* s, double* x, double* y, double* z, double* se, double* si, pqp_info* info)
{
  if (!s) return PQP_EINVAL;
  for (size_t k = 0; k < s->shard.size(); ++k) {
    const int64_t f = s->first[k], c = s->count[k];
    if (int rc = pqp_batch_results(s->shard[k], 0, c, at(x, f, s->n), at(y, f, s->ne), at(z, f, s->nc), at(se, f, s->ne), at(si, f, s->nc), info ? info + f : nullptr)) return rc;
  }
  return 0;
}

} // extern "C"
