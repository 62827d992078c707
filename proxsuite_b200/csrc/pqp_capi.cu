// Host side of the C-ABI (include/pqp.h): device memory management, the
// init / update / warm-start / solve state machine of the reference's QP<T>
// object applied to a whole batch, shared-memory layout policy and launches.
//
// Reference semantics followed here (relative to
// /root/reference/include/proxsuite/proxqp):
//   dense/wrapper.hpp:354-498, 520-703   QP::init
//   dense/wrapper.hpp:723-918            QP::update
//   dense/wrapper.hpp:922-962            QP::solve / cleanup
//   dense/helpers.hpp:176-189, 502-572, 680-763   setup / proximal parameters / warm_start
//   dense/solver.hpp:1125-1377           qp_solve prologue (dirty / initial guess matrix)
//   results.hpp:149-203                  Results::cleanup / cold_start / cleanup_statistics
//   dense/workspace.hpp:330-377          Workspace::cleanup (flags)
// There is no CPU fallback anywhere in this file: without a usable CUDA
// device every entry point that computes returns PQP_ECUDA.
#include "pqp_device.h"
#include "random_qp.hpp"
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

int
fail(int code, const std::string& msg)
{
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                                                                                                                                                                                         \
  do {                                                                                                                                                                                                                                                         \
    cudaError_t e__ = (expr);                                                                                                                                                                                                                                  \
    if (e__ != cudaSuccess) return fail(PQP_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));                                                                                                                                                     \
  } while (0)

struct QpFlags
{
  bool dirty = false, refactorize = false, proximal_parameter_update = false, is_initialized = false;
};

} // namespace

#define PQP_UPLOAD_CHUNKS 8
#define PQP_CSTREAMS 4 // compute streams the chunks of a pipelined init + solve rotate over
#define PQP_FEED_CHUNKS 16 // upload chunks of the fused feed (a 4-byte progress word follows each)

struct pqp_batch
{
  int64_t B = 0;
  PqpDims d{};
  int backend = PQP_BACKEND_PRIMAL_DUAL_LDLT;
  int device = 0;
  PqpBatchPtrs p{};
  std::vector<void*> allocs;
  std::vector<PqpQpParams> hparams; // settings (host truth) + launch parameters
  std::vector<pqp_info> hinfo;      // results.info (host truth between solves)
  std::vector<QpFlags> flags;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr; // host->device uploads of init(), chunked so that the set-up kernels overlap them
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  cudaEvent_t ev_main = nullptr, ev_chunk[PQP_UPLOAD_CHUNKS] = {};
  cudaStream_t cstream[PQP_CSTREAMS] = {};
  cudaEvent_t ev_cdone[PQP_CSTREAMS] = {};
  // chunks of the last chunked init() whose set-up kernels are enqueued on cstream[k % PQP_CSTREAMS];
  // a solve() that follows directly runs one launch per chunk on the same streams, so early
  // chunks are solved while later ones are still crossing PCIe
  int nchunks_pending = 0;
  int64_t chunk_first[PQP_UPLOAD_CHUNKS] = {}, chunk_count[PQP_UPLOAD_CHUNKS] = {};
  int64_t ws_slot_doubles = 0; // workspace stride between concurrently running chunk launches
  bool setup_timed = false, solve_timed = false;
  PqpLayout lay{};      // primary layout (fast kernel when the inverse blocks are in shared memory)
  PqpLayout lay_gen{};  // fallback: everything but the vectors in global memory, full capacity
  PqpLayout lay_big{};  // tile layout with the largest S^-1 capacity one CTA per SM allows (first retry level); kind 0 = unused
  int grid = 0, grid_gen = 0, grid_big = 0;
  int32_t* counter = nullptr;
  double* ws = nullptr;
  double* dbg = nullptr;
  int dbg_cap = 0;
  long long* prof = nullptr;
  int64_t launches = 0;
  int64_t overflow_retries = 0;
  bool solve_pending = false;
  // Fused feed: a whole-batch init() / update() only uploads; the equilibration is done by the persistent solve
  // kernel (the CTA that pops a QP sets it up, then solves it), gated per QP on the progress of the upload.
  int deferred = 0;            // 0: nothing pending; else PqpSolveArgs::fused_setup code of the pending set-up
  bool deferred_gated = false; // the upload runs on copy_stream and announces its progress through d_ready
  bool fused_ok = false;       // the solve kernel's shared memory holds the set-up scratch
  int32_t* d_ready = nullptr;  // [0] QPs uploaded so far, [1] abort flag
  int32_t* h_ready = nullptr;  // pinned: cumulative QP count behind each upload chunk
  cudaEvent_t ev_feed = nullptr;
  cudaEvent_t ev_r0 = nullptr, ev_r1 = nullptr; // around the retry launches of a sync()
  float retry_ms = 0;                           // device time of the retry launches of the last solve
  std::vector<uint8_t> selected;                // pqp_batch_select: QPs the next solve addresses (empty: all)
  int64_t l2_window_bytes = 0, l2_persist_bytes = 0; // access-policy window of the workspace (0: unsupported)
  // QPLayer backward (allocated on first use): loss derivatives in, BackwardData out
  double *bw_loss = nullptr, *bw_dH = nullptr, *bw_dg = nullptr, *bw_dA = nullptr, *bw_db = nullptr, *bw_dC = nullptr, *bw_du = nullptr, *bw_dl = nullptr;
};

namespace {

template<class T>
int
dev_alloc(pqp_batch* b, T** out, size_t count)
{
  void* ptr = nullptr;
  size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  cudaError_t e = cudaMalloc(&ptr, bytes);
  if (e != cudaSuccess) return fail(PQP_ECUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  cudaMemset(ptr, 0, bytes);
  b->allocs.push_back(ptr);
  *out = static_cast<T*>(ptr);
  return 0;
}

void
info_defaults(pqp_info& i, const pqp_settings* s, int backend)
{
  // results.hpp:90-143 (ctor) + cold_start(settings) :175-194
  i.rho = (backend == PQP_BACKEND_PRIMAL_LDLT) ? 1e-5 : 1e-6;
  i.mu_eq = 1e-3;
  i.mu_eq_inv = 1e3;
  i.mu_in = 1e-1;
  i.mu_in_inv = 1e1;
  i.nu = 1.0;
  i.minimal_H_eigenvalue_estimate = 0;
  if (s) {
    i.rho = s->default_rho;
    i.mu_eq = s->default_mu_eq;
    i.mu_eq_inv = 1.0 / i.mu_eq;
    i.mu_in = s->default_mu_in;
    i.mu_in_inv = 1.0 / i.mu_in;
    i.minimal_H_eigenvalue_estimate = s->default_H_eigenvalue_estimate;
  }
}
void
cleanup_statistics(pqp_info& i)
{
  // results.hpp:157-174
  i.run_time = i.setup_time = i.solve_time = 0;
  i.objValue = 0;
  i.iter = i.iter_ext = i.mu_updates = i.rho_updates = 0;
  i.pri_res = i.dua_res = i.duality_gap = i.iterative_residual = 0;
  i.status = PQP_MAX_ITER_REACHED;
}
void
cold_start(pqp_info& i, const pqp_settings* s, int backend)
{
  info_defaults(i, s, backend);
  cleanup_statistics(i);
}

// Shared-memory placement policy. Three layouts, all served by the same device
// code (arrays are addressed through per-array pointers):
//   compact : vectors + S^-1 (capacity si_cap) in shared memory, P^-1 / A_s / G
//             in the per-CTA global workspace (L2-resident) -> TWO CTAs per SM.
//             The solver is latency bound, so a second
//             resident QP per SM nearly doubles throughput.
//   full    : vectors, P^-1, S^-1 (full capacity) and A_s in shared memory,
//             one CTA per SM.
//   generic : only the vectors in shared memory; used for shapes that fit
//             neither, and to re-solve the rare QPs whose active set outgrows
//             si_cap in the compact layout.
int
fill_layout(const PqpDims& d, PqpLayout& L, int64_t budget_bytes, bool want_m1, bool want_ms, bool want_as, int si_cap, int ctas)
{
  std::memset(&L, 0, sizeof(L));
  const int n = d.n, ne = d.ne, nc = d.nc, cap = d.cap;
  auto rnd = [](int64_t v) { return (v + 1) & ~int64_t(1); };
  int vsz[V_COUNT];
  for (int& v : vsz) v = 0;
  vsz[V_X] = n; vsz[V_Y] = ne; vsz[V_Z] = nc; vsz[V_XP] = n; vsz[V_YP] = ne; vsz[V_ZP] = nc;
  vsz[V_DX] = n; vsz[V_DS] = cap; vsz[V_DZ] = nc;
  vsz[V_RX] = n; vsz[V_RS] = cap; vsz[V_EX] = n; vsz[V_ES] = cap;
  vsz[V_DUAL] = n; vsz[V_SE] = ne; vsz[V_RUP] = nc; vsz[V_SI] = nc;
  vsz[V_HDX] = n; vsz[V_ADX] = ne; vsz[V_ATDY] = n; vsz[V_CDX] = nc; vsz[V_CTDZ] = n; vsz[V_Q] = n;
  vsz[V_GS] = n; vsz[V_BS] = ne; vsz[V_US] = nc; vsz[V_LS] = nc; vsz[V_IS] = n; vsz[V_DELTA] = n + ne + nc;
  vsz[V_B] = ne; vsz[V_U] = nc; vsz[V_L] = nc;
  vsz[V_D1INV] = n; vsz[V_DSV] = 2; vsz[V_DSINV] = 2;
  vsz[V_T1] = n; vsz[V_T2] = n; vsz[V_T3] = n;
  vsz[V_S1] = cap + 1; vsz[V_S2] = cap + 1; vsz[V_S3] = cap + 1; vsz[V_S4] = cap + 1;
  vsz[V_ALPHAS] = 2 * nc + 2; vsz[V_GRADS] = 4;
  // partial sums: NW x 32*NG for the symmetric mat-vec, NW x n for the row passes
  const int ncols = (n <= 128 && cap <= 128) ? 128 : ((n <= 160 && cap <= 160) ? 160 : 256);
  vsz[V_SCRATCH] = std::max<int>(PQP_NW * ncols, (n <= 256 && (n % 2) == 0) ? PQP_NW * n : PQP_NT);
  vsz[V_SCRATCH] = std::max<int>(vsz[V_SCRATCH], 8 * ((std::max(n, cap) + 2) & ~1)); // 8 panel vectors of the blocked sweep
  vsz[V_SCRATCH] = std::max<int>(vsz[V_SCRATCH], PQP_NW * (std::max(n, cap) + 2));     // NW partial vectors of the any-n row passes / symmetric mat-vec
  vsz[V_RED] = PQP_NW * 16; // block_reduce: up to 10 values per warp
  vsz[V_KT] = 2;
  vsz[V_KT2] = 2;
  int off = 0;
  for (int v = 0; v < V_COUNT; ++v) {
    L.voff[v] = off;
    off += (int)rnd(vsz[v]);
  }
  L.vec_doubles = off;
  L.scratch_doubles = vsz[V_SCRATCH];
  L.si_cap = si_cap;
  L.ctas_per_sm = ctas;
  int64_t sz[PA_COUNT];
  sz[PA_M1] = (d.hess == PQP_HESSIAN_DENSE) ? rnd((int64_t)n * (n + 1) / 2) : 2; // P^-1, packed with diagonal
  sz[PA_AS] = rnd((int64_t)ne * n);
  sz[PA_MS] = rnd((int64_t)si_cap * (si_cap + 1) / 2 + 2);                       // S^-1, packed with diagonal
  // build_Pi sweeps P inside the S^-1 region whenever P^-1 itself is not in shared memory: the region must hold
  // n (n + 1) / 2 doubles as well (fewer constraint rows than variables: cap < n)
  if (d.hess == PQP_HESSIAN_DENSE) sz[PA_MS] = std::max(sz[PA_MS], rnd((int64_t)n * (n + 1) / 2 + 2));
  sz[PA_G] = rnd((int64_t)cap * (cap + 1) / 2 + 2);
  sz[PA_Y] = 2;
  sz[PA_VEC] = L.vec_doubles;
  const int64_t nlist = std::max(nc, cap);
  L.smem_int_bytes = (int32_t)((4 * (nc + cap + nlist + nc + 2 * PQP_NW + 8) + 2 * nc + 15) & ~15);
  const int64_t budget = budget_bytes - 1664 /*static shared memory of the kernel (general kernels: 1568 B in the fused instantiation); the budget must hold for the FUSED instantiation too: round 2 found it running at one CTA per SM, 16 bytes over half an SM*/ - L.smem_int_bytes;
  int64_t smem_d = 0, ws_d = 0;
  auto put = [&](int id, bool want_smem) {
    if (want_smem && (smem_d + sz[id]) * 8 <= budget) {
      L.in_smem[id] = 1;
      L.off[id] = smem_d;
      smem_d += sz[id];
    } else {
      L.in_smem[id] = 0;
      L.off[id] = ws_d;
      ws_d += sz[id];
    }
  };
  put(PA_VEC, true);
  put(PA_MS, want_ms);
  put(PA_M1, want_m1);
  put(PA_AS, want_as);
  put(PA_G, false);
  put(PA_Y, false);
  L.smem_doubles = (int32_t)smem_d;
  L.ws_doubles = std::max<int64_t>(ws_d, 2);
  return 0;
}

// Doubles of the 32 x 32 tile storage (row stride 33) of a symmetric matrix of capacity cap
// (pqp_fast_body.inl: ts_extent(cap, cap)).
int64_t
tile_doubles(int64_t cap)
{
  const int64_t nb = (cap + 31) / 32;
  if (nb == 0) return 2;
  const int64_t rows_last = cap - 32 * (nb - 1);
  return ((nb - 1) * nb / 2) * (32 * 33) + nb * rows_last * 33 + 2;
}

// Tile layout (kind 1) of the specialised kernel: vectors + S^-1 (tile storage, capacity
// si_cap <= 128) in shared memory; P^-1 (n x ldn), Bt (n x ldb) and G in the per-CTA workspace.
int
fill_layout_tile(const PqpDims& d, PqpLayout& L, int64_t budget_bytes, int si_cap, int ctas, bool pi_smem = false)
{
  std::memset(&L, 0, sizeof(L));
  const int n = d.n, ne = d.ne, nc = d.nc, cap = d.cap;
  auto rnd = [](int64_t v) { return (v + 1) & ~int64_t(1); };
  const int sc = si_cap + 2; // slot-indexed vectors
  int vsz[V_COUNT];
  for (int& v : vsz) v = 0;
  vsz[V_X] = n; vsz[V_Y] = ne; vsz[V_Z] = nc; vsz[V_XP] = n; vsz[V_YP] = ne; vsz[V_ZP] = nc;
  vsz[V_DX] = n; vsz[V_DS] = sc; vsz[V_DZ] = nc;
  vsz[V_RX] = n; vsz[V_RS] = sc; vsz[V_EX] = n; vsz[V_ES] = sc;
  vsz[V_DUAL] = n; vsz[V_SE] = ne + nc + 2; vsz[V_RUP] = 0; vsz[V_SI] = nc; // [A x; C x] contiguous
  vsz[V_HDX] = n; vsz[V_ADX] = ne + nc + 2; vsz[V_ATDY] = n; vsz[V_CDX] = 0; vsz[V_CTDZ] = n; vsz[V_Q] = n; // [A dx; C dx] contiguous
  vsz[V_GS] = n; vsz[V_BS] = ne; vsz[V_US] = nc; vsz[V_LS] = nc; vsz[V_IS] = d.box ? n : 2; vsz[V_DELTA] = n + ne + nc;
  vsz[V_B] = ne; vsz[V_U] = nc; vsz[V_L] = nc;
  vsz[V_D1INV] = 2; vsz[V_DSV] = 2; vsz[V_DSINV] = 2;
  vsz[V_T1] = n; vsz[V_T2] = n; vsz[V_T3] = n;
  vsz[V_S1] = sc; vsz[V_S2] = sc; vsz[V_S3] = sc; vsz[V_S4] = sc;
  vsz[V_ALPHAS] = 2 * nc + 2; vsz[V_GRADS] = 4;
  const int uv_ld = (std::max(n, si_cap) + 2) & ~1;
  vsz[V_SCRATCH] = std::max(std::max(8 * uv_ld, PQP_NW * 128), PQP_NW * (std::max(n, ne + nc) + 2));
  vsz[V_RED] = PQP_NW * 16;
  vsz[V_KT] = ne + nc + 4;
  vsz[V_KT2] = ne + nc + 4;
  // Arena diet (every double saved here is S^-1 capacity: cfg 2 needs 113 slots for seeds 0..4095, 111 fitted):
  //   s4 is unused by the tile kernel; t3 (C^T z of the global passes) and q (per Newton step) are never live together;
  //   kt2 (second coefficient vector of the global passes) and alphas (line search breakpoints) neither.
  const bool alias_kt2 = vsz[V_ALPHAS] >= vsz[V_KT2];
  vsz[V_S4] = 0;
  vsz[V_T3] = 0;
  if (alias_kt2) vsz[V_KT2] = 0;
  int off = 0;
  for (int v = 0; v < V_COUNT; ++v) {
    L.voff[v] = off;
    off += (int)rnd(vsz[v]);
  }
  L.voff[V_S4] = L.voff[V_S3];
  L.voff[V_T3] = L.voff[V_Q];
  if (alias_kt2) L.voff[V_KT2] = L.voff[V_ALPHAS];
  L.voff[V_RUP] = L.voff[V_SE] + ne;
  L.voff[V_CDX] = L.voff[V_ADX] + ne;
  L.vec_doubles = off;
  L.scratch_doubles = vsz[V_SCRATCH];
  L.si_cap = si_cap;
  L.ctas_per_sm = ctas;
  L.kind = 1;
  const int ldn = (n + 1) & ~1, ldb = (ne + nc + 1) & ~1;
  int64_t sz[PA_COUNT];
  sz[PA_M1] = rnd((int64_t)n * ldn + 4);
  sz[PA_AS] = rnd((int64_t)n * ldb + 4);              // Bt
  sz[PA_MS] = rnd(tile_doubles(si_cap));
  sz[PA_G] = rnd((int64_t)(ne + nc) * ldb + 4);       // G, full square
  sz[PA_Y] = rnd((int64_t)n * ldb + 4);               // W
  sz[PA_VEC] = L.vec_doubles;
  const int64_t nlist = std::max(nc, cap);
  L.smem_int_bytes = (int32_t)((4 * (nc + cap + nlist + nc + 2 * PQP_NW + 8) + 2 * nc + 15) & ~15);
  const int64_t budget = budget_bytes - 1664 /*static shared memory of the kernel (tile kernel: 736 B plain, 1616 B in the fused instantiation); the budget must hold for the FUSED instantiation too: round 2 found it running at one CTA per SM, 16 bytes over half an SM*/ - L.smem_int_bytes;
  int64_t smem_d = 0, ws_d = 0;
  auto put = [&](int id, bool want_smem) {
    if (want_smem && (smem_d + sz[id]) * 8 <= budget) {
      L.in_smem[id] = 1;
      L.off[id] = smem_d;
      smem_d += sz[id];
    } else {
      L.in_smem[id] = 0;
      L.off[id] = ws_d;
      ws_d += sz[id];
    }
  };
  put(PA_VEC, true);
  put(PA_MS, true);
  put(PA_M1, pi_smem);
  put(PA_AS, false);
  put(PA_G, false);
  put(PA_Y, false);
  L.smem_doubles = (int32_t)smem_d;
  L.ws_doubles = std::max<int64_t>(ws_d, 2);
  return (L.in_smem[PA_VEC] && L.in_smem[PA_MS] && (!pi_smem || L.in_smem[PA_M1])) ? 0 : 1;
}

// Big layout (kind 2) of the BIG variant of the tile body (pqp_kernels.cu, namespace bigk): pass scratch, reduction
// scratch and the two coefficient vectors in shared memory (absolute offsets in voff), then the vector arena if it
// fits (else in the workspace); packed S^-1 (order max(n, capacity): P is inverted there as well), P^-1 (n x ldn),
// Bt, W (n x ldb) and G (m x ldb) in the per-CTA global workspace. No capacity limit, no overflow retries.
int
fill_layout_big(const PqpDims& d, PqpLayout& L, int64_t budget_bytes, int ctas)
{
  std::memset(&L, 0, sizeof(L));
  const int n = d.n, ne = d.ne, nc = d.nc, cap = d.cap;
  auto rnd = [](int64_t v) { return (v + 1) & ~int64_t(1); };
  const int m = ne + nc, sc = cap + 2;
  int vsz[V_COUNT];
  for (int& v : vsz) v = 0;
  vsz[V_X] = n; vsz[V_Y] = ne; vsz[V_Z] = nc; vsz[V_XP] = n; vsz[V_YP] = ne; vsz[V_ZP] = nc;
  vsz[V_DX] = n; vsz[V_DS] = sc; vsz[V_DZ] = nc;
  vsz[V_RX] = n; vsz[V_RS] = sc; vsz[V_EX] = n; vsz[V_ES] = sc;
  vsz[V_DUAL] = n; vsz[V_SE] = m + 2; vsz[V_RUP] = 0; vsz[V_SI] = nc;
  vsz[V_HDX] = n; vsz[V_ADX] = m + 2; vsz[V_ATDY] = n; vsz[V_CDX] = 0; vsz[V_CTDZ] = n; vsz[V_Q] = n;
  vsz[V_GS] = n; vsz[V_BS] = ne; vsz[V_US] = nc; vsz[V_LS] = nc; vsz[V_IS] = d.box ? n : 2; vsz[V_DELTA] = n + ne + nc;
  // unscaled b, u, l: read from the model arrays when there are no box constraints (u, l of a box QP are concatenations)
  vsz[V_B] = d.box ? ne : 0; vsz[V_U] = d.box ? nc : 0; vsz[V_L] = d.box ? nc : 0;
  vsz[V_D1INV] = (d.hess == PQP_HESSIAN_DENSE) ? 2 : n; vsz[V_DSV] = 2; vsz[V_DSINV] = 2;
  vsz[V_T1] = n; vsz[V_T2] = n; vsz[V_T3] = 0; // t3 shares q (never live together, see fill_layout_tile)
  vsz[V_S1] = sc; vsz[V_S2] = sc; vsz[V_S3] = sc; vsz[V_S4] = 0;
  vsz[V_ALPHAS] = std::max(2 * nc + 2, m + 4); vsz[V_GRADS] = 4; // the second coefficient vector kt2 shares alphas
  // shared-memory part (absolute offsets)
  const int ord = std::max(n, cap);
  const int uv_ld = (ord + 2) & ~1;
  const int scratch = std::max(8 * uv_ld, PQP_NW * (std::max(ord, m) + 2));
  int64_t sm = 0;
  L.voff[V_SCRATCH] = (int32_t)sm; sm += rnd(scratch);
  L.voff[V_RED] = (int32_t)sm;     sm += rnd(PQP_NW * 16);
  L.voff[V_KT] = (int32_t)sm;      sm += rnd(m + 4);
  L.scratch_doubles = scratch;
  // vector arena (relative offsets)
  int off = 0;
  for (int v = 0; v < V_COUNT; ++v) {
    if (v == V_SCRATCH || v == V_RED || v == V_KT || v == V_KT2) continue;
    L.voff[v] = off;
    off += (int)rnd(vsz[v]);
  }
  L.voff[V_S4] = L.voff[V_S3];
  L.voff[V_T3] = L.voff[V_Q];
  L.voff[V_RUP] = L.voff[V_SE] + ne;
  L.voff[V_CDX] = L.voff[V_ADX] + ne;
  L.vec_doubles = off;
  L.si_cap = cap;
  L.ctas_per_sm = ctas;
  L.kind = 2;
  const int ldn = (n + 1) & ~1, ldb = (m + 1) & ~1;
  int64_t sz[PA_COUNT];
  sz[PA_M1] = (d.hess == PQP_HESSIAN_DENSE) ? rnd((int64_t)n * ldn + 4) : 2;
  sz[PA_AS] = rnd((int64_t)n * ldb + 4);                 // Bt
  // packed S^-1 (and P during its inversion); sized for the fallback's K^-1 of order n + capacity (pqp_fast_body.inl, kkt_factor)
  const int64_t ordk = (int64_t)n + cap;
  // (rows padded to an even length, pqp_fast_body.inl ts_idx: order m takes ((m + 1) / 2) (m + 2 - m % 2) doubles)
  auto packed = [](int64_t mm) { return ((mm + 1) / 2) * (mm + 2 - (mm & 1)); };
  sz[PA_MS] = rnd(std::max<int64_t>(packed(ord), packed(ordk)) + 4);
  sz[PA_G] = rnd((int64_t)m * ldb + 4);                  // G, full square (lower block triangle used)
  // W; later the fallback's panels (8), right-hand side, solution and NW partial vectors of length n + capacity
  sz[PA_Y] = rnd(std::max<int64_t>((int64_t)n * ldb, (int64_t)(8 + 2 + PQP_NW) * ((ordk + 2) & ~int64_t(1))) + 4);
  sz[PA_VEC] = L.vec_doubles;
  const int64_t nlist = std::max(nc, cap);
  L.smem_int_bytes = (int32_t)((4 * (nc + cap + nlist + nc + 2 * PQP_NW + 8) + 2 * nc + 15) & ~15);
  const int64_t budget = budget_bytes - 1664 /*static shared memory of the kernel (752 B plain, 1632 B in the fused instantiation; cuobjdump's SHARED figure adds the 1 KB system reserve); the budget must hold for the FUSED instantiation too: round 2 found it running at one CTA per SM, 16 bytes over half an SM*/ - L.smem_int_bytes;
  if (sm * 8 > budget) return 1; // not even the scratch fits
  int64_t smem_d = sm, ws_d = 0;
  if ((smem_d + sz[PA_VEC]) * 8 <= budget) {
    L.in_smem[PA_VEC] = 1;
    L.off[PA_VEC] = smem_d;
    smem_d += sz[PA_VEC];
  } else {
    L.in_smem[PA_VEC] = 0;
    L.off[PA_VEC] = ws_d;
    ws_d += sz[PA_VEC];
  }
  for (int id : { (int)PA_MS, (int)PA_M1, (int)PA_AS, (int)PA_G, (int)PA_Y }) {
    L.in_smem[id] = 0;
    L.off[id] = ws_d;
    ws_d += sz[id];
  }
  L.smem_doubles = (int32_t)smem_d;
  L.ws_doubles = std::max<int64_t>(ws_d, 2);
  return 0;
}

int64_t
sym_doubles(int64_t m)
{
  return m * (m + 1) / 2 + 2;
}

int
make_layout(pqp_batch* b)
{
  const PqpDims& d = b->d;
  int max_smem = pqp_solve_max_smem();
  if (max_smem <= 0) return fail(PQP_ECUDA, "no CUDA device / cannot query shared memory");
  int smem_sm = 0;
  cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, b->device);
  if (smem_sm <= 0) smem_sm = max_smem + 1024;
  const char* mode = std::getenv("PQP_LAYOUT"); // "compact" | "full" | "generic" | unset (auto)
  const std::string m = mode ? mode : "auto";
  // generic fallback first (always valid as long as the vectors fit somewhere)
  fill_layout(d, b->lay_gen, max_smem, false, false, false, d.cap, 1);
  bool done = false;
  // tile layout (specialised kernel): dense Hessian, n even and <= 128, at most 254 constraint rows
  if ((m == "auto" || m == "tile") && d.hess == PQP_HESSIAN_DENSE && (d.n % 2) == 0 && d.n <= 128 && d.n >= 2 && d.ne + d.nc <= 254 && d.nc > 0) {
    // experiment hook: PQP_TILE_CTAS=1 -> one CTA per SM with P^-1 in shared memory as well
    const char* tce = std::getenv("PQP_TILE_CTAS");
    const int tctas = (tce && std::atoi(tce) == 1) ? 1 : 2;
    const bool pis = tctas == 1;
    const int64_t per_cta = tctas == 1 ? (int64_t)max_smem : ((int64_t)smem_sm - 2 * 1024) / 2;
    int best = 0;
    bool forced_tile_cap = false;
    for (int cnd = std::min(std::max(d.cap, d.n), 128); cnd >= std::max(d.n, d.ne + 1); --cnd) { // P is inverted inside the S^-1 storage: cap >= n
      PqpLayout probe;
      if (fill_layout_tile(d, probe, per_cta, cnd, tctas, pis) == 0) {
        best = cnd;
        break;
      }
    }
    if (const char* e = std::getenv("PQP_SI_CAP")) { // test hook: force a small capacity to exercise the retry path
      int v = std::atoi(e);
      if (v >= std::max(d.n, d.ne + 1) && v < best) {
        best = v;
        forced_tile_cap = true;
      }
    }
    // worth it only if the shared-memory S^-1 holds the equalities plus half of the inequality rows
    // (cfg 3, box constraints: ~all QPs exceed 128 slots and would go through the retry levels)
    const int need = d.ne + std::min(d.nc, std::max(8, (d.nc + 1) / 2));
    if (best > 0 && (best >= std::min(d.cap, need) || forced_tile_cap)) {
      fill_layout_tile(d, b->lay, per_cta, best, tctas, pis);
      done = true;
      // first retry level for QPs whose active set outgrows `best`: same kernel, one CTA per SM
      std::memset(&b->lay_big, 0, sizeof(b->lay_big));
      for (int cnd = std::min(std::max(d.cap, d.n), 128); cnd > best; --cnd) {
        PqpLayout probe;
        if (fill_layout_tile(d, probe, (int64_t)max_smem, cnd, 1, false) == 0) {
          b->lay_big = probe;
          break;
        }
      }
    }
  }
  // big layout (BIG variant of the tile body): everything the tile kernel proper cannot hold, n even
  if (!done && (m == "auto" || m == "big") && (d.n % 2) == 0 && d.n >= 2 && d.nc > 0 && !std::getenv("PQP_NO_BIG")) {
    const int64_t half = ((int64_t)smem_sm - 2 * 1024) / 2;
    PqpLayout two, one;
    const int r2 = fill_layout_big(d, two, half, 2);           // 0: the shared-memory scratch part fits twice per SM
    const int r1 = fill_layout_big(d, one, (int64_t)max_smem, 1);

    // two CTAs per SM when the vector arena fits in half an SM as well - or when it would not fit in a whole one either
    if (r2 == 0 && (two.in_smem[PA_VEC] || r1 != 0 || !one.in_smem[PA_VEC])) {
      b->lay = two;
      done = true;
    } else if (r1 == 0) {
      b->lay = one;
      done = true;
    }
  }
  if (!done && (m == "auto" || m == "compact")) {
    // two CTAs per SM: each gets half of the SM's shared memory minus the 1 KB system reserve
    const int64_t per_cta = ((int64_t)smem_sm - 2 * 1024) / 2;
    PqpLayout probe;
    fill_layout(d, probe, per_cta, false, false, false, 1, 2); // vectors only, to measure them
    const int64_t left = (per_cta - 1024 - probe.smem_int_bytes) / 8 - probe.vec_doubles;
    int cap2 = 0;
    for (int cnd = d.cap; cnd >= 1; --cnd) {
      if (sym_doubles(cnd) + 2 <= left) {
        cap2 = cnd;
        break;
      }
    }
    bool forced_cap = false;
    if (const char* e = std::getenv("PQP_SI_CAP")) { // test hook: force a small capacity to exercise the retry path
      int v = std::atoi(e);
      if (v >= d.ne + 1 && v < cap2) {
        cap2 = v;
        forced_cap = true;
      }
    }
    // worth it only if the shared-memory S^-1 holds the equalities plus a healthy share of the
    // inequalities, and P^-1 can be swept inside that region
    const int need = d.ne + std::min(d.nc, std::max(8, (3 * d.nc + 3) / 4));
    const bool pi_ok = d.hess != PQP_HESSIAN_DENSE || sym_doubles(d.n) <= sym_doubles(cap2);
    if (probe.in_smem[PA_VEC] && (cap2 >= std::min(d.cap, need) || forced_cap) && pi_ok) {
      fill_layout(d, b->lay, per_cta, false, true, false, cap2, 2);
      done = b->lay.in_smem[PA_MS] != 0;
    }
  }
  if (!done && (m == "auto" || m == "full" || m == "compact")) {
    fill_layout(d, b->lay, max_smem, true, true, true, d.cap, 1);
    done = b->lay.in_smem[PA_VEC] && b->lay.in_smem[PA_MS] && b->lay.in_smem[PA_M1];
  }
  if (!done) b->lay = b->lay_gen;
  return 0;
}

int
check_range(pqp_batch* b, int64_t first, int64_t count)
{
  if (!b) return fail(PQP_EINVAL, "null batch");
  if (first < 0 || count < 0 || first + count > b->B) return fail(PQP_EINVAL, "wrong argument size: QP index range out of bounds");
  return 0;
}

int
copy_in(pqp_batch* b, double* dst_base, const double* src, int64_t first, int64_t count, int64_t per_qp, bool src_is_device, cudaStream_t st = nullptr)
{
  if (!src || per_qp == 0 || count == 0) return 0;
  CUDA_TRY(cudaMemcpyAsync(dst_base + first * per_qp, src, sizeof(double) * (size_t)(count * per_qp), src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st ? st : b->stream));
  return 0;
}

// helpers.hpp:680-705
void
update_proximal_parameters(pqp_batch* b, int64_t i, const double* rho, const double* mu_eq, const double* mu_in)
{
  pqp_settings& s = b->hparams[i].s;
  pqp_info& info = b->hinfo[i];
  QpFlags& f = b->flags[i];
  if (rho) {
    s.default_rho = *rho;
    info.rho = *rho;
    f.proximal_parameter_update = true;
  }
  if (mu_eq) {
    s.default_mu_eq = *mu_eq;
    info.mu_eq = *mu_eq;
    info.mu_eq_inv = 1.0 / info.mu_eq;
    f.proximal_parameter_update = true;
  }
  if (mu_in) {
    s.default_mu_in = *mu_in;
    info.mu_in = *mu_in;
    info.mu_in_inv = 1.0 / info.mu_in;
    f.proximal_parameter_update = true;
  }
}
// helpers.hpp:176-189
void
update_default_rho(pqp_batch* b, int64_t i, const double* manual)
{
  pqp_settings& s = b->hparams[i].s;
  pqp_info& info = b->hinfo[i];
  if (manual) {
    s.default_H_eigenvalue_estimate = *manual;
    info.minimal_H_eigenvalue_estimate = s.default_H_eigenvalue_estimate;
  }
  s.default_rho += std::fabs(info.minimal_H_eigenvalue_estimate);
  info.rho = s.default_rho;
}

// zero x, y, z, se, si of the QPs [i, i + cnt)
int
zero_results(pqp_batch* b, int64_t i, int64_t cnt = 1)
{
  const PqpDims& d = b->d;
  if (cnt <= 0) return 0;
  CUDA_TRY(cudaMemsetAsync(b->p.x + i * d.n, 0, sizeof(double) * d.n * cnt, b->stream));
  if (d.ne) CUDA_TRY(cudaMemsetAsync(b->p.y + i * d.ne, 0, sizeof(double) * d.ne * cnt, b->stream));
  if (d.ne) CUDA_TRY(cudaMemsetAsync(b->p.se + i * d.ne, 0, sizeof(double) * d.ne * cnt, b->stream));
  if (d.nc) CUDA_TRY(cudaMemsetAsync(b->p.z + i * d.nc, 0, sizeof(double) * d.nc * cnt, b->stream));
  if (d.nc) CUDA_TRY(cudaMemsetAsync(b->p.si + i * d.nc, 0, sizeof(double) * d.nc * cnt, b->stream));
  return 0;
}
// the same for the QPs flagged in need[0 .. count): one memset per array and contiguous run
int
zero_results_runs(pqp_batch* b, int64_t first, const std::vector<char>& need)
{
  const int64_t count = (int64_t)need.size();
  int64_t k = 0;
  while (k < count) {
    if (!need[k]) {
      ++k;
      continue;
    }
    int64_t e = k;
    while (e < count && need[e]) ++e;
    if (int rc = zero_results(b, first + k, e - k)) return rc;
    k = e;
  }
  return 0;
}

// helpers.hpp:522-572: what setup() does to results / workspace flags
int
setup_results_and_flags(pqp_batch* b, int64_t i, char* need_zero = nullptr)
{
  pqp_settings& s = b->hparams[i].s;
  pqp_info& info = b->hinfo[i];
  QpFlags& f = b->flags[i];
  auto work_cleanup = [&]() { f = QpFlags(); };
  switch (s.initial_guess) {
    case PQP_EQUALITY_CONSTRAINED_INITIAL_GUESS:
    case PQP_NO_INITIAL_GUESS:
    case PQP_WARM_START: {
      bool ppu = f.proximal_parameter_update;
      if (need_zero)
        *need_zero = 1; // the caller zeroes whole runs of QPs with one memset per array
      else if (int rc = zero_results(b, i))
        return rc;
      if (ppu)
        cleanup_statistics(info);
      else
        cold_start(info, &s, b->backend);
      work_cleanup();
    } break;
    case PQP_COLD_START_WITH_PREVIOUS_RESULT: {
      if (f.proximal_parameter_update)
        cleanup_statistics(info);
      else
        cold_start(info, &s, b->backend);
      work_cleanup();
    } break;
    case PQP_WARM_START_WITH_PREVIOUS_RESULT: {
      if (f.refactorize || f.proximal_parameter_update) {
        work_cleanup();
        f.refactorize = true;
      }
      cleanup_statistics(info);
    } break;
    default:
      return fail(PQP_EINVAL, "invalid initial_guess");
  }
  return 0;
}

int
launch_setup(pqp_batch* b, int64_t first, int64_t count, bool execute, bool reset_scaling, bool time_begin = true, bool time_end = true, cudaStream_t st = nullptr)
{
  if (!st) st = b->stream;
  // device needs the settings (preconditioner parameters) of these QPs
  CUDA_TRY(cudaMemcpyAsync(b->p.params + first, b->hparams.data() + first, sizeof(PqpQpParams) * (size_t)count, cudaMemcpyHostToDevice, st));
  PqpSetupArgs a{};
  a.d = b->d;
  a.p = b->p;
  a.first = (int32_t)first;
  a.count = (int32_t)count;
  a.execute = execute ? 1 : 0;
  a.reset_scaling = reset_scaling ? 1 : 0;
  if (time_begin) CUDA_TRY(cudaEventRecord(b->ev0, st));
  int rc = pqp_launch_setup(&a, st);
  if (rc != 0) return fail(PQP_ECUDA, std::string("setup kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
  if (time_end) CUDA_TRY(cudaEventRecord(b->ev1, st));
  b->setup_timed = true;
  b->launches += 1;
  return 0;
}

// later work on the main stream is ordered after everything enqueued on the chunk streams
int
join_cstreams(pqp_batch* b)
{
  for (int j = 0; j < PQP_CSTREAMS; ++j) {
    CUDA_TRY(cudaEventRecord(b->ev_cdone[j], b->cstream[j]));
    CUDA_TRY(cudaStreamWaitEvent(b->stream, b->ev_cdone[j], 0));
  }
  return 0;
}

// A deferred set-up that the next call cannot fuse (anything but a solve of the whole batch): run it now.
int
flush_deferred(pqp_batch* b)
{
  if (!b->deferred) return 0;
  const int code = b->deferred;
  b->deferred = 0;
  if (b->deferred_gated) {
    CUDA_TRY(cudaEventRecord(b->ev_feed, b->copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(b->stream, b->ev_feed, 0));
    b->deferred_gated = false;
  }
  return launch_setup(b, 0, b->B, (code & 1) != 0, (code & 4) != 0);
}

int
do_init(pqp_batch* b, int64_t first, int64_t count, const double* H, const double* g, const double* A, const double* b_, const double* C, const double* l, const double* u, const double* l_box, const double* u_box, int compute_preconditioner,
        const double* rho, const double* mu_eq, const double* mu_in, const double* manual_eig, bool dev_ptrs)
{
  if (int rc = check_range(b, first, count)) return rc;
  const PqpDims& d = b->d;
  if (!d.box && (l_box || u_box))
    return fail(PQP_EINVAL, "wrong model setup: the QP object is designed without box constraints, but is initialized with lower or upper box inequalities.");
  CUDA_TRY(cudaSetDevice(b->device));
  std::vector<char> need_zero((size_t)count, 0);
  for (int64_t i = first; i < first + count; ++i) {
    pqp_settings& s = b->hparams[i].s;
    QpFlags& f = b->flags[i];
    s.compute_preconditioner = compute_preconditioner ? 1 : 0;
    f.refactorize = (s.initial_guess == PQP_WARM_START_WITH_PREVIOUS_RESULT); // wrapper.hpp:452-459
    f.proximal_parameter_update = false;
    update_proximal_parameters(b, i, rho, mu_eq, mu_in);
    update_default_rho(b, i, manual_eig);
    if (int rc = setup_results_and_flags(b, i, &need_zero[(size_t)(i - first)])) return rc;
    f.is_initialized = true;
  }
  if (int rc = zero_results_runs(b, first, need_zero)) return rc;
  const int64_t n = d.n, ne = d.ne, ni = d.ni;
  const bool whole = first == 0 && count == b->B && count > 0;
  const char* mode_env = std::getenv("PQP_E2E"); // "fused" (default) | "chunks" (per-chunk set-up + solve launches) | "plain"
  const std::string mode = mode_env ? mode_env : "fused";
  auto upload = [&](int64_t f, int64_t cnt, cudaStream_t st) -> int {
    const int64_t o = f - first;
    auto at = [&](const double* src, int64_t per) { return src ? src + o * per : nullptr; };
    if (int rc = copy_in(b, b->p.H, at(H, n * n), f, cnt, n * n, dev_ptrs, st)) return rc;
    if (int rc = copy_in(b, b->p.g, at(g, n), f, cnt, n, dev_ptrs, st)) return rc;
    if (int rc = copy_in(b, b->p.A, at(A, ne * n), f, cnt, ne * n, dev_ptrs, st)) return rc;
    if (int rc = copy_in(b, b->p.b, at(b_, ne), f, cnt, ne, dev_ptrs, st)) return rc;
    if (int rc = copy_in(b, b->p.C, at(C, ni * n), f, cnt, ni * n, dev_ptrs, st)) return rc;
    if (int rc = copy_in(b, b->p.l, at(l, ni), f, cnt, ni, dev_ptrs, st)) return rc;
    if (int rc = copy_in(b, b->p.u, at(u, ni), f, cnt, ni, dev_ptrs, st)) return rc;
    if (d.box) {
      if (int rc = copy_in(b, b->p.l_box, at(l_box, n), f, cnt, n, dev_ptrs, st)) return rc;
      if (int rc = copy_in(b, b->p.u_box, at(u_box, n), f, cnt, n, dev_ptrs, st)) return rc;
    }
    return 0;
  };
  b->nchunks_pending = 0;
  if (whole && b->fused_ok && mode == "fused") {
    // Fused feed: upload only. The persistent solve kernel equilibrates each QP right before solving it and,
    // for host inputs, starts while later chunks are still crossing PCIe (gated on d_ready).
    if (b->deferred && b->deferred_gated) { // an unconsumed feed: its uploads must land before the new ones
      CUDA_TRY(cudaEventRecord(b->ev_feed, b->copy_stream));
      CUDA_TRY(cudaStreamWaitEvent(b->stream, b->ev_feed, 0));
    }
    b->deferred = compute_preconditioner ? 1 : (2 | 4);
    b->deferred_gated = false;
    if (!dev_ptrs && count >= 256) {
      const int nchunks = (int)std::min<int64_t>(PQP_FEED_CHUNKS, count / 64);
      CUDA_TRY(cudaEventRecord(b->ev_main, b->stream)); // uploads must not overtake kernels still reading the buffers
      CUDA_TRY(cudaStreamWaitEvent(b->copy_stream, b->ev_main, 0));
      CUDA_TRY(cudaMemsetAsync(b->d_ready, 0, 2 * sizeof(int32_t), b->copy_stream));
      CUDA_TRY(cudaEventRecord(b->ev_feed, b->copy_stream));
      CUDA_TRY(cudaStreamWaitEvent(b->stream, b->ev_feed, 0)); // no kernel may see the progress word of the previous feed
      for (int k = 0; k < nchunks; ++k) {
        const int64_t f = count * k / nchunks, e = count * (k + 1) / nchunks;
        if (int rc = upload(f, e - f, b->copy_stream)) return rc;
        b->h_ready[k] = (int32_t)e;
        CUDA_TRY(cudaMemcpyAsync(b->d_ready, b->h_ready + k, sizeof(int32_t), cudaMemcpyHostToDevice, b->copy_stream));
      }
      b->deferred_gated = true;
    } else if (int rc = upload(first, count, b->stream)) {
      return rc;
    }
    return 0;
  }
  if (int rc = flush_deferred(b)) return rc;
  // Host inputs of a large range are uploaded in chunks on a second stream; the set-up kernel of
  // chunk k runs while chunk k+1 is still crossing PCIe.
  const int nchunks = (!dev_ptrs && count >= 256 && mode != "plain") ? (int)std::min<int64_t>(PQP_UPLOAD_CHUNKS, count / 128) : 1;
  if (nchunks > 1) {
    CUDA_TRY(cudaEventRecord(b->ev_main, b->stream)); // uploads must not overtake kernels still reading the buffers
    CUDA_TRY(cudaStreamWaitEvent(b->copy_stream, b->ev_main, 0));
    for (int j = 0; j < PQP_CSTREAMS; ++j) CUDA_TRY(cudaStreamWaitEvent(b->cstream[j], b->ev_main, 0));
  }
  for (int k = 0; k < nchunks; ++k) {
    const int64_t f = first + count * k / nchunks, e = first + count * (k + 1) / nchunks, cnt = e - f;
    if (int rc = upload(f, cnt, nchunks > 1 ? b->copy_stream : b->stream)) return rc;
    cudaStream_t cs = b->stream;
    if (nchunks > 1) {
      cs = b->cstream[k % PQP_CSTREAMS];
      CUDA_TRY(cudaEventRecord(b->ev_chunk[k], b->copy_stream));
      CUDA_TRY(cudaStreamWaitEvent(cs, b->ev_chunk[k], 0));
      b->chunk_first[k] = f;
      b->chunk_count[k] = cnt;
    }
    if (int rc = launch_setup(b, f, cnt, compute_preconditioner != 0, compute_preconditioner == 0, k == 0, k == nchunks - 1, cs)) return rc;
  }
  if (nchunks > 1) {
    if (int rc = join_cstreams(b)) return rc;
    // the whole batch, freshly initialised: a solve() issued next may run chunk by chunk
    if (whole && b->ws_slot_doubles > 0) b->nchunks_pending = nchunks;
  }
  return 0;
}

void
fill_vec(std::vector<double>& v, double val)
{
  std::fill(v.begin(), v.end(), val);
}

// upload the per-QP launch parameters and enqueue one persistent solve kernel
int
enqueue_solve(pqp_batch* b, cudaStream_t st, const PqpLayout& lay, int grid, int64_t first = 0, int64_t count = -1, int slot = 0, bool timed = true, int fused_code = 0, int32_t* ready = nullptr)
{
  if (count < 0) count = b->B;
  CUDA_TRY(cudaMemcpyAsync(b->p.params + first, b->hparams.data() + first, sizeof(PqpQpParams) * (size_t)count, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(b->counter + slot, 0, sizeof(int32_t), st));
  PqpSolveArgs a{};
  a.d = b->d;
  a.p = b->p;
  a.lay = lay;
  a.batch = (int32_t)count;
  a.first = (int32_t)first;
  a.counter = b->counter + slot;
  a.ws = b->ws + (size_t)slot * (size_t)b->ws_slot_doubles;
  a.dbg = b->dbg;
  a.dbg_cap = b->dbg_cap;
  a.dbg_qp = 0;
  a.prof = b->prof;
  if (const char* e = std::getenv("PQP_DEBUG_TRACE")) a.dbg_qp = std::atoi(e);
  if (const char* e = std::getenv("PQP_WATCHDOG_MS")) a.watchdog_ns = 1000000ull * (unsigned long long)std::atoll(e);
  if (const char* e = std::getenv("PQP_FORCE_KKT")) a.force_kkt = std::atoi(e);
  // big variant: the per-CTA workspaces of the large shapes (cfg 4: 4.9 MB, cfg 5: 27 MB, times 296 CTAs) live in HBM;
  // the streaming passes prefetch the rows of the next warp iteration into L2. Measured (profiles/r02_summary.md):
  // distance 1: cfg 4 +6.7 %, cfg 5 +6 %; 2 and 4: less; cfg 3 (1 MB per CTA, mostly cache resident): -2 % -> off there.
  // PQP_PREFETCH=<distance> overrides (0 = off).
  a.prefetch = (lay.kind == 2 && lay.ws_doubles >= 250000) ? 1 : 0;
  if (const char* e = std::getenv("PQP_PREFETCH")) a.prefetch = std::atoi(e);
  a.fused_setup = fused_code;
  a.ready = ready;
  {
    // smallest per-QP array the upload writes, in bytes: QPs that close together can share a cache line
    int64_t mn = (int64_t)b->d.n;
    if (b->d.ne > 0) mn = std::min<int64_t>(mn, b->d.ne);
    if (b->d.ni > 0) mn = std::min<int64_t>(mn, b->d.ni);
    a.feed_margin = (int32_t)((128 + 8 * mn - 1) / (8 * mn));
  }
#ifndef PQP_CPU_EMU
  // L2 residency experiment (PQP_L2_PERSIST=1; OFF by default): an access-policy window that marks the per-CTA workspace
  // (P^-1, Bt, G: ~300 KB x 296 CTAs at cfg 2) as persisting in L2 and everything else as streaming. Measured (round 2,
  // profiles/r02_summary.md section 6): no effect on the tile kernel (31.4 vs 31.5 ms, DRAM traffic 9.1 GB either way)
  // and 20 % SLOWER on the large shapes, whose multi-GB workspaces the window cannot cover.
  {
    static const bool persist = std::getenv("PQP_L2_PERSIST") && std::atoi(std::getenv("PQP_L2_PERSIST")) == 1;
    if (persist && b->ws && b->l2_window_bytes > 0) {
      cudaStreamAttrValue at{};
      at.accessPolicyWindow.base_ptr = (void*)a.ws;
      at.accessPolicyWindow.num_bytes = (size_t)std::min<int64_t>(b->l2_window_bytes, (int64_t)grid * lay.ws_doubles * 8);
      at.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)b->l2_persist_bytes / (double)std::max<size_t>(at.accessPolicyWindow.num_bytes, 1));
      at.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      at.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &at); // best effort: an error here must not fail the solve
      (void)cudaGetLastError();
    }
  }
#endif
  if (timed) CUDA_TRY(cudaEventRecord(b->ev2, st));
  int rc = pqp_launch_solve(&a, grid, st);
  if (rc != 0) return fail(PQP_ECUDA, std::string("solve kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
  if (timed) CUDA_TRY(cudaEventRecord(b->ev3, st));
  return 0;
}

} // namespace

extern "C" {

const char*
pqp_last_error(void)
{
  return g_err.c_str();
}
const char*
pqp_version(void)
{
  return "proxsuite_b200 0.1.0 (sm_100a)";
}

void
pqp_settings_default(pqp_settings* s, int dense_backend)
{
  // settings.hpp:213-315
  std::memset(s, 0, sizeof(*s));
  s->default_rho = (dense_backend == PQP_BACKEND_PRIMAL_LDLT) ? 1e-5 : 1e-6;
  s->default_mu_eq = 1e-3;
  s->default_mu_in = 1e-1;
  s->alpha_bcl = 0.1;
  s->beta_bcl = 0.9;
  s->refactor_dual_feasibility_threshold = 1e-2;
  s->refactor_rho_threshold = 1e-7;
  s->mu_min_eq = 1e-9;
  s->mu_min_in = 1e-8;
  s->mu_max_eq_inv = 1e9;
  s->mu_max_in_inv = 1e8;
  s->mu_update_factor = 0.1;
  s->mu_update_inv_factor = 10;
  s->cold_reset_mu_eq = 1. / 1.1;
  s->cold_reset_mu_in = 1. / 1.1;
  s->cold_reset_mu_eq_inv = 1.1;
  s->cold_reset_mu_in_inv = 1.1;
  s->eps_abs = 1e-5;
  s->eps_rel = 0;
  s->eps_refact = 1e-6;
  s->eps_duality_gap_abs = 1e-4;
  s->eps_duality_gap_rel = 0;
  s->preconditioner_accuracy = 1e-3;
  s->eps_primal_inf = 1e-4;
  s->eps_dual_inf = 1e-4;
  s->alpha_gpdal = 0.95;
  s->default_H_eigenvalue_estimate = 0;
  s->max_iter = 10000;
  s->max_iter_in = 1500;
  s->safe_guard = 10000;
  s->nb_iterative_refinement = 10;
  s->preconditioner_max_iter = 10;
  s->frequence_infeasibility_check = 1;
  s->verbose = 0;
  s->initial_guess = PQP_EQUALITY_CONSTRAINED_INITIAL_GUESS;
  s->update_preconditioner = 0;
  s->compute_preconditioner = 1;
  s->compute_timings = 0;
  s->check_duality_gap = 0;
  s->bcl_update = 1;
  s->merit_function_type = PQP_MERIT_GPDAL;
  s->primal_infeasibility_solving = 0;
}

int
pqp_dense_backend_choice(int backend, int64_t dim, int64_t n_eq, int64_t n_in, int box)
{
  // wrapper.hpp:82-113
  if (backend != PQP_BACKEND_AUTOMATIC) return backend;
  int64_t ncons = n_in + (box ? dim : 0);
  double d = double(dim);
  double pd = 0.5 * std::pow(double(n_eq) / d, 2) + 0.17 * (std::pow(double(n_eq) / d, 3) + std::pow(double(ncons) / d, 3)) + 0.2 * std::pow(double(n_eq + ncons) / d, 2) / d;
  double pl = 1.5 * ((0.5 * double(n_eq) + double(ncons)) / d + 0.2 / d);
  return pd > pl ? PQP_BACKEND_PRIMAL_LDLT : PQP_BACKEND_PRIMAL_DUAL_LDLT;
}

pqp_batch*
pqp_batch_create(int64_t batch, int64_t dim, int64_t n_eq, int64_t n_in, int box, int hessian, int backend, int device)
{
  if (dim <= 0) {
    fail(PQP_EINVAL, "wrong argument size: the dimension wrt the primal variable x should be strictly positive.");
    return nullptr;
  }
  if (batch < 0 || n_eq < 0 || n_in < 0 || hessian < 0 || hessian > 2) {
    fail(PQP_EINVAL, "wrong argument: negative size or invalid hessian type");
    return nullptr;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fail(PQP_ECUDA, "no CUDA device available: proxsuite_b200 has no CPU fallback");
    return nullptr;
  }
  if (device < 0) cudaGetDevice(&device);
  if (cudaSetDevice(device) != cudaSuccess) {
    fail(PQP_ECUDA, "cudaSetDevice failed");
    return nullptr;
  }
  pqp_batch* b = new pqp_batch;
  b->B = batch;
  b->device = device;
  // Both dense backends are served by the same block factorisation on the
  // device; the choice only selects default_rho.
  b->backend = pqp_dense_backend_choice(backend, dim, n_eq, n_in, box);
  PqpDims& d = b->d;
  d.n = (int)dim;
  d.ne = (int)n_eq;
  d.ni = (int)n_in;
  d.box = box ? 1 : 0;
  d.nc = d.ni + (d.box ? d.n : 0);
  d.hess = hessian;
  d.cap = d.ne + d.nc;
  const size_t B = (size_t)batch;
  const size_t n = d.n, ne = d.ne, ni = d.ni, nc = d.nc;
  int rc = 0;
  PqpBatchPtrs& p = b->p;
  rc |= dev_alloc(b, &p.H, B * n * n);
  rc |= dev_alloc(b, &p.g, B * n);
  rc |= dev_alloc(b, &p.A, B * ne * n);
  rc |= dev_alloc(b, &p.b, B * ne);
  rc |= dev_alloc(b, &p.C, B * ni * n);
  rc |= dev_alloc(b, &p.l, B * ni);
  rc |= dev_alloc(b, &p.u, B * ni);
  rc |= dev_alloc(b, &p.l_box, B * n);
  rc |= dev_alloc(b, &p.u_box, B * n);
  rc |= dev_alloc(b, &p.Hs, B * n * n);
  rc |= dev_alloc(b, &p.gs, B * n);
  rc |= dev_alloc(b, &p.As, B * ne * n);
  rc |= dev_alloc(b, &p.bs, B * ne);
  rc |= dev_alloc(b, &p.Cs, B * ni * n);
  rc |= dev_alloc(b, &p.us, B * nc);
  rc |= dev_alloc(b, &p.ls, B * nc);
  rc |= dev_alloc(b, &p.is, B * n);
  rc |= dev_alloc(b, &p.delta, B * (n + ne + nc));
  rc |= dev_alloc(b, &p.c, B);
  rc |= dev_alloc(b, &p.x, B * n);
  rc |= dev_alloc(b, &p.y, B * ne);
  rc |= dev_alloc(b, &p.z, B * nc);
  rc |= dev_alloc(b, &p.se, B * ne);
  rc |= dev_alloc(b, &p.si, B * nc);
  rc |= dev_alloc(b, &p.info, B * PQP_INFO_DOUBLES);
  rc |= dev_alloc(b, &p.params, B);
  rc |= dev_alloc(b, &b->counter, PQP_CSTREAMS + 1);
  rc |= dev_alloc(b, &b->d_ready, 2);
  if (cudaHostAlloc((void**)&b->h_ready, sizeof(int32_t) * (PQP_FEED_CHUNKS + 2), cudaHostAllocDefault) != cudaSuccess) {
    b->h_ready = nullptr;
    rc |= fail(PQP_ECUDA, "cudaHostAlloc failed");
  }
  if (rc != 0) {
    pqp_batch_destroy(b);
    return nullptr;
  }
  // model defaults (model.hpp:70-91): u = +inf_bound, l = -inf_bound
  {
    const double inf_b = std::sqrt(1.7976931348623157e308);
    std::vector<double> tmp(std::max<size_t>(B * std::max(ni, n), 1));
    fill_vec(tmp, inf_b);
    cudaMemcpy(p.u, tmp.data(), sizeof(double) * B * ni, cudaMemcpyHostToDevice);
    cudaMemcpy(p.u_box, tmp.data(), sizeof(double) * B * n, cudaMemcpyHostToDevice);
    fill_vec(tmp, -inf_b);
    cudaMemcpy(p.l, tmp.data(), sizeof(double) * B * ni, cudaMemcpyHostToDevice);
    cudaMemcpy(p.l_box, tmp.data(), sizeof(double) * B * n, cudaMemcpyHostToDevice);
    std::vector<double> ones(std::max<size_t>(B * (n + ne + nc), 1), 1.0);
    cudaMemcpy(p.delta, ones.data(), sizeof(double) * B * (n + ne + nc), cudaMemcpyHostToDevice);
    cudaMemcpy(p.c, ones.data(), sizeof(double) * B, cudaMemcpyHostToDevice);
    cudaMemcpy(p.is, ones.data(), sizeof(double) * B * n, cudaMemcpyHostToDevice);
  }
  b->hparams.resize(B);
  b->hinfo.resize(B);
  b->flags.resize(B);
  for (size_t i = 0; i < B; ++i) {
    std::memset(&b->hparams[i], 0, sizeof(PqpQpParams));
    pqp_settings_default(&b->hparams[i].s, b->backend);
    std::memset(&b->hinfo[i], 0, sizeof(pqp_info));
    info_defaults(b->hinfo[i], nullptr, b->backend);
    b->hinfo[i].status = PQP_NOT_RUN;
  }
  bool aux_ok = cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreateWithFlags(&b->ev_main, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&b->ev_feed, cudaEventDisableTiming) == cudaSuccess && cudaEventCreate(&b->ev_r0) == cudaSuccess && cudaEventCreate(&b->ev_r1) == cudaSuccess;
  for (int k = 0; k < PQP_UPLOAD_CHUNKS; ++k) aux_ok = aux_ok && cudaEventCreateWithFlags(&b->ev_chunk[k], cudaEventDisableTiming) == cudaSuccess;
  for (int k = 0; k < PQP_CSTREAMS; ++k)
    aux_ok = aux_ok && cudaStreamCreateWithFlags(&b->cstream[k], cudaStreamNonBlocking) == cudaSuccess && cudaEventCreateWithFlags(&b->ev_cdone[k], cudaEventDisableTiming) == cudaSuccess;
  if (!aux_ok || cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&b->ev0) != cudaSuccess || cudaEventCreate(&b->ev1) != cudaSuccess || cudaEventCreate(&b->ev2) != cudaSuccess ||
      cudaEventCreate(&b->ev3) != cudaSuccess) {
    fail(PQP_ECUDA, "stream/event creation failed");
    pqp_batch_destroy(b);
    return nullptr;
  }
  if (make_layout(b) != 0) {
    pqp_batch_destroy(b);
    return nullptr;
  }
  // the fused feed runs the set-up inside the solve kernel, in the shared memory of the primary layout
  b->fused_ok = (int64_t)sizeof(double) * b->lay.smem_doubles >= pqp_setup_smem_bytes(d.n, d.ne, d.ni, d.nc);
  // persistent grids: resident CTAs per SM x SM count, never more than the batch
  {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    auto grid_for = [&](const PqpLayout& L) {
      int per_sm = std::max(1, (int)L.ctas_per_sm);
      per_sm = std::min(per_sm, 2048 / PQP_NT);
      if (const char* e = std::getenv("PQP_CTAS_PER_SM")) per_sm = std::max(1, std::atoi(e));
      int64_t g = (int64_t)sms * per_sm;
      return (int)std::max<int64_t>(1, std::min<int64_t>(g, std::max<int64_t>(batch, 1)));
    };
    b->grid = grid_for(b->lay);
    b->grid_gen = grid_for(b->lay_gen);
    size_t ws = std::max((size_t)b->grid * (size_t)b->lay.ws_doubles, (size_t)b->grid_gen * (size_t)b->lay_gen.ws_doubles);
    if (b->lay_big.kind == 1) {
      b->grid_big = grid_for(b->lay_big);
      ws = std::max(ws, (size_t)b->grid_big * (size_t)b->lay_big.ws_doubles);
    }
    // pipelined init + solve: up to PQP_CSTREAMS chunk launches run concurrently, each with its own workspace slice
    const int64_t nch = batch >= 256 ? std::min<int64_t>(PQP_UPLOAD_CHUNKS, batch / 128) : 1;
    if (nch > 1) {
      const int64_t cmax = (batch + nch - 1) / nch;
      b->ws_slot_doubles = std::min<int64_t>(b->grid, cmax) * b->lay.ws_doubles;
      ws = std::max(ws, (size_t)PQP_CSTREAMS * (size_t)b->ws_slot_doubles);
    }
    if (dev_alloc(b, &b->ws, ws) != 0) {
      pqp_batch_destroy(b);
      return nullptr;
    }
#ifndef PQP_CPU_EMU
    {
      int max_persist = 0, max_window = 0;
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
      cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
      // (only on request: setting the limit carves the persisting part out of the L2 for EVERY kernel of the context -
      //  measured -18 % on cfg 4 / cfg 5 with the window itself switched off)
      const bool want_persist = std::getenv("PQP_L2_PERSIST") && std::atoi(std::getenv("PQP_L2_PERSIST")) == 1;
      if (want_persist && max_persist > 0 && max_window > 0) {
        const size_t want = std::min<size_t>((size_t)max_persist, ws * sizeof(double));
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
          b->l2_persist_bytes = (int64_t)want;
          b->l2_window_bytes = (int64_t)std::min<size_t>((size_t)max_window, ws * sizeof(double));
        }
        (void)cudaGetLastError();
      }
    }
#endif
  }
  if (std::getenv("PQP_PROFILE")) dev_alloc(b, &b->prof, 16);
  if (std::getenv("PQP_DEBUG_TRACE")) {
    b->dbg_cap = 6 * 4096;
    dev_alloc(b, &b->dbg, (size_t)b->dbg_cap);
  }
  return b;
}

void
pqp_batch_destroy(pqp_batch* b)
{
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->copy_stream) cudaStreamSynchronize(b->copy_stream);
  if (b->stream) cudaStreamSynchronize(b->stream);
  for (void* p : b->allocs) cudaFree(p);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->ev2) cudaEventDestroy(b->ev2);
  if (b->ev3) cudaEventDestroy(b->ev3);
  if (b->ev_main) cudaEventDestroy(b->ev_main);
  if (b->ev_feed) cudaEventDestroy(b->ev_feed);
  if (b->ev_r0) cudaEventDestroy(b->ev_r0);
  if (b->ev_r1) cudaEventDestroy(b->ev_r1);
  if (b->h_ready) cudaFreeHost(b->h_ready);
  for (int k = 0; k < PQP_UPLOAD_CHUNKS; ++k)
    if (b->ev_chunk[k]) cudaEventDestroy(b->ev_chunk[k]);
  for (int k = 0; k < PQP_CSTREAMS; ++k) {
    if (b->cstream[k]) {
      cudaStreamSynchronize(b->cstream[k]);
      cudaStreamDestroy(b->cstream[k]);
    }
    if (b->ev_cdone[k]) cudaEventDestroy(b->ev_cdone[k]);
  }
  if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

int64_t
pqp_batch_size(const pqp_batch* b)
{
  return b ? b->B : 0;
}
int
pqp_batch_dims(const pqp_batch* b, int64_t* dim, int64_t* n_eq, int64_t* n_in, int* box, int* hessian, int* backend)
{
  if (!b) return fail(PQP_EINVAL, "null batch");
  if (dim) *dim = b->d.n;
  if (n_eq) *n_eq = b->d.ne;
  if (n_in) *n_in = b->d.ni;
  if (box) *box = b->d.box;
  if (hessian) *hessian = b->d.hess;
  if (backend) *backend = b->backend;
  return 0;
}

int
pqp_batch_settings_get(const pqp_batch* b, int64_t index, pqp_settings* out)
{
  if (!b || !out || index >= b->B) return fail(PQP_EINVAL, "bad arguments");
  *out = b->hparams[std::max<int64_t>(index, 0)].s;
  return 0;
}
int
pqp_batch_settings_set(pqp_batch* b, int64_t index, const pqp_settings* in)
{
  if (!b || !in || index >= b->B) return fail(PQP_EINVAL, "bad arguments");
  if (index < 0) {
    for (auto& p : b->hparams) p.s = *in;
  } else {
    b->hparams[index].s = *in;
  }
  return 0;
}

int
pqp_batch_init(pqp_batch* b, int64_t first, int64_t count, const double* H, const double* g, const double* A, const double* b_, const double* C, const double* l, const double* u, const double* l_box, const double* u_box, int compute_preconditioner,
               const double* rho, const double* mu_eq, const double* mu_in, const double* manual_eig)
{
  return do_init(b, first, count, H, g, A, b_, C, l, u, l_box, u_box, compute_preconditioner, rho, mu_eq, mu_in, manual_eig, false);
}
int
pqp_batch_init_device(pqp_batch* b, int64_t first, int64_t count, const double* H, const double* g, const double* A, const double* b_, const double* C, const double* l, const double* u, const double* l_box, const double* u_box,
                      int compute_preconditioner, const double* rho, const double* mu_eq, const double* mu_in, const double* manual_eig)
{
  return do_init(b, first, count, H, g, A, b_, C, l, u, l_box, u_box, compute_preconditioner, rho, mu_eq, mu_in, manual_eig, true);
}

int
pqp_batch_update(pqp_batch* b, int64_t first, int64_t count, const double* H, const double* g, const double* A, const double* b_, const double* C, const double* l, const double* u, const double* l_box, const double* u_box, int update_preconditioner,
                 const double* rho, const double* mu_eq, const double* mu_in, const double* manual_eig)
{
  if (int rc = check_range(b, first, count)) return rc;
  const PqpDims& d = b->d;
  if (!d.box && (l_box || u_box))
    return fail(PQP_EINVAL, "wrong model setup: the QP object is designed without box constraints, but the update includes lower or upper box inequalities.");
  CUDA_TRY(cudaSetDevice(b->device));
  if (int rc = flush_deferred(b)) return rc;
  b->nchunks_pending = 0; // work enqueued on the main stream from here on: the next solve is a single launch
  // wrapper.hpp:743-746: update before init == init (per QP); handle the
  // common case where the whole range is in the same state.
  bool all_init = true, none_init = true;
  for (int64_t i = first; i < first + count; ++i) {
    all_init = all_init && b->flags[i].is_initialized;
    none_init = none_init && !b->flags[i].is_initialized;
  }
  if (none_init && count > 0) return do_init(b, first, count, H, g, A, b_, C, l, u, l_box, u_box, update_preconditioner, rho, mu_eq, mu_in, nullptr, false);
  if (!all_init) return fail(PQP_ESTATE, "update on a range that mixes initialised and non-initialised QPs");
  std::vector<char> need_zero((size_t)count, 0);
  for (int64_t i = first; i < first + count; ++i) {
    pqp_settings& s = b->hparams[i].s;
    QpFlags& f = b->flags[i];
    s.update_preconditioner = update_preconditioner ? 1 : 0;
    f.refactorize = false;
    f.proximal_parameter_update = false;
    if (H || A || C) f.refactorize = true; // helpers.hpp:466-468
    update_proximal_parameters(b, i, rho, mu_eq, mu_in);
    update_default_rho(b, i, manual_eig);
    if (int rc = setup_results_and_flags(b, i, &need_zero[(size_t)(i - first)])) return rc;
    // Workspace::cleanup clears is_initialized (workspace.hpp:330-377); the
    // reference only re-sets it in qp_solve. Model data stay valid, so the
    // batch keeps the QP solvable.
    f.is_initialized = true;
  }
  if (int rc = zero_results_runs(b, first, need_zero)) return rc;
  const int64_t n = d.n, ne = d.ne, ni = d.ni;
  if (int rc = copy_in(b, b->p.H, H, first, count, n * n, false)) return rc;
  if (int rc = copy_in(b, b->p.g, g, first, count, n, false)) return rc;
  if (int rc = copy_in(b, b->p.A, A, first, count, ne * n, false)) return rc;
  if (int rc = copy_in(b, b->p.b, b_, first, count, ne, false)) return rc;
  if (int rc = copy_in(b, b->p.C, C, first, count, ni * n, false)) return rc;
  if (int rc = copy_in(b, b->p.l, l, first, count, ni, false)) return rc;
  if (int rc = copy_in(b, b->p.u, u, first, count, ni, false)) return rc;
  if (d.box) {
    if (int rc = copy_in(b, b->p.l_box, l_box, first, count, n, false)) return rc;
    if (int rc = copy_in(b, b->p.u_box, u_box, first, count, n, false)) return rc;
  }
  // EXECUTE recomputes the scaling, KEEP re-applies the stored one
  {
    const char* mode_env = std::getenv("PQP_E2E");
    if (first == 0 && count == b->B && count > 0 && b->fused_ok && (!mode_env || std::string(mode_env) == "fused")) {
      b->deferred = update_preconditioner ? 1 : 2; // done by the next solve's persistent kernel, QP by QP
      b->deferred_gated = false;
      return 0;
    }
  }
  return launch_setup(b, first, count, update_preconditioner != 0, false);
}

int
pqp_batch_warm_start(pqp_batch* b, int64_t first, int64_t count, const double* x, const double* y, const double* z)
{
  if (int rc = check_range(b, first, count)) return rc;
  if (!x && !y && !z) return 0;
  CUDA_TRY(cudaSetDevice(b->device));
  b->nchunks_pending = 0;
  for (int64_t i = first; i < first + count; ++i) b->hparams[i].s.initial_guess = PQP_WARM_START; // sticky, helpers.hpp:727
  if (int rc = copy_in(b, b->p.x, x, first, count, b->d.n, false)) return rc;
  if (int rc = copy_in(b, b->p.y, y, first, count, b->d.ne, false)) return rc;
  if (int rc = copy_in(b, b->p.z, z, first, count, b->d.nc, false)) return rc;
  return 0;
}

int
pqp_batch_solve_async(pqp_batch* b, void* stream_)
{
  if (!b) return fail(PQP_EINVAL, "null batch");
  CUDA_TRY(cudaSetDevice(b->device));
  cudaStream_t st = stream_ ? (cudaStream_t)stream_ : b->stream;
  // qp_solve prologue (solver.hpp:1125-1377) decided per QP on the host
  for (int64_t i = 0; i < b->B; ++i) {
    PqpQpParams& p = b->hparams[i];
    pqp_info& info = b->hinfo[i];
    QpFlags& f = b->flags[i];
    p.active = (f.is_initialized && (b->selected.empty() || b->selected[(size_t)i])) ? 1 : 0;
    if (!p.active) continue;
    const int ig = p.s.initial_guess;
    if (f.dirty) {
      switch (ig) {
        case PQP_EQUALITY_CONSTRAINED_INITIAL_GUESS:
        case PQP_NO_INITIAL_GUESS:
          cold_start(info, &p.s, b->backend); // results.cleanup(settings)
          p.start_mode = (ig == PQP_NO_INITIAL_GUESS) ? PQP_START_COLD : PQP_START_EQ_GUESS;
          break;
        case PQP_COLD_START_WITH_PREVIOUS_RESULT:
        case PQP_WARM_START:
          cold_start(info, &p.s, b->backend);
          p.start_mode = PQP_START_WARM;
          break;
        default:
          cleanup_statistics(info);
          p.start_mode = PQP_START_WARM_KEEP;
          break;
      }
    } else {
      switch (ig) {
        case PQP_EQUALITY_CONSTRAINED_INITIAL_GUESS: p.start_mode = PQP_START_EQ_GUESS; break;
        case PQP_NO_INITIAL_GUESS: p.start_mode = PQP_START_COLD; break;
        case PQP_COLD_START_WITH_PREVIOUS_RESULT:
        case PQP_WARM_START: p.start_mode = PQP_START_WARM; break;
        default: p.start_mode = PQP_START_WARM_KEEP; break;
      }
    }
    p.rho = info.rho;
    p.mu_eq = info.mu_eq;
    p.mu_in = info.mu_in;
  }
  bool all_active = true, any_active = false;
  for (int64_t i = 0; i < b->B; ++i) {
    all_active = all_active && b->hparams[i].active;
    any_active = any_active || b->hparams[i].active;
  }
  if (!any_active) { // solve before init (pqp.h: PQP_ESTATE), or a selection without an initialised QP
    b->selected.clear();
    return fail(PQP_ESTATE, "solve on a batch without an initialised QP (call init first)");
  }
  const bool partial = !b->selected.empty();
  b->selected.clear(); // the selection is consumed by this solve
  if (b->deferred && (b->prof || b->dbg || partial)) { // (a partial solve must not leave the other QPs without their set-up)
    if (int rc = flush_deferred(b)) return rc;
  }
  if (b->deferred) {
    // fused feed: one persistent launch that equilibrates and solves, consuming QPs as their inputs arrive
    const int code = b->deferred;
    const bool gated = b->deferred_gated;
    b->deferred = 0;
    b->deferred_gated = false;
    b->setup_timed = false; // the set-up is part of the solve kernel: solve_ms covers both
    if (st != b->stream) { // uploads / memsets of init() were enqueued on the batch's own stream
      CUDA_TRY(cudaEventRecord(b->ev_main, b->stream));
      CUDA_TRY(cudaStreamWaitEvent(st, b->ev_main, 0));
    }
    if (int rc = enqueue_solve(b, st, b->lay, b->grid, 0, -1, 0, true, code, gated ? b->d_ready : nullptr)) return rc;
    if (gated) { // later work on this stream must also see the complete upload
      CUDA_TRY(cudaEventRecord(b->ev_feed, b->copy_stream));
      CUDA_TRY(cudaStreamWaitEvent(st, b->ev_feed, 0));
    }
  } else if (!stream_ && b->nchunks_pending > 1 && all_active && !b->prof && !b->dbg && !std::getenv("PQP_NO_PIPELINE")) {
    // pipelined: one launch per uploaded chunk, on the stream that runs the chunk's set-up kernel
    CUDA_TRY(cudaEventRecord(b->ev2, b->cstream[0])); // solve_ms then spans first chunk start .. last chunk end
    for (int k = 0; k < b->nchunks_pending; ++k) {
      const int64_t cnt = b->chunk_count[k];
      const int grid = (int)std::min<int64_t>(b->grid, cnt);
      if (int rc = enqueue_solve(b, b->cstream[k % PQP_CSTREAMS], b->lay, grid, b->chunk_first[k], cnt, k % PQP_CSTREAMS, false)) return rc;
    }
    if (int rc = join_cstreams(b)) return rc;
    CUDA_TRY(cudaEventRecord(b->ev3, b->stream));
    b->launches += b->nchunks_pending - 1;
  } else if (int rc = enqueue_solve(b, st, b->lay, b->grid)) {
    return rc;
  }
  b->nchunks_pending = 0;
  b->solve_timed = true;
  b->launches += 1;
  b->solve_pending = true;
  for (int64_t i = 0; i < b->B; ++i) {
    if (b->hparams[i].active) {
      b->flags[i].dirty = true; // solver.hpp:1835-1836
      b->flags[i].is_initialized = true;
    }
  }
  return 0;
}

int
pqp_batch_select(pqp_batch* b, const int64_t* indices, int64_t count)
{
  if (!b) return fail(PQP_EINVAL, "null batch");
  b->selected.clear();
  if (!indices || count < 0) return 0;
  b->selected.assign((size_t)b->B, 0);
  for (int64_t k = 0; k < count; ++k) {
    if (indices[k] < 0 || indices[k] >= b->B) {
      b->selected.clear();
      return fail(PQP_EINVAL, "wrong argument size: QP index out of bounds");
    }
    b->selected[(size_t)indices[k]] = 1;
  }
  return 0;
}

int
pqp_batch_sync(pqp_batch* b)
{
  if (!b) return fail(PQP_EINVAL, "null batch");
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  CUDA_TRY(cudaDeviceSynchronize());
  if (b->solve_pending && b->d_ready) {
    // fused feed: a CTA that waited 20 s for its QP's inputs raised the abort flag and the rest of the batch was left
    // NOT_RUN - that must not look like a successful solve
    int32_t flag[2] = { 0, 0 };
    CUDA_TRY(cudaMemcpy(flag, b->d_ready, sizeof(flag), cudaMemcpyDeviceToHost));
    if (flag[1] != 0) {
      b->solve_pending = false;
      return fail(PQP_ECUDA, "fused feed aborted: the inputs of a QP did not arrive within 20 s (status NOT_RUN on the affected QPs)");
    }
  }
  if (b->solve_pending) {
    std::vector<double> raw((size_t)b->B * PQP_INFO_DOUBLES);
    CUDA_TRY(cudaMemcpy(raw.data(), b->p.info, sizeof(double) * raw.size(), cudaMemcpyDeviceToHost));
    // QPs whose active set outgrew the shared-memory S^-1 of the compact layout
    // report the internal status 99: re-solve exactly those with the generic
    // kernel (full capacity, inverse blocks in global memory).
    {
      std::vector<int64_t> retry;
      for (int64_t i = 0; i < b->B; ++i) {
        if (b->hparams[i].active && raw[(size_t)i * PQP_INFO_DOUBLES + 10] == 99.0) retry.push_back(i);
      }
      if (!retry.empty()) b->overflow_retries += (int64_t)retry.size();
      b->retry_ms = 0;
      // level 1: the tile kernel with its largest capacity (one CTA per SM); level 2: the general kernel
      for (int level = (b->lay_big.kind == 1 ? 1 : 2); level <= 2 && !retry.empty(); ++level) {
        std::vector<int32_t> saved((size_t)b->B);
        for (int64_t i = 0; i < b->B; ++i) {
          saved[i] = b->hparams[i].active;
          b->hparams[i].active = 0;
        }
        for (int64_t i : retry) b->hparams[i].active = 1;
        // (not `timed`: the events of the main launch must survive; the retry launches are timed on their own pair and added)
        CUDA_TRY(cudaEventRecord(b->ev_r0, b->stream));
        int rc = (level == 1) ? enqueue_solve(b, b->stream, b->lay_big, b->grid_big, 0, -1, 0, false) : enqueue_solve(b, b->stream, b->lay_gen, b->grid_gen, 0, -1, 0, false);
        for (int64_t i = 0; i < b->B; ++i) b->hparams[i].active = saved[i];
        if (rc != 0) return rc;
        CUDA_TRY(cudaEventRecord(b->ev_r1, b->stream));
        b->launches += 1;
        CUDA_TRY(cudaStreamSynchronize(b->stream));
        {
          float ms_r = 0;
          if (cudaEventElapsedTime(&ms_r, b->ev_r0, b->ev_r1) == cudaSuccess) b->retry_ms += ms_r;
        }
        CUDA_TRY(cudaMemcpy(raw.data(), b->p.info, sizeof(double) * raw.size(), cudaMemcpyDeviceToHost));
        std::vector<int64_t> still;
        for (int64_t i : retry) {
          if (raw[(size_t)i * PQP_INFO_DOUBLES + 10] == 99.0) still.push_back(i);
        }
        retry.swap(still);
      }
    }
    float ms_solve = 0, ms_setup = 0;
    if (b->solve_timed) cudaEventElapsedTime(&ms_solve, b->ev2, b->ev3);
    ms_solve += b->retry_ms;
    if (b->setup_timed) cudaEventElapsedTime(&ms_setup, b->ev0, b->ev1);
    int64_t nact = 0;
    for (int64_t i = 0; i < b->B; ++i) nact += b->hparams[i].active ? 1 : 0;
    for (int64_t i = 0; i < b->B; ++i) {
      if (!b->hparams[i].active) continue;
      const double* I = raw.data() + (size_t)i * PQP_INFO_DOUBLES;
      pqp_info& o = b->hinfo[i];
      o.mu_eq = I[0];
      o.mu_eq_inv = I[1];
      o.mu_in = I[2];
      o.mu_in_inv = I[3];
      o.rho = I[4];
      o.nu = I[5];
      o.iter = (int64_t)I[6];
      o.iter_ext = (int64_t)I[7];
      o.mu_updates = (int64_t)I[8];
      o.rho_updates = (int64_t)I[9];
      o.status = (int64_t)I[10];
      o.objValue = I[14];
      o.pri_res = I[15];
      o.dua_res = I[16];
      o.duality_gap = I[17];
      o.iterative_residual = I[18];
      if (b->hparams[i].s.compute_timings && nact > 0) {
        // microseconds; the batch runs as one kernel, so per-QP time = batch time / batch size
        o.solve_time = 1e3 * ms_solve / double(nact);
        o.setup_time = 1e3 * ms_setup / double(nact);
        o.run_time = o.solve_time + o.setup_time;
      }
    }
    b->solve_pending = false;
  }
  return 0;
}

int
pqp_batch_solve(pqp_batch* b)
{
  if (int rc = pqp_batch_solve_async(b, nullptr)) return rc;
  return pqp_batch_sync(b);
}

int
pqp_batch_results(pqp_batch* b, int64_t first, int64_t count, double* x, double* y, double* z, double* se, double* si, pqp_info* info)
{
  if (int rc = check_range(b, first, count)) return rc;
  if (int rc = pqp_batch_sync(b)) return rc;
  const PqpDims& d = b->d;
  auto out = [&](double* dst, const double* src, int64_t per) -> int {
    if (!dst || per == 0 || count == 0) return 0;
    CUDA_TRY(cudaMemcpy(dst, src + first * per, sizeof(double) * (size_t)(count * per), cudaMemcpyDeviceToHost));
    return 0;
  };
  if (int rc = out(x, b->p.x, d.n)) return rc;
  if (int rc = out(y, b->p.y, d.ne)) return rc;
  if (int rc = out(z, b->p.z, d.nc)) return rc;
  if (int rc = out(se, b->p.se, d.ne)) return rc;
  if (int rc = out(si, b->p.si, d.nc)) return rc;
  if (info) {
    for (int64_t i = 0; i < count; ++i) info[i] = b->hinfo[first + i];
  }
  return 0;
}

int
pqp_batch_results_device(pqp_batch* b, double** x, double** y, double** z, double** info20)
{
  if (!b) return fail(PQP_EINVAL, "null batch");
  if (x) *x = b->p.x;
  if (y) *y = b->p.y;
  if (z) *z = b->p.z;
  if (info20) *info20 = b->p.info;
  return 0;
}

int
pqp_batch_scaled(pqp_batch* b, int64_t index, double* H, double* g, double* A, double* b_, double* C, double* u, double* l, double* delta, double* c)
{
  if (b) {
    if (int rc = flush_deferred(b)) return rc;
  }
  if (int rc = check_range(b, index, 1)) return rc;
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  const PqpDims& d = b->d;
  const size_t n = d.n, ne = d.ne, ni = d.ni, nc = d.nc, i = (size_t)index;
  auto out = [&](double* dst, const double* src, size_t per) -> int {
    if (!dst || per == 0) return 0;
    CUDA_TRY(cudaMemcpy(dst, src + i * per, sizeof(double) * per, cudaMemcpyDeviceToHost));
    return 0;
  };
  if (int rc = out(H, b->p.Hs, n * n)) return rc;
  if (int rc = out(g, b->p.gs, n)) return rc;
  if (int rc = out(A, b->p.As, ne * n)) return rc;
  if (int rc = out(b_, b->p.bs, ne)) return rc;
  if (int rc = out(C, b->p.Cs, ni * n)) return rc;
  if (int rc = out(u, b->p.us, nc)) return rc;
  if (int rc = out(l, b->p.ls, nc)) return rc;
  if (int rc = out(delta, b->p.delta, n + ne + nc)) return rc;
  if (int rc = out(c, b->p.c, 1)) return rc;
  return 0;
}

static int
do_backward(pqp_batch* b, int64_t first, int64_t count, const double* loss_derivative, double eps, double rho_new, double mu_new, double* dL_dH, double* dL_dg, double* dL_dA, double* dL_db, double* dL_dC, double* dL_du, double* dL_dl, bool dev_ptrs)
{
  if (int rc = check_range(b, first, count)) return rc;
  if (!loss_derivative) return fail(PQP_EINVAL, "wrong argument size: loss_derivative is required");
  const PqpDims& d = b->d;
  if (d.box) return fail(PQP_EINVAL, "compute_backward: QPs with box constraints are not handled by the backward pass");
  CUDA_TRY(cudaSetDevice(b->device));
  if (int rc = flush_deferred(b)) return rc;
  if (b->solve_pending) {
    if (int rc = pqp_batch_sync(b)) return rc;
  }
  for (int64_t i = first; i < first + count; ++i) {
    if (!b->flags[i].is_initialized || b->hinfo[i].status == PQP_NOT_RUN) return fail(PQP_ESTATE, "compute_backward on a QP that has not been solved");
    if (b->hinfo[i].status == PQP_DUAL_INFEASIBLE) // compute_ECJ.hpp:37-46
      return fail(PQP_EINVAL, "the QP problem is not feasible, so computing the derivatives is not valid in this setting. Try enabling infeasible solving if the problem is only primally infeasible.");
  }
  if (count == 0) return 0;
  const size_t B = (size_t)b->B, n = d.n, ne = d.ne, ni = d.ni, nt = n + ne + ni;
  if (!b->bw_loss) {
    int rc = 0;
    rc |= dev_alloc(b, &b->bw_loss, B * nt);
    rc |= dev_alloc(b, &b->bw_dH, B * n * n);
    rc |= dev_alloc(b, &b->bw_dg, B * n);
    rc |= dev_alloc(b, &b->bw_dA, B * ne * n);
    rc |= dev_alloc(b, &b->bw_db, B * ne);
    rc |= dev_alloc(b, &b->bw_dC, B * ni * n);
    rc |= dev_alloc(b, &b->bw_du, B * ni);
    rc |= dev_alloc(b, &b->bw_dl, B * ni);
    if (rc != 0) {
      b->bw_loss = nullptr;
      return PQP_ECUDA;
    }
  }
  cudaStream_t st = b->stream;
  CUDA_TRY(cudaMemcpyAsync(b->bw_loss + (size_t)first * nt, loss_derivative, sizeof(double) * (size_t)count * nt, dev_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  std::vector<int32_t> saved((size_t)b->B);
  for (int64_t i = 0; i < b->B; ++i) {
    saved[(size_t)i] = b->hparams[i].active;
    b->hparams[i].active = (i >= first && i < first + count) ? 1 : 0;
  }
  cudaError_t e1 = cudaMemcpyAsync(b->p.params, b->hparams.data(), sizeof(PqpQpParams) * (size_t)b->B, cudaMemcpyHostToDevice, st);
  for (int64_t i = 0; i < b->B; ++i) b->hparams[i].active = saved[(size_t)i];
  if (e1 != cudaSuccess) return fail(PQP_ECUDA, std::string("cudaMemcpyAsync(params): ") + cudaGetErrorString(e1));
  CUDA_TRY(cudaMemsetAsync(b->counter, 0, sizeof(int32_t), st));
  PqpSolveArgs a{};
  a.d = b->d;
  a.p = b->p;
  a.lay = b->lay_gen; // general kernel body, full-capacity layout
  a.batch = (int32_t)b->B;
  a.first = 0;
  a.counter = b->counter;
  a.ws = b->ws;
  PqpBackwardArgs k{};
  k.loss_derivative = b->bw_loss;
  k.eps = eps;
  k.rho_new = rho_new;
  k.mu_new = mu_new;
  k.dL_dH = b->bw_dH;
  k.dL_dg = b->bw_dg;
  k.dL_dA = b->bw_dA;
  k.dL_db = b->bw_db;
  k.dL_dC = b->bw_dC;
  k.dL_du = b->bw_du;
  k.dL_dl = b->bw_dl;
  int rc = pqp_launch_backward(&a, &k, b->grid_gen, st);
  if (rc != 0) return fail(PQP_ECUDA, std::string("backward kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
  b->launches += 1;
  auto out = [&](double* dst, const double* src, size_t per) -> int {
    if (!dst || per == 0) return 0;
    CUDA_TRY(cudaMemcpyAsync(dst, src + (size_t)first * per, sizeof(double) * (size_t)count * per, dev_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    return 0;
  };
  if (int r = out(dL_dH, b->bw_dH, n * n)) return r;
  if (int r = out(dL_dg, b->bw_dg, n)) return r;
  if (int r = out(dL_dA, b->bw_dA, ne * n)) return r;
  if (int r = out(dL_db, b->bw_db, ne)) return r;
  if (int r = out(dL_dC, b->bw_dC, ni * n)) return r;
  if (int r = out(dL_du, b->bw_du, ni)) return r;
  if (int r = out(dL_dl, b->bw_dl, ni)) return r;
  CUDA_TRY(cudaStreamSynchronize(st));
  for (int64_t i = first; i < first + count; ++i) { // compute_ECJ.hpp:66-68
    b->hinfo[i].rho = rho_new;
    b->hinfo[i].mu_eq = mu_new;
    b->hinfo[i].mu_in = mu_new;
  }
  return 0;
}

int
pqp_batch_backward(pqp_batch* b, int64_t first, int64_t count, const double* loss_derivative, double eps, double rho_new, double mu_new, double* dL_dH, double* dL_dg, double* dL_dA, double* dL_db, double* dL_dC, double* dL_du, double* dL_dl)
{
  return do_backward(b, first, count, loss_derivative, eps, rho_new, mu_new, dL_dH, dL_dg, dL_dA, dL_db, dL_dC, dL_du, dL_dl, false);
}
int
pqp_batch_backward_device(pqp_batch* b, int64_t first, int64_t count, const double* loss_derivative, double eps, double rho_new, double mu_new, double* dL_dH, double* dL_dg, double* dL_dA, double* dL_db, double* dL_dC, double* dL_du, double* dL_dl)
{
  return do_backward(b, first, count, loss_derivative, eps, rho_new, mu_new, dL_dH, dL_dg, dL_dA, dL_db, dL_dC, dL_du, dL_dl, true);
}

// x, y, z of the QPs [first, first + count) into caller-owned DEVICE buffers (torch CUDA tensors)
int
pqp_batch_results_copy_device(pqp_batch* b, int64_t first, int64_t count, double* x, double* y, double* z)
{
  if (int rc = check_range(b, first, count)) return rc;
  CUDA_TRY(cudaSetDevice(b->device));
  if (b->solve_pending) {
    if (int rc = pqp_batch_sync(b)) return rc;
  }
  const PqpDims& d = b->d;
  if (x && count) CUDA_TRY(cudaMemcpyAsync(x, b->p.x + first * d.n, sizeof(double) * (size_t)(count * d.n), cudaMemcpyDeviceToDevice, b->stream));
  if (y && count && d.ne) CUDA_TRY(cudaMemcpyAsync(y, b->p.y + first * d.ne, sizeof(double) * (size_t)(count * d.ne), cudaMemcpyDeviceToDevice, b->stream));
  if (z && count && d.nc) CUDA_TRY(cudaMemcpyAsync(z, b->p.z + first * d.nc, sizeof(double) * (size_t)(count * d.nc), cudaMemcpyDeviceToDevice, b->stream));
  CUDA_TRY(cudaStreamSynchronize(b->stream));
  return 0;
}

int
pqp_batch_cleanup(pqp_batch* b, int64_t first, int64_t count)
{
  if (int rc = check_range(b, first, count)) return rc;
  CUDA_TRY(cudaSetDevice(b->device));
  b->nchunks_pending = 0;
  if (int rc = zero_results(b, first, count)) return rc;
  for (int64_t i = first; i < first + count; ++i) {
    cold_start(b->hinfo[i], &b->hparams[i].s, b->backend);
    // workspace.hpp:330-377 clears every flag; the reference's next solve() re-runs the set-up from the model it
    // still holds. The model data of this batch stay resident as well, so the QP stays solvable (is_initialized is
    // this library's "has a model" flag: without it the next solve would silently skip the QP).
    const bool had_model = b->flags[i].is_initialized;
    b->flags[i] = QpFlags();
    b->flags[i].is_initialized = had_model;
  }
  return 0;
}

// test / diagnostic hook (not in pqp.h): resident CTAs per SM of the primary layout's solve kernel, plain or fused
int
pqp_batch_occupancy(pqp_batch* b, int fused)
{
  if (!b) return -1;
  cudaSetDevice(b->device);
  PqpSolveArgs a{};
  a.d = b->d;
  a.lay = b->lay;
  return pqp_solve_occupancy(&a, fused);
}

int
pqp_batch_timings(const pqp_batch* b, double* setup_ms, double* solve_ms, int64_t* launches)
{
  if (!b) return fail(PQP_EINVAL, "null batch");
  float ms = 0;
  if (setup_ms) {
    *setup_ms = 0;
    if (b->setup_timed && cudaEventElapsedTime(&ms, b->ev0, b->ev1) == cudaSuccess) *setup_ms = ms;
  }
  if (solve_ms) {
    *solve_ms = 0;
    if (b->solve_timed && cudaEventElapsedTime(&ms, b->ev2, b->ev3) == cudaSuccess) *solve_ms = ms + b->retry_ms; // main launch + retry launches
  }
  if (launches) *launches = b->launches;
  return 0;
}

// per-phase cycle counters accumulated since creation (PQP_PROFILE=1 in the environment)
int
pqp_batch_profile(pqp_batch* b, long long* out12, int reset)
{
  if (!b || !b->prof) return 0;
  cudaDeviceSynchronize();
  cudaMemcpy(out12, b->prof, sizeof(long long) * 12, cudaMemcpyDeviceToHost);
  if (reset) cudaMemset(b->prof, 0, sizeof(long long) * 16);
  return 12;
}

// debug trace access (PQP_DEBUG_TRACE=<qp index> in the environment)
int
pqp_batch_debug_trace(pqp_batch* b, double* out, int64_t cap)
{
  if (!b || !b->dbg) return 0;
  cudaMemcpy(out, b->dbg, sizeof(double) * (size_t)std::min<int64_t>(cap, b->dbg_cap), cudaMemcpyDeviceToHost);
  return (int)std::min<int64_t>(cap, b->dbg_cap);
}

int
pqp_batch_launch_config(const pqp_batch* b, int* grid, int* smem_bytes, int* in_smem_mask, int64_t* ws_doubles)
{
  if (!b) return fail(PQP_EINVAL, "null batch");
  if (grid) *grid = b->grid;
  if (smem_bytes) *smem_bytes = (int)(sizeof(double) * (size_t)b->lay.smem_doubles + (size_t)b->lay.smem_int_bytes);
  if (in_smem_mask) {
    int m = 0;
    for (int i = 0; i < PA_COUNT; ++i) m |= (b->lay.in_smem[i] ? 1 : 0) << i;
    *in_smem_mask = m;
  }
  if (ws_doubles) *ws_doubles = b->lay.ws_doubles + ((int64_t)b->lay.si_cap << 32) + ((int64_t)b->overflow_retries << 48);
  return 0;
}

int
pqp_random_qp(int kind, uint64_t seed, int64_t dim, int64_t n_eq, int64_t n_in, double sparsity, double strong_convexity, double* H, double* g, double* A, double* b_, double* C, double* u, double* l, double* u_box, double* l_box)
{
  pqp::randqp::Lehmer rng;
  rng.set_seed(seed);
  const int n = (int)dim, ne = (int)n_eq, ni = (int)n_in;
  switch (kind) {
    case 0: pqp::randqp::dense_strongly_convex_qp(rng, n, ne, ni, sparsity, strong_convexity, H, g, A, b_, C, u, l); return 0;
    case 1: pqp::randqp::dense_not_strongly_convex_qp(rng, n, ne, ni, sparsity, H, g, A, b_, C, u, l); return 0;
    case 2: pqp::randqp::dense_degenerate_qp(rng, n, ne, ni, sparsity, strong_convexity, H, g, A, b_, C, u, l); return 0;
    case 3: pqp::randqp::dense_box_constrained_qp(rng, n, ne, ni, sparsity, strong_convexity, H, g, A, b_, C, u, l); return 0;
    case 4: pqp::randqp::dense_box_benchmark_qp(rng, n, ne, ni, sparsity, strong_convexity, 0, H, g, A, b_, C, u, l, u_box, l_box); return 0;
    case 5: pqp::randqp::dense_box_benchmark_qp(rng, n, ne, ni, sparsity, strong_convexity, 1, H, g, A, b_, C, u, l, u_box, l_box); return 0;
  }
  return fail(PQP_EINVAL, "unknown generator kind");
}

} // extern "C"
