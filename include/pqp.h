/* pqp.h — C-ABI of the B200-native batched dense ProxQP path.
 *
 * Drop-in boundary for ONE path of Simple-Robotics/proxsuite:
 *   proxsuite::proxqp::dense::QP<T> / BatchQP<T> + dense::solve_in_parallel
 * (reference files, relative to /root/reference/include/proxsuite/proxqp):
 *   parallel/qp_solve.hpp:17-60     solve_in_parallel(std::vector<QP>&) / (BatchQP&)
 *   dense/wrapper.hpp:115-963       QP<T>: ctor / init / update / solve / cleanup
 *   dense/wrapper.hpp:1253-1311     BatchQP<T>
 *   settings.hpp:88-316             Settings<T>
 *   results.hpp:28-203              Info<T> / Results<T>
 *   status.hpp:17-43                enums
 * The reference has no C ABI (Eigen types cannot cross one); every entry point
 * below names the reference call it replaces. Plain pointers and sizes only,
 * no torch / Eigen types. All matrices are ROW-MAJOR (dense/fwd.hpp:16-33),
 * batch-major: QP i's H starts at H + i*n*n. T = double (fp64) throughout.
 *
 * Error convention: 0 = success, negative = error (PQP_EINVAL for what the
 * reference throws as std::invalid_argument, macros.hpp:18-35); the message
 * is available from pqp_last_error(). Nothing throws across the ABI.
 * There is NO CPU fallback: every solve runs the CUDA kernels and fails with
 * PQP_ECUDA when no device is usable.
 */
#ifndef PQP_H
#define PQP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PQP_OK 0
#define PQP_EINVAL (-1)  /* std::invalid_argument in the reference */
#define PQP_ECUDA (-2)   /* CUDA runtime / no device */
#define PQP_ESTATE (-3)  /* call order violation (e.g. solve before init) */

/* status.hpp:17-26 QPSolverOutput */
enum pqp_status {
  PQP_SOLVED = 0,
  PQP_MAX_ITER_REACHED = 1,
  PQP_PRIMAL_INFEASIBLE = 2,
  PQP_SOLVED_CLOSEST_PRIMAL_FEASIBLE = 3,
  PQP_DUAL_INFEASIBLE = 4,
  PQP_NOT_RUN = 5
};
/* status.hpp:28-35 InitialGuessStatus */
enum pqp_initial_guess {
  PQP_NO_INITIAL_GUESS = 0,
  PQP_EQUALITY_CONSTRAINED_INITIAL_GUESS = 1,
  PQP_WARM_START_WITH_PREVIOUS_RESULT = 2,
  PQP_WARM_START = 3,
  PQP_COLD_START_WITH_PREVIOUS_RESULT = 4
};
/* settings.hpp:26-45 */
enum pqp_dense_backend { PQP_BACKEND_AUTOMATIC = 0, PQP_BACKEND_PRIMAL_DUAL_LDLT = 1, PQP_BACKEND_PRIMAL_LDLT = 2 };
enum pqp_merit_function { PQP_MERIT_GPDAL = 0, PQP_MERIT_PDAL = 1 };
enum pqp_hessian_type { PQP_HESSIAN_ZERO = 0, PQP_HESSIAN_DENSE = 1, PQP_HESSIAN_DIAGONAL = 2 };

/* Field-for-field mirror of Settings<double> (settings.hpp:88-210, defaults
 * :213-315). bool -> int32. `sparse_backend` is not on this path. */
typedef struct pqp_settings {
  double default_rho;
  double default_mu_eq;
  double default_mu_in;
  double alpha_bcl;
  double beta_bcl;
  double refactor_dual_feasibility_threshold;
  double refactor_rho_threshold;
  double mu_min_eq;
  double mu_min_in;
  double mu_max_eq_inv;
  double mu_max_in_inv;
  double mu_update_factor;
  double mu_update_inv_factor;
  double cold_reset_mu_eq;
  double cold_reset_mu_in;
  double cold_reset_mu_eq_inv;
  double cold_reset_mu_in_inv;
  double eps_abs;
  double eps_rel;
  double eps_refact;
  double eps_duality_gap_abs;
  double eps_duality_gap_rel;
  double preconditioner_accuracy;
  double eps_primal_inf;
  double eps_dual_inf;
  double alpha_gpdal;
  double default_H_eigenvalue_estimate;
  int64_t max_iter;
  int64_t max_iter_in;
  int64_t safe_guard;
  int64_t nb_iterative_refinement;
  int64_t preconditioner_max_iter;
  int64_t frequence_infeasibility_check;
  int32_t verbose;
  int32_t initial_guess; /* enum pqp_initial_guess */
  int32_t update_preconditioner;
  int32_t compute_preconditioner;
  int32_t compute_timings;
  int32_t check_duality_gap;
  int32_t bcl_update;
  int32_t merit_function_type; /* enum pqp_merit_function */
  int32_t primal_infeasibility_solving;
  int32_t reserved_;
} pqp_settings;

/* Mirror of Info<double> (results.hpp:28-58). */
typedef struct pqp_info {
  double mu_eq;
  double mu_eq_inv;
  double mu_in;
  double mu_in_inv;
  double rho;
  double nu;
  int64_t iter;
  int64_t iter_ext;
  int64_t mu_updates;
  int64_t rho_updates;
  int64_t status; /* enum pqp_status */
  double setup_time; /* microseconds, batch time / batch size */
  double solve_time;
  double run_time;
  double objValue;
  double pri_res;
  double dua_res;
  double duality_gap;
  double iterative_residual;
  double minimal_H_eigenvalue_estimate;
} pqp_info;

typedef struct pqp_batch pqp_batch; /* opaque: a device-resident batch of same-shaped QPs */

/* Settings<double>::Settings(dense_backend) defaults, settings.hpp:213-315. */
void pqp_settings_default(pqp_settings* s, int dense_backend);

/* dense_backend_choice<T>, wrapper.hpp:82-113. */
int pqp_dense_backend_choice(int dense_backend, int64_t dim, int64_t n_eq, int64_t n_in, int box_constraints);

/* BatchQP<T>(batch_size) + init_qp_in_place(dim, n_eq, n_in) for every slot
 * (wrapper.hpp:1263-1283), extended with the QP<T> ctor arguments
 * box_constraints / HessianType / DenseBackend (wrapper.hpp:140-333) that
 * init_qp_in_place cannot express. `device` is the CUDA ordinal (-1: current).
 * dim == 0 -> PQP_EINVAL (model.hpp:65-68). Returns NULL on error. */
pqp_batch* pqp_batch_create(int64_t batch, int64_t dim, int64_t n_eq, int64_t n_in, int box_constraints, int hessian_type, int dense_backend, int device);
void pqp_batch_destroy(pqp_batch* b);

int64_t pqp_batch_size(const pqp_batch* b); /* BatchQP::size(), wrapper.hpp:1310 */
int pqp_batch_dims(const pqp_batch* b, int64_t* dim, int64_t* n_eq, int64_t* n_in, int* box_constraints, int* hessian_type, int* dense_backend);

/* qp.settings access (wrapper.hpp:126). index = -1 addresses every QP. */
int pqp_batch_settings_get(const pqp_batch* b, int64_t index, pqp_settings* out);
int pqp_batch_settings_set(pqp_batch* b, int64_t index, const pqp_settings* in);

/* QP<T>::init(H,g,A,b,C,l,u[,l_box,u_box], compute_preconditioner, rho, mu_eq,
 * mu_in, manual_minimal_H_eigenvalue) for QPs [first, first+count)
 * (wrapper.hpp:354-498, 520-703). Host pointers, batch-major over `count`
 * QPs; NULL = absent (nullopt). rho/mu_eq/mu_in/manual_eig: NULL = nullopt,
 * else ONE value applied to every addressed QP. Copies the data (the caller's
 * buffers are not retained, helpers.hpp:573-612), clamps the bounds and runs
 * the Ruiz equilibration on the device.
 * Asynchrony: pageable host buffers are consumed before the call returns;
 * PINNED (page-locked) buffers are read by the copy engine after it returns
 * and must stay unchanged until the next pqp_batch_solve / pqp_batch_sync /
 * pqp_batch_scaled returns. An init / update of the WHOLE batch only uploads:
 * the equilibration is done by the solve kernel itself (the CTA that takes a
 * QP sets it up, then solves it, and starts while later QPs are still being
 * uploaded); any other call in between runs it as a separate launch first. */
int pqp_batch_init(pqp_batch* b, int64_t first, int64_t count, const double* H, const double* g, const double* A, const double* b_, const double* C, const double* l, const double* u, const double* l_box, const double* u_box,
                   int compute_preconditioner, const double* rho, const double* mu_eq, const double* mu_in, const double* manual_minimal_H_eigenvalue);

/* Same with DEVICE pointers (no host round trip; torch CUDA tensors). */
int pqp_batch_init_device(pqp_batch* b, int64_t first, int64_t count, const double* H, const double* g, const double* A, const double* b_, const double* C, const double* l, const double* u, const double* l_box, const double* u_box,
                          int compute_preconditioner, const double* rho, const double* mu_eq, const double* mu_in, const double* manual_minimal_H_eigenvalue);

/* QP<T>::update(...) (wrapper.hpp:723-918): NULL = unchanged. */
int pqp_batch_update(pqp_batch* b, int64_t first, int64_t count, const double* H, const double* g, const double* A, const double* b_, const double* C, const double* l, const double* u, const double* l_box, const double* u_box,
                     int update_preconditioner, const double* rho, const double* mu_eq, const double* mu_in, const double* manual_minimal_H_eigenvalue);

/* QP<T>::solve(x, y, z) warm start part (helpers.hpp:715-763): stores the
 * guess in results and sticky-sets initial_guess = WARM_START. NULL = absent. */
int pqp_batch_warm_start(pqp_batch* b, int64_t first, int64_t count, const double* x, const double* y, const double* z);

/* solve_in_parallel(BatchQP&) (parallel/qp_solve.hpp:40-60): solves every
 * initialised QP of the batch with the persistent CUDA kernel. Blocking. */
int pqp_batch_solve(pqp_batch* b);
/* Enqueue only (no host synchronisation); results become valid after
 * pqp_batch_sync(). `stream` is a cudaStream_t passed as void* (NULL: the
 * batch's own stream). Used by bench.py to time with CUDA events. */
int pqp_batch_solve_async(pqp_batch* b, void* stream);
int pqp_batch_sync(pqp_batch* b);
/* Restricts the NEXT pqp_batch_solve / pqp_batch_solve_async to the listed QPs (QP<T>::solve() of one member of a
 * BatchQP touches that QP only, wrapper.hpp:922-954; solve_in_parallel(std::vector<QP>&) the listed ones,
 * parallel/qp_solve.hpp:17-38). indices == NULL or count < 0: every QP again. The selection is consumed by that
 * solve; results, iteration counts and warm-start state of the other QPs are left untouched. */
int pqp_batch_select(pqp_batch* b, const int64_t* indices, int64_t count);

/* qp.results (results.hpp:67-88): x[dim], y[n_eq], z[n_in (+dim)], se, si, info
 * for QPs [first, first+count); any pointer may be NULL. */
int pqp_batch_results(pqp_batch* b, int64_t first, int64_t count, double* x, double* y, double* z, double* se, double* si, pqp_info* info);
int pqp_batch_results_device(pqp_batch* b, double** x, double** y, double** z, double** info20);
/* x, y, z of the QPs [first, first+count) copied into caller-owned DEVICE buffers (any may be NULL). */
int pqp_batch_results_copy_device(pqp_batch* b, int64_t first, int64_t count, double* x, double* y, double* z);

/* qp.work scaled model + qp.ruiz (workspace.hpp:35-44, ruiz.hpp:319-320) of one
 * QP, for the equilibration identity test. Any pointer may be NULL. */
int pqp_batch_scaled(pqp_batch* b, int64_t index, double* H, double* g, double* A, double* b_, double* C, double* u, double* l, double* delta, double* c);

/* dense::compute_backward (dense/compute_ECJ.hpp:29-190) for the solved QPs
 * [first, first+count) == solve_backward_in_parallel (parallel/qp_solve.hpp:84-138):
 * one more solve of the KKT system regularised with rho_new / mu_new, with the
 * active set at the solution, then the jacobian-vector products of
 * BackwardData (backward_data.hpp:27-50). Host pointers, batch-major:
 * loss_derivative[count][dim + n_eq + n_in] = (dL/dx, dL/dy, dL/dz); outputs
 * dL_dH[count][dim*dim], dL_dg[count][dim], dL_dA[count][n_eq*dim], dL_db,
 * dL_dC[count][n_in*dim], dL_du, dL_dl (any may be NULL). Leaves
 * info.rho = rho_new, info.mu_eq = info.mu_in = mu_new like the reference.
 * PQP_EINVAL: a QP is dual infeasible (std::invalid_argument in the reference)
 * or the batch has box constraints (the QP layer has none); PQP_ESTATE: a QP of
 * the range has not been solved. */
int pqp_batch_backward(pqp_batch* b, int64_t first, int64_t count, const double* loss_derivative, double eps, double rho_new, double mu_new, double* dL_dH, double* dL_dg, double* dL_dA, double* dL_db, double* dL_dC, double* dL_du,
                       double* dL_dl);
/* Same with DEVICE pointers for the loss derivatives and the outputs (torch CUDA tensors, no host round trip). */
int pqp_batch_backward_device(pqp_batch* b, int64_t first, int64_t count, const double* loss_derivative, double eps, double rho_new, double mu_new, double* dL_dH, double* dL_dg, double* dL_dA, double* dL_db, double* dL_dC, double* dL_du,
                              double* dL_dl);

/* QP<T>::cleanup() (wrapper.hpp:958-962). */
int pqp_batch_cleanup(pqp_batch* b, int64_t first, int64_t count);

/* Device time (CUDA events, milliseconds) of the last init/update set-up
 * kernel and of the last solve kernel; number of kernels this library launched
 * since the batch was created. */
int pqp_batch_timings(const pqp_batch* b, double* setup_ms, double* solve_ms, int64_t* kernel_launches);

/* Deterministic synthetic inputs of the reference's tests and benchmarks
 * (utils/random_qp_problems.hpp:104-147, 308-368, 463-628;
 * benchmark/timings-box-constraints.cpp:30-51, timings-diagonal-hessian.cpp:52-56).
 * kind: 0 strongly convex, 1 not strongly convex, 2 degenerate (C,u,l hold
 * 2*n_in rows), 3 box-constrained (C = I), 4 box benchmark, 5 diagonal-Hessian
 * benchmark. Row-major outputs; u_box/l_box only for kinds 4, 5. */
int pqp_random_qp(int kind, uint64_t seed, int64_t dim, int64_t n_eq, int64_t n_in, double sparsity_factor, double strong_convexity_factor, double* H, double* g, double* A, double* b_, double* C, double* u, double* l, double* u_box,
                  double* l_box);

/* ---- One batch sharded over several GPUs of one node -------------------------------------------------------------
 * solve_in_parallel has no cross-QP state (parallel/qp_solve.hpp:55-59): a batch shards into contiguous slices
 * [k B / G, (k + 1) B / G), one pqp_batch per listed device, no cross-device dependency inside the iteration; the
 * only data movement is the scatter of the inputs (each device receives its own slice from the caller's host
 * buffers) and the gather of the solutions into the caller's buffers. `devices` lists CUDA ordinals (an ordinal may
 * be listed more than once). All calls address the WHOLE batch with host pointers laid out like pqp_batch_init's. */
typedef struct pqp_sharded pqp_sharded;
pqp_sharded* pqp_sharded_create(int64_t batch, int64_t dim, int64_t n_eq, int64_t n_in, int box_constraints, int hessian_type, int dense_backend, const int* devices, int n_devices);
void pqp_sharded_destroy(pqp_sharded* s);
int pqp_sharded_count(const pqp_sharded* s);
/* shard k: its per-device batch (for every pqp_batch_* call, e.g. the device-pointer entry points) and its slice */
pqp_batch* pqp_sharded_shard(pqp_sharded* s, int k, int64_t* first, int64_t* count);
int pqp_sharded_settings_set(pqp_sharded* s, const pqp_settings* in); /* every QP */
int pqp_sharded_init(pqp_sharded* s, const double* H, const double* g, const double* A, const double* b_, const double* C, const double* l, const double* u, const double* l_box, const double* u_box, int compute_preconditioner,
                     const double* rho, const double* mu_eq, const double* mu_in, const double* manual_minimal_H_eigenvalue);
int pqp_sharded_update(pqp_sharded* s, const double* H, const double* g, const double* A, const double* b_, const double* C, const double* l, const double* u, const double* l_box, const double* u_box, int update_preconditioner,
                       const double* rho, const double* mu_eq, const double* mu_in, const double* manual_minimal_H_eigenvalue);
/* solve_in_parallel over all devices: every shard is enqueued first, then all are synchronised. */
int pqp_sharded_solve(pqp_sharded* s);
int pqp_sharded_results(pqp_sharded* s, double* x, double* y, double* z, double* se, double* si, pqp_info* info);

const char* pqp_last_error(void);
const char* pqp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PQP_H */
