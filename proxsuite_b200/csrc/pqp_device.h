// Shared host/device declarations of the B200 batched dense ProxQP kernels.
//
// Data layout in HBM (all fp64, batch-major, every matrix row-major — the
// reference's layout, dense/fwd.hpp:16-33):
//   model   : H[B][n*n] g[B][n] A[B][ne*n] b[B][ne] C[B][ni*n] l,u[B][ni] l_box,u_box[B][n]
//             (copies of the caller's data, reference helpers.hpp:573-612)
//   scaled  : Hs, gs, As, bs, Cs, us, ls [B][ncons] (constraint order: C rows then box), is[B][n]
//             delta[B][n+ne+ncons], c[B]          (workspace.hpp:35-44, ruiz.hpp:319-320)
//   results : x[B][n] y[B][ne] z[B][ncons] se[B][ne] si[B][ncons] info[B][20]
//   params  : one PqpQpParams per QP (settings + proximal parameters + start mode)
// Per-CTA (not per-QP) scratch lives in a global workspace that stays
// L2-resident: the Gram matrix G of constraint rows, temporaries, and any of
// the factor arrays that do not fit in shared memory.
#pragma once
#include "../../include/pqp.h"
#include <stdint.h>

#ifndef PQP_NT
#define PQP_NT 256          // threads per CTA: two warp-groups own one QP
#endif
#ifndef PQP_MIN_CTAS
#define PQP_MIN_CTAS 2       // resident CTAs per SM the register budget is sized for (compact layout)
#endif
#define PQP_NW (PQP_NT / 32)
#define PQP_INFO_DOUBLES 20

// start modes derived on the host from settings.initial_guess and the
// init/update/solve state machine (solver.hpp:1125-1377)
enum PqpStartMode {
  PQP_START_COLD = 0,        // x = y = z = 0 (NO_INITIAL_GUESS)
  PQP_START_EQ_GUESS = 1,    // + equality constrained initial guess (helpers.hpp:201-228)
  PQP_START_WARM = 2,        // x, y, z from results (unscaled), proximal parameters reset
  PQP_START_WARM_KEEP = 3    // WARM_START_WITH_PREVIOUS_RESULT: keep rho / mu as they are
};

struct PqpQpParams
{
  pqp_settings s;
  double rho, mu_eq, mu_in; // proximal parameters to start from (results.info)
  int32_t start_mode;
  int32_t active;           // 0: skip this QP (not initialised)
};

struct PqpDims
{
  int n, ne, ni, nc; // nc = ni + (box ? n : 0)
  int box, hess;
  int cap;           // ne + nc : capacity of the dual block
};

struct PqpBatchPtrs
{
  // model (unscaled)
  double *H, *g, *A, *b, *C, *l, *u, *l_box, *u_box;
  // scaled
  double *Hs, *gs, *As, *bs, *Cs, *us, *ls, *is;
  double *delta, *c;
  // results
  double *x, *y, *z, *se, *si, *info;
  PqpQpParams* params;
};

// Arrays the solve kernel places either in shared memory or in the per-CTA
// global workspace (decided on the host, see pqp_layout.cpp).
enum PqpArr {
  PA_M1 = 0,   // P^-1 = (Hs + rho I)^-1, packed lower with diagonal : n(n+1)/2
  PA_AS,       // scaled equality matrix ne x n
  PA_MS,       // S^-1 (inverse of the dual Schur complement), packed lower with diagonal : cap(cap+1)/2
  PA_G,        // packed lower (with diagonal) Gram matrix B P^-1 B^T by row id : cap(cap+1)/2
  PA_Y,        // n x max(ne, 1) temporary (P^-1 A^T)
  PA_VEC,      // all vectors, one arena (sub-offsets below)
  PA_COUNT
};

// vector sub-arena (offsets in doubles from the arena base)
enum PqpVec {
  V_X = 0, V_Y, V_Z, V_XP, V_YP, V_ZP,
  V_DX, V_DS, V_DZ,          // dw: x part, dual slots part, dz in constraint order
  V_RX, V_RS,                // rhs
  V_EX, V_ES,                // err
  V_DUAL, V_SE, V_RUP, V_SI,
  V_HDX, V_ADX, V_ATDY, V_CDX, V_CTDZ, V_Q,
  V_GS, V_BS, V_US, V_LS, V_IS, V_DELTA,
  V_B, V_U, V_L,             // unscaled b, u, l (u, l in constraint order incl. box)
  V_D1INV, V_DSV, V_DSINV,
  V_T1, V_T2, V_T3,          // n-sized temporaries
  V_S1, V_S2, V_S3, V_S4,    // cap-sized temporaries
  V_ALPHAS, V_GRADS,         // 2*nc + 2
  V_SCRATCH,                 // partial-sum scratch
  V_RED,                     // reduction scratch (64)
  V_KT,                      // ne + ni products of one Bt pass (tile layout)
  V_KT2,                     // second dense coefficient vector (tile layout)
  V_COUNT
};

struct PqpLayout
{
  int64_t off[PA_COUNT];      // offset in doubles inside smem or the CTA workspace
  int32_t in_smem[PA_COUNT];
  int32_t voff[V_COUNT];      // offsets inside the vector arena
  int32_t vec_doubles;
  int32_t scratch_doubles;
  int32_t smem_doubles;       // total doubles of dynamic shared memory
  int32_t smem_int_bytes;     // bytes of int scratch that follow the doubles
  int64_t ws_doubles;         // per-CTA global workspace in doubles
  int32_t si_cap;             // dual-block capacity of the S^-1 storage (<= dims.cap)
  int32_t ctas_per_sm;        // resident CTAs per SM this layout is sized for
  int32_t kind;               // 0: packed layouts (pqp_solver_body.inl), 1: tile layout (pqp_fast_body.inl)
};

struct PqpSolveArgs
{
  PqpDims d;
  PqpBatchPtrs p;
  PqpLayout lay;
  int32_t batch;     // number of QPs this launch owns
  int32_t first;     // first QP of this launch
  int32_t* counter;  // dynamic work queue (replaces OpenMP schedule(dynamic), qp_solve.hpp:55)
  double* ws;        // per-CTA workspace base
  double* dbg;       // optional debug trace buffer (NULL = off)
  int32_t dbg_qp;
  int32_t dbg_cap;
  long long* prof;   // optional per-phase cycle counters (12 entries), NULL = off
  int32_t prefetch;   // big variant: software L2 prefetch distance of the streaming passes, in warp iterations (0 = off)
  int32_t force_kkt;  // test hook (PQP_FORCE_KKT=1): the big variant solves every QP through the whole-KKT inverse fallback
  unsigned long long watchdog_ns; // 0 = off; per-QP time budget after which the QP is abandoned (status MAX_ITER_REACHED)
  // Fused feed (end-to-end path: init() from host buffers directly followed by solve()):
  int32_t* ready;      // NULL: every input is resident. Else ready[0] = number of QPs of this launch whose inputs have
                       // arrived (written by a 4-byte copy behind each uploaded chunk), ready[1] = abort flag
  int32_t fused_setup; // 0: the scaled data are ready; bit 0: run Ruiz (EXECUTE), bit 1: re-apply the stored scaling,
                       // bit 2: reset delta / c first (IDENTITY) -- the CTA that pops a QP equilibrates it, then solves it
  int32_t feed_margin; // a QP is consumed once the inputs of the `feed_margin` QPs after it have arrived too: no 128-byte
                       // line of its arrays can then still be changed by the upload (a stale copy in an SM's L1 would
                       // otherwise survive until that SM solves the neighbour)
};

// QPLayer backward pass (dense/compute_ECJ.hpp:29-190) of the QPs a launch owns: device pointers, batch-major
struct PqpBackwardArgs
{
  const double* loss_derivative; // [B][n + ne + ni] : dL/dx, dL/dy, dL/dz
  double eps, rho_new, mu_new;
  double *dL_dH, *dL_dg, *dL_dA, *dL_db, *dL_dC, *dL_du, *dL_dl; // backward_data.hpp:27-50
};

struct PqpSetupArgs
{
  PqpDims d;
  PqpBatchPtrs p;
  int32_t first, count;
  int32_t execute;           // 1: run Ruiz (EXECUTE); 0: apply stored delta/c (KEEP / IDENTITY)
  int32_t reset_scaling;     // 1: set delta = 1, c = 1 first (IDENTITY)
};

#ifdef __cplusplus
extern "C" {
#endif
// host-side launchers implemented in pqp_kernels.cu
int pqp_launch_setup(const PqpSetupArgs* a, void* stream);
int pqp_launch_solve(const PqpSolveArgs* a, int grid, void* stream);
int pqp_solve_occupancy(const PqpSolveArgs* a, int fused);
int pqp_launch_backward(const PqpSolveArgs* a, const PqpBackwardArgs* k, int grid, void* stream);
int pqp_solve_max_smem(void);
int64_t pqp_setup_smem_bytes(int n, int ne, int ni, int nc);
#ifdef __cplusplus
}
#endif
