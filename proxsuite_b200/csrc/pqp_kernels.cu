// Hand-written sm_100a CUDA kernels of the batched dense ProxQP path.
//
// One CTA (8 warps) owns one QP at a time; CTAs are persistent and pull QPs
// from an atomic work queue (the device-side equivalent of the reference's
// `#pragma omp parallel for schedule(dynamic)`, parallel/qp_solve.hpp:55-59).
//
// Linear algebra (B200-first restructuring of the reference's LDLT, see
// DESIGN.md section 3): the regularised KKT matrix
//     K = [ P   B^T ]     P = H_s + rho I          (n x n, SPD)
//         [ B  -Dlt ]     B = [A_s ; C_s(active)]   Dlt = diag(mu_eq.., mu_in..)
// is factorised as a block LDL^T in the fixed elimination order [x | y, z_act]:
//     P = L1 D1 L1^T,   S = Dlt + B P^-1 B^T = Ls Ds Ls^T      (S is SPD)
// and BOTH unit-triangular factors are kept as explicit INVERSES
// M1 = L1^-1, Ms = Ls^-1 (packed strict lower, shared memory). Every solve is
// then a sequence of triangular mat-vecs (no dependent TRSV chain), a new
// active constraint is a bordering step (two mat-vecs), a removed one is a
// rank-one modification of the trailing rows of Ms driven by a prefix sum, and
// a mu update is a re-bordering from the cached Gram matrix G = B P^-1 B^T.
// This replaces, operation for operation, the reference's
//   Ldlt::factorize / solve_in_place / insert_block_at / delete_at /
//   diagonal_update_clobber_indices   (linalg/dense/ldlt.hpp:340-782)
// while the ProxQP iteration around it follows dense/solver.hpp:1088-1843,
// dense/linesearch.hpp and dense/utils.hpp step by step (cited below).
#include "pqp_device.h"
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <cstdlib>

#define NT PQP_NT
#define NW PQP_NW
#define FULL 0xffffffffu
// tests/emu builds this file with g++ against a functional emulator of the CUDA execution model (cooperative
// fibers for the threads of a CTA); PQP_CPU_EMU is never defined in the product build
#ifdef PQP_CPU_EMU
#define PQP_LAUNCH(kern, grid, block, smem, stream, arg) emu::launch(kern, (int)(grid), (int)(block), (size_t)(smem), arg)
#else
#define PQP_LAUNCH(kern, grid, block, smem, stream, arg) kern<<<grid, block, smem, (cudaStream_t)stream>>>(arg)
#endif

// The solver body is compiled twice: `fastk` assumes that the vector arena and
// the two inverse blocks are in shared memory (the layout chosen whenever they
// fit), which lets ptxas emit LDS/STS with 32-bit addresses instead of generic
// LD/ST; `genk` makes no assumption (large problems spilling to global memory).
namespace setupk {
// dims / pointers are handed over through a shared-memory copy: taking the address of the kernel parameter
// itself makes the compiler keep the whole parameter block addressable, which costs the solve part registers
struct FeedArgs
{
  PqpDims d;
  PqpBatchPtrs p;
  int32_t* ready;
  int32_t batch, fused_setup, feed_margin;
};
__device__ bool feed_and_setup(const FeedArgs* F, int cur_q, int q, double* sm);
}
#define PQP_SM(p) __builtin_assume(__isShared(p))
namespace fastk {
#include "pqp_solver_body.inl"
}
namespace tilek {
#include "pqp_fast_body.inl"
}
#undef PQP_SM
#define PQP_SM(p) ((void)0)
// BIG variant of the tile body (layout kind 2): same driver, Gram precompute and AXPY-form passes, but loop-based
// primitives on packed symmetric storage in the per-CTA global workspace: any n (even), any number of constraints,
// dense / diagonal / zero Hessian, box constraints. Serves the shapes the tile kernel proper cannot hold in shared
// memory (BASELINE cfg 3: dual block 200; cfg 4: n = 256; cfg 5: n = 500, diagonal Hessian, 1000 rows).
#define PQP_BIG 1
namespace bigk {
#include "pqp_fast_body.inl"
}
#undef PQP_BIG
#define PQP_WITH_BACKWARD 1
namespace genk {
#include "pqp_solver_body.inl"
}
#undef PQP_WITH_BACKWARD
#undef PQP_SM

namespace setupk {
using genk::warp_max;
using genk::warp_sum;
// ---------------------------------------------------------------------------
// Set-up of one QP: model -> scaled copies, bound clamping, Ruiz equilibration.
// helpers.hpp:573-666, ruiz.hpp:31-311 (execute) and :425-511 (re-apply).
// One CTA per QP, operating on the scaled arrays in place (L2 resident).
//
// The reference makes two passes over H, A, C per Ruiz iteration (norms, then
// scaling). Here the scaling pass of iteration k also accumulates the column /
// row infinity norms of the values it writes, which ARE the norms iteration
// k + 1 needs, and the very first norms come out of the model -> scaled copy:
// one read-modify-write pass per iteration, same numbers bit for bit.
// ---------------------------------------------------------------------------

struct SetupCtx
{
  int n, ne, ni, nc, box, hess;
  int cstride;    // columns of one block: min(256, n rounded up to 32)
  double *Hs, *As, *Cs, *gs, *bs, *us, *ls, *is, *delta;
  double *dcur;   // n + ne + nc
  double *colmax; // NW * cstride per-warp partial column maxima of one column block
  double *rowmax; // ne + ni
  double *colH;   // n : column norms of H_s (dense)
  double *colAC;  // n : column norms of [A_s; C_s]
  double *red;    // 64
};

__host__ __device__ inline int setup_cstride(int n)
{
  const int r = (n + 31) & ~31;
  return r < 256 ? r : 256;
}
__host__ __device__ inline size_t setup_smem_doubles(int n, int ne, int ni, int nc)
{
  return (size_t)(n + ne + nc) + (size_t)NW * setup_cstride(n) + (size_t)(ne + ni) + 2 * (size_t)n + 64;
}

// combine the per-warp partial column maxima of the block [cb, cb + cw) into out[cb ..]
template<int NCH>
__device__ __forceinline__ void setup_combine_cols(const SetupCtx& s, const double* cm, int cb, int cw, double* out)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cs = s.cstride;
#pragma unroll
  for (int u = 0; u < NCH; ++u) {
    if (32 * u < cs) s.colmax[warp * cs + lane + 32 * u] = cm[u];
  }
  __syncthreads();
  for (int j = tid; j < cw; j += NT) {
    double v = 0;
    for (int w = 0; w < NW; ++w) v = fmax(v, s.colmax[w * cs + j]);
    out[cb + j] = v;
  }
  __syncthreads();
}

// One pass over the rows of a row-major matrix: dst[r][j] = f(src[r][j]) with
//   f(v) = v                       (d == nullptr : the model -> scaled copy)
//   f(v) = (d_r[r] * v) * d[j]     (a Ruiz scaling step; d_r == nullptr means d_r = d)
// accumulating |f| into the caller's column partials and (optionally) rowmax.
// Two rows per warp are in flight so that the loads of the second overlap the first.
template<int NCH>
__device__ __forceinline__ void setup_pass(const SetupCtx& s, const double* src, double* dst, int rows, int cb, int cw, const double* d, const double* d_r, double* cm /*regs*/, double* rowmax_out)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = s.n;
  for (int r = warp; r < rows; r += 2 * NW) {
    const int r2 = r + NW;
    const bool two = r2 < rows;
    const double* a0 = src + (size_t)r * n;
    const double* a1 = src + (size_t)(two ? r2 : r) * n;
    double v0[NCH], v1[NCH];
#pragma unroll
    for (int u = 0; u < NCH; ++u) {
      const int j = cb + lane + 32 * u;
      const bool in = j < cb + cw;
      // L2 loads: a neighbouring QP's pass may have left a stale copy of a shared 128-byte line in
      // this SM's L1 before the upload of this QP's inputs landed (fused feed, see feed_and_setup)
      v0[u] = in ? __ldcg(a0 + j) : 0.0;
      v1[u] = (in && two) ? __ldcg(a1 + j) : 0.0;
    }
    double dr0 = 1.0, dr1 = 1.0;
    if (d) {
      const double* dd = d_r ? d_r : d;
      dr0 = dd[r];
      dr1 = two ? dd[r2] : 1.0;
    }
    double rm0 = 0, rm1 = 0;
    double* o0 = dst + (size_t)r * n;
    double* o1 = dst + (size_t)(two ? r2 : r) * n;
#pragma unroll
    for (int u = 0; u < NCH; ++u) {
      const int j = cb + lane + 32 * u;
      if (j < cb + cw) {
        double w0 = v0[u], w1 = v1[u];
        if (d) {
          const double dj = d[j];
          w0 = dr0 * w0 * dj;
          w1 = dr1 * w1 * dj;
        }
        o0[j] = w0;
        const double f0 = fabs(w0);
        cm[u] = fmax(cm[u], f0);
        rm0 = fmax(rm0, f0);
        if (two) {
          o1[j] = w1;
          const double f1 = fabs(w1);
          cm[u] = fmax(cm[u], f1);
          rm1 = fmax(rm1, f1);
        }
      }
    }
    if (rowmax_out) {
      rm0 = warp_max(rm0);
      rm1 = warp_max(rm1);
      if (lane == 0) {
        rowmax_out[r] = fmax(rowmax_out[r], rm0);
        if (two) rowmax_out[r2] = fmax(rowmax_out[r2], rm1);
      }
    }
  }
}

// The same pass for even n <= 256 (one column block): a lane owns column PAIRS (16-byte L2 loads / stores on
// pointers known to be global), the scaling vector is read from shared memory as pairs, and the three variants
// (copy, scale, scale + row norms) are separate instantiations without run-time branches.
// max of two non-negative finite numbers: one compare + select (fmax adds NaN handling)
__device__ __forceinline__ double pmax(double a, double b)
{
  return b > a ? b : a;
}
template<int NP, bool SCALE, bool ROWMAX>
__device__ __forceinline__ void setup_pass2(const double* src, double* dst, int rows, int n, const double* d, const double* d_r, double2* cm /*regs*/, double* rowmax_out)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int np = n >> 1;
  const double2* d2 = reinterpret_cast<const double2*>(d);
  _Pragma("unroll 1") for (int r = warp; r < rows; r += 2 * NW) {
    const int r2 = r + NW;
    const bool two = r2 < rows;
    const double2* a0 = reinterpret_cast<const double2*>(src + (size_t)r * n);
    const double2* a1 = reinterpret_cast<const double2*>(src + (size_t)(two ? r2 : r) * n);
    double2 v0[NP], v1[NP];
#pragma unroll
    for (int u = 0; u < NP; ++u) {
      const int p = lane + 32 * u;
      const bool in = p < np;
      v0[u] = in ? __ldcg(a0 + p) : make_double2(0.0, 0.0);
      v1[u] = (in && two) ? __ldcg(a1 + p) : make_double2(0.0, 0.0);
    }
    double dr0 = 1.0, dr1 = 1.0;
    if (SCALE) {
      dr0 = d_r[r];
      dr1 = two ? d_r[r2] : 1.0;
    }
    double rm0 = 0, rm1 = 0;
    double2* o0 = reinterpret_cast<double2*>(dst + (size_t)r * n);
    double2* o1 = reinterpret_cast<double2*>(dst + (size_t)(two ? r2 : r) * n);
#pragma unroll
    for (int u = 0; u < NP; ++u) {
      const int p = lane + 32 * u;
      if (p < np) {
        double2 w0 = v0[u], w1 = v1[u];
        if (SCALE) {
          const double2 dj = d2[p];
          w0.x = dr0 * w0.x * dj.x;
          w0.y = dr0 * w0.y * dj.y;
          w1.x = dr1 * w1.x * dj.x;
          w1.y = dr1 * w1.y * dj.y;
        }
        __stcg(o0 + p, w0);
        const double f0x = fabs(w0.x), f0y = fabs(w0.y);
        cm[u].x = pmax(cm[u].x, f0x);
        cm[u].y = pmax(cm[u].y, f0y);
        if (ROWMAX) rm0 = pmax(rm0, pmax(f0x, f0y));
        if (two) {
          __stcg(o1 + p, w1);
          const double f1x = fabs(w1.x), f1y = fabs(w1.y);
          cm[u].x = pmax(cm[u].x, f1x);
          cm[u].y = pmax(cm[u].y, f1y);
          if (ROWMAX) rm1 = pmax(rm1, pmax(f1x, f1y));
        }
      }
    }
    if (ROWMAX) {
      rm0 = warp_max(rm0);
      rm1 = warp_max(rm1);
      if (lane == 0) {
        rowmax_out[r] = pmax(rowmax_out[r], rm0);
        if (two) rowmax_out[r2] = pmax(rowmax_out[r2], rm1);
      }
    }
  }
}
// per-warp column partials of the pair layout -> out[0 .. n)
template<int NP>
__device__ __forceinline__ void setup_combine_cols2(const SetupCtx& s, const double2* cm, int n, double* out)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cs = s.cstride, np = n >> 1;
  double* const colmax = s.colmax;
  __builtin_assume(__isShared(colmax));
  __builtin_assume(__isShared(out));
#pragma unroll
  for (int u = 0; u < NP; ++u) {
    const int p = lane + 32 * u;
    if (p < np) { // colmax is only 8-byte aligned (it follows n + ne + nc doubles)
      colmax[warp * cs + 2 * p] = cm[u].x;
      colmax[warp * cs + 2 * p + 1] = cm[u].y;
    }
  }
  __syncthreads();
  for (int j = tid; j < n; j += NT) {
    double v = 0;
    for (int w = 0; w < NW; ++w) v = fmax(v, colmax[w * cs + j]);
    out[j] = v;
  }
  __syncthreads();
}

// `sm` : setup_smem_doubles(...) doubles of shared memory; execute: 1 = run Ruiz, 0 = apply the stored delta / c
template<int NCH>
__device__ void setup_one_t(const PqpDims& D, const PqpBatchPtrs& P, int q, int execute, int reset_scaling, double* sm)
{
  __shared__ SetupCtx s;
  const int n = D.n, ne = D.ne, ni = D.ni, nc = D.nc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nd = n + ne + nc;
  const double machine_eps = 2.220446049250313e-16;
  __syncthreads(); // the previous user of `s` / `sm` is done
  if (tid == 0) {
    s.n = n;
    s.ne = ne;
    s.ni = ni;
    s.nc = nc;
    s.box = D.box;
    s.hess = D.hess;
    s.cstride = setup_cstride(n);
    s.Hs = P.Hs + (size_t)q * n * n;
    s.As = P.As + (size_t)q * ne * n;
    s.Cs = P.Cs + (size_t)q * ni * n;
    s.gs = P.gs + (size_t)q * n;
    s.bs = P.bs + (size_t)q * ne;
    s.us = P.us + (size_t)q * nc;
    s.ls = P.ls + (size_t)q * nc;
    s.is = P.is + (size_t)q * n;
    s.delta = P.delta + (size_t)q * nd;
    s.dcur = sm;
    s.colmax = sm + nd;
    s.rowmax = s.colmax + NW * s.cstride;
    s.colH = s.rowmax + ne + ni;
    s.colAC = s.colH + n;
    s.red = s.colAC + n;
  }
  __syncthreads();
  const double* Hm = P.H + (size_t)q * n * n;
  const double* Am = P.A + (size_t)q * ne * n;
  const double* Cm = P.C + (size_t)q * ni * n;
  for (int r = tid; r < ne + ni; r += NT) s.rowmax[r] = 0.0;
  __syncthreads();
  // scaled <- model (helpers.hpp:614-649), with the norms of the first Ruiz iteration
  if (D.hess == PQP_HESSIAN_ZERO) {
    for (size_t i = tid; i < (size_t)n * n; i += NT) s.Hs[i] = 0.0;
  }
  // pair layout: even n, one column block, 16-byte aligned rows (cudaMalloc'ed bases, even row length)
  constexpr int NP = (NCH + 1) / 2;
  const bool vec2 = (n % 2 == 0) && n <= 256;
  double* const gHs = P.Hs + (size_t)q * n * n;
  double* const gAs = P.As + (size_t)q * ne * n;
  double* const gCs = P.Cs + (size_t)q * ni * n;
  if (vec2) {
    double2 cm2[NP];
#pragma unroll
    for (int u = 0; u < NP; ++u) cm2[u] = make_double2(0.0, 0.0);
    if (D.hess != PQP_HESSIAN_ZERO) {
      setup_pass2<NP, false, false>(Hm, gHs, n, n, sm, sm, cm2, sm);
      setup_combine_cols2<NP>(s, cm2, n, s.colH);
#pragma unroll
      for (int u = 0; u < NP; ++u) cm2[u] = make_double2(0.0, 0.0);
    }
    setup_pass2<NP, false, true>(Am, gAs, ne, n, sm, sm, cm2, s.rowmax);
    setup_pass2<NP, false, true>(Cm, gCs, ni, n, sm, sm, cm2, s.rowmax + ne);
    setup_combine_cols2<NP>(s, cm2, n, s.colAC);
  } else {
  for (int cb = 0; cb < n; cb += 256) {
    const int cw = min(256, n - cb);
    double cm[NCH];
#pragma unroll
    for (int u = 0; u < NCH; ++u) cm[u] = 0.0;
    if (D.hess != PQP_HESSIAN_ZERO) {
      setup_pass<NCH>(s, Hm, s.Hs, n, cb, cw, nullptr, nullptr, cm, nullptr);
      setup_combine_cols<NCH>(s, cm, cb, cw, s.colH);
#pragma unroll
      for (int u = 0; u < NCH; ++u) cm[u] = 0.0;
    }
    setup_pass<NCH>(s, Am, s.As, ne, cb, cw, nullptr, nullptr, cm, s.rowmax);
    setup_pass<NCH>(s, Cm, s.Cs, ni, cb, cw, nullptr, nullptr, cm, s.rowmax + ne);
    setup_combine_cols<NCH>(s, cm, cb, cw, s.colAC);
  }
  }
  for (int j = tid; j < n; j += NT) s.gs[j] = __ldcg(P.g + (size_t)q * n + j);
  for (int j = tid; j < ne; j += NT) s.bs[j] = __ldcg(P.b + (size_t)q * ne + j);
  for (int j = tid; j < nc; j += NT) {
    double u = (j < ni) ? __ldcg(P.u + (size_t)q * ni + j) : __ldcg(P.u_box + (size_t)q * n + j - ni);
    double l = (j < ni) ? __ldcg(P.l + (size_t)q * ni + j) : __ldcg(P.l_box + (size_t)q * n + j - ni);
    s.us[j] = u <= 1e20 ? u : 1e20;
    s.ls[j] = l >= -1e20 ? l : -1e20;
  }
  // i_scaled always restarts from ones (the reference only does so when the
  // preconditioner is executed and lets it drift on re-application; see
  // DESIGN.md "deliberate deviations")
  for (int j = tid; j < n; j += NT) s.is[j] = 1.0;
  if (reset_scaling || execute) {
    for (int j = tid; j < nd; j += NT) s.delta[j] = 1.0;
    if (tid == 0) P.c[q] = 1.0;
  }
  __syncthreads();

  auto scale_vecs = [&](const double* d) {
    for (int j = tid; j < n; j += NT) s.gs[j] *= d[j];
    for (int j = tid; j < ne; j += NT) s.bs[j] *= d[n + j];
    for (int j = tid; j < nc; j += NT) {
      s.us[j] *= d[n + ne + j];
      s.ls[j] *= d[n + ne + j];
    }
    if (D.box) {
      for (int j = tid; j < n; j += NT) {
        s.is[j] *= d[j];
        s.is[j] *= d[n + ne + ni + j];
      }
    }
  };

  if (!execute) {
    // ruiz.hpp:425-511: re-apply the stored scaling
    const double cq = P.c[q];
    for (int j = tid; j < nd; j += NT) s.dcur[j] = s.delta[j];
    __syncthreads();
    for (int r = warp; r < ne + ni; r += NW) {
      double* row = (r < ne) ? s.As + (size_t)r * n : s.Cs + (size_t)(r - ne) * n;
      const double dr = s.dcur[n + r];
      for (int j = lane; j < n; j += 32) row[j] = dr * row[j] * s.dcur[j];
    }
    if (D.hess == PQP_HESSIAN_DENSE) {
      for (int r = warp; r < n; r += NW) {
        double* row = s.Hs + (size_t)r * n;
        const double dr = s.dcur[r];
        for (int j = lane; j < n; j += 32) row[j] = (dr * row[j] * s.dcur[j]) * cq;
      }
    } else if (D.hess == PQP_HESSIAN_DIAGONAL) {
      for (int j = tid; j < n; j += NT) {
        double h = s.Hs[(size_t)j * n + j];
        h *= s.dcur[j];
        h *= s.dcur[j];
        s.Hs[(size_t)j * n + j] = h * cq;
      }
    }
    scale_vecs(s.dcur);
    __syncthreads();
    for (int j = tid; j < n; j += NT) s.gs[j] *= cq;
    __syncthreads();
    return;
  }

  // ruiz.hpp:31-311
  const PqpQpParams& prm = P.params[q];
  const long long max_iter = prm.s.preconditioner_max_iter;
  const double epsilon = prm.s.preconditioner_accuracy;
  const bool for_infeasible = prm.s.primal_infeasibility_solving != 0;
  for (int j = tid; j < nd; j += NT) s.dcur[j] = 0.0;
  double cacc = 1.0;
  __syncthreads();
  long long iter = 1;
  while (true) {
    double m = 0;
    for (int j = tid; j < nd; j += NT) m = fmax(m, fabs(1.0 - s.dcur[j]));
    m = warp_max(m);
    if (lane == 0) s.red[warp] = m;
    __syncthreads();
    m = s.red[0];
    for (int w = 1; w < NW; ++w) m = fmax(m, s.red[w]);
    __syncthreads();
    if (!(m > epsilon)) break;
    if (iter == max_iter) break;
    ++iter;
    // --- scaling factors from the norms of the current matrices (ruiz.hpp:68-173)
    for (int k = tid; k < n; k += NT) {
      double v = s.colAC[k];
      if (D.hess == PQP_HESSIAN_DENSE) v = fmax(v, s.colH[k]);
      if (D.hess == PQP_HESSIAN_DIAGONAL) v = fmax(v, fabs(s.Hs[(size_t)k * n + k]));
      if (D.box) v = fmax(v, s.is[k]);
      const double aux = sqrt(v);
      s.dcur[k] = (aux == 0.0) ? 1.0 : 1.0 / (aux + machine_eps);
    }
    if (for_infeasible) {
      for (int j = tid; j < ne + nc; j += NT) s.dcur[n + j] = 1.0;
    } else {
      for (int r = tid; r < ne + ni; r += NT) {
        const double aux = sqrt(s.rowmax[r]);
        s.dcur[n + r] = (aux == 0.0) ? 1.0 : 1.0 / (aux + machine_eps);
      }
      if (D.box) {
        for (int k = tid; k < n; k += NT) s.dcur[n + ne + ni + k] = 1.0 / sqrt(s.is[k] + machine_eps);
      }
    }
    __syncthreads();
    for (int r = tid; r < ne + ni; r += NT) s.rowmax[r] = 0.0;
    __syncthreads();
    // --- scale (ruiz.hpp:175-290); the values written are the next iteration's norms
    double colsum = 0;
    if (vec2) {
      double2 cm2[NP];
#pragma unroll
      for (int u = 0; u < NP; ++u) cm2[u] = make_double2(0.0, 0.0);
      setup_pass2<NP, true, true>(gAs, gAs, ne, n, s.dcur, s.dcur + n, cm2, s.rowmax);
      setup_pass2<NP, true, true>(gCs, gCs, ni, n, s.dcur, s.dcur + n + ne, cm2, s.rowmax + ne);
      setup_combine_cols2<NP>(s, cm2, n, s.colAC);
      if (D.hess == PQP_HESSIAN_DENSE) {
#pragma unroll
        for (int u = 0; u < NP; ++u) cm2[u] = make_double2(0.0, 0.0);
        setup_pass2<NP, true, false>(gHs, gHs, n, n, s.dcur, s.dcur, cm2, sm);
        setup_combine_cols2<NP>(s, cm2, n, s.colH);
        double part = 0;
        for (int j = tid; j < n; j += NT) part += s.colH[j];
        part = warp_sum(part);
        if (lane == 0) s.red[warp] = part;
        __syncthreads();
        for (int w = 0; w < NW; ++w) colsum += s.red[w];
        __syncthreads();
      }
    } else
    for (int cb = 0; cb < n; cb += 256) {
      const int cw = min(256, n - cb);
      double cm[NCH];
#pragma unroll
      for (int u = 0; u < NCH; ++u) cm[u] = 0.0;
      setup_pass<NCH>(s, s.As, s.As, ne, cb, cw, s.dcur, s.dcur + n, cm, s.rowmax);
      setup_pass<NCH>(s, s.Cs, s.Cs, ni, cb, cw, s.dcur, s.dcur + n + ne, cm, s.rowmax + ne);
      setup_combine_cols<NCH>(s, cm, cb, cw, s.colAC);
      if (D.hess == PQP_HESSIAN_DENSE) {
#pragma unroll
        for (int u = 0; u < NCH; ++u) cm[u] = 0.0;
        setup_pass<NCH>(s, s.Hs, s.Hs, n, cb, cw, s.dcur, nullptr, cm, nullptr);
        setup_combine_cols<NCH>(s, cm, cb, cw, s.colH);
        double part = 0;
        for (int j = tid; j < cw; j += NT) part += s.colH[cb + j];
        part = warp_sum(part);
        if (lane == 0) s.red[warp] = part;
        __syncthreads();
        for (int w = 0; w < NW; ++w) colsum += s.red[w];
        __syncthreads();
      }
    }
    scale_vecs(s.dcur);
    double gamma = 1.0;
    if (D.hess == PQP_HESSIAN_DENSE) {
      gamma = 1.0 / fmax(1.0, colsum / (double)n);
      // quirk 1: H itself is NOT multiplied by gamma here
    } else if (D.hess == PQP_HESSIAN_DIAGONAL) {
      double dm = 0;
      for (int j = tid; j < n; j += NT) {
        double h = s.Hs[(size_t)j * n + j];
        h *= s.dcur[j];
        h *= s.dcur[j];
        s.Hs[(size_t)j * n + j] = h;
        dm = fmax(dm, fabs(h));
      }
      dm = warp_max(dm);
      if (lane == 0) s.red[warp] = dm;
      __syncthreads();
      dm = s.red[0];
      for (int w = 1; w < NW; ++w) dm = fmax(dm, s.red[w]);
      __syncthreads();
      gamma = 1.0 / fmax(1.0, dm / (double)n);
      for (int j = tid; j < n; j += NT) s.Hs[(size_t)j * n + j] *= gamma;
    }
    __syncthreads();
    for (int j = tid; j < n; j += NT) s.gs[j] *= gamma;
    for (int j = tid; j < nd; j += NT) s.delta[j] *= s.dcur[j];
    cacc *= gamma;
    __syncthreads();
  }
  if (tid == 0) P.c[q] = cacc;
  __syncthreads();
}

// column chunks per lane: n <= 32 -> 1, <= 64 -> 2, <= 128 -> 4, else 8 (blocks of 256 columns)
__device__ __noinline__ void setup_one(const PqpDims& D, const PqpBatchPtrs& P, int q, int execute, int reset_scaling, double* sm)
{
  if (D.n <= 32)
    setup_one_t<1>(D, P, q, execute, reset_scaling, sm);
  else if (D.n <= 64)
    setup_one_t<2>(D, P, q, execute, reset_scaling, sm);
  else if (D.n <= 128)
    setup_one_t<4>(D, P, q, execute, reset_scaling, sm);
  else
    setup_one_t<8>(D, P, q, execute, reset_scaling, sm);
}

__global__ void __launch_bounds__(NT) pqp_setup_kernel(const __grid_constant__ PqpSetupArgs A)
{
#ifdef PQP_CPU_EMU
  double* const setup_sm = emu::dyn_smem;
#else
  extern __shared__ __align__(16) double setup_sm[];
#endif
  setup_one(A.d, A.p, A.first + blockIdx.x, A.execute, A.reset_scaling, setup_sm);
}

// Feed gate + fused set-up of the persistent solve kernels (declared in pqp_device.h terms):
// a QP is consumed only once the host -> device upload of its inputs has been
// announced through A.ready[0] (written by a 4-byte copy that follows the
// chunk's data on the upload stream); then the CTA that owns the QP equilibrates
// it and goes straight on to the solve. Returns false when the feed was aborted.
__device__ __noinline__ bool feed_and_setup(const FeedArgs* F, int cur_q, int q, double* sm)
{
  const FeedArgs& A = *F;
  __shared__ int feed_ok;
  if (A.ready) {
    if (threadIdx.x == 0) {
      volatile int32_t* rdy = A.ready;
      int ok = 1;
      int need = cur_q + 1 + A.feed_margin;
      if (need > A.batch) need = A.batch;
      if (rdy[0] < need) {
        unsigned long long t0;
#ifdef PQP_CPU_EMU
        t0 = emu::globaltimer();
#else
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#endif
        while (rdy[0] < need) {
          if (rdy[1] != 0) {
            ok = 0;
            break;
          }
          unsigned long long t1;
#ifdef PQP_CPU_EMU
          t1 = emu::globaltimer();
#else
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
#endif
          if (t1 - t0 > 20000000000ull) { // 20 s without progress: the upload will never come
            rdy[1] = 1;
            ok = 0;
            break;
          }
          __nanosleep(200);
        }
      }
      __threadfence();
      feed_ok = ok;
    }
    __syncthreads();
    const bool ok = feed_ok != 0;
    __syncthreads();
    if (!ok) {
      if (threadIdx.x == 0) A.p.info[(size_t)q * PQP_INFO_DOUBLES + 10] = (double)PQP_NOT_RUN;
      return false;
    }
  }
#ifndef PQP_EXPERIMENT_EMPTY_SETUP
  if (A.fused_setup && A.p.params[q].active) setup_one(A.d, A.p, q, (A.fused_setup & 1) ? 1 : 0, (A.fused_setup & 4) ? 1 : 0, sm);
#endif
  return true;
}

} // namespace setupk
using setupk::pqp_setup_kernel;

extern "C" int64_t
pqp_setup_smem_bytes(int n, int ne, int ni, int nc)
{
  return (int64_t)(sizeof(double) * setupk::setup_smem_doubles(n, ne, ni, nc));
}

extern "C" int
pqp_solve_max_smem(void)
{
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
  return v;
}

extern "C" int
pqp_launch_setup(const PqpSetupArgs* a, void* stream)
{
  if (a->count <= 0) return 0;
  size_t smem = sizeof(double) * setupk::setup_smem_doubles(a->d.n, a->d.ne, a->d.ni, a->d.nc);
  cudaError_t e = cudaFuncSetAttribute(pqp_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  PQP_LAUNCH(pqp_setup_kernel, a->count, NT, smem, stream, *a);
  return (int)cudaGetLastError();
}

// the backward pass runs on the general kernel body (any shape, full-capacity layout)
extern "C" int
pqp_launch_backward(const PqpSolveArgs* a, const PqpBackwardArgs* k, int grid, void* stream)
{
  size_t smem = sizeof(double) * (size_t)a->lay.smem_doubles + (size_t)a->lay.smem_int_bytes;
  cudaError_t e = cudaFuncSetAttribute(genk::pqp_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
#ifdef PQP_CPU_EMU
  emu::launch2(genk::pqp_backward_kernel, grid, NT, smem, *a, *k);
#else
  genk::pqp_backward_kernel<<<grid, NT, smem, (cudaStream_t)stream>>>(*a, *k);
#endif
  return (int)cudaGetLastError();
}

// Resident CTAs per SM the runtime grants the solve kernel this layout selects (plain or fused instantiation): the
// layout budgets of pqp_capi.cu count on two; a kernel whose static shared memory pushes it 16 bytes over half an SM
// silently runs at one (round 2: the fused tile kernel did, 56 ms instead of 36 per 4096 QPs).
extern "C" int
pqp_solve_occupancy(const PqpSolveArgs* a, int fused)
{
#ifdef PQP_CPU_EMU
  (void)fused;
  return a->lay.ctas_per_sm;
#else
  size_t smem = sizeof(double) * (size_t)a->lay.smem_doubles + (size_t)a->lay.smem_int_bytes;
  const int64_t symn = (int64_t)a->d.n * (a->d.n + 1) / 2, symc = (int64_t)a->lay.si_cap * (a->lay.si_cap + 1) / 2;
  const bool fast = a->lay.in_smem[PA_VEC] && a->lay.in_smem[PA_MS] && (a->lay.in_smem[PA_M1] || a->d.hess != PQP_HESSIAN_DENSE || symn <= symc);
  auto kern = fused ? ((a->lay.kind == 1) ? tilek::pqp_solve_kernel_fused : (a->lay.kind == 2) ? bigk::pqp_solve_kernel_fused : (fast ? fastk::pqp_solve_kernel_fused : genk::pqp_solve_kernel_fused))
                    : ((a->lay.kind == 1) ? tilek::pqp_solve_kernel : (a->lay.kind == 2) ? bigk::pqp_solve_kernel : (fast ? fastk::pqp_solve_kernel : genk::pqp_solve_kernel));
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, NT, smem) != cudaSuccess) return -1;
  return nb;
#endif
}

extern "C" int
pqp_launch_solve(const PqpSolveArgs* a, int grid, void* stream)
{
  size_t smem = sizeof(double) * (size_t)a->lay.smem_doubles + (size_t)a->lay.smem_int_bytes;
  // fast kernel: vectors and S^-1 in shared memory; P^-1 either there too or swept inside the S^-1 region
  const int64_t symn = (int64_t)a->d.n * (a->d.n + 1) / 2, symc = (int64_t)a->lay.si_cap * (a->lay.si_cap + 1) / 2;
  const bool fast = a->lay.in_smem[PA_VEC] && a->lay.in_smem[PA_MS] && (a->lay.in_smem[PA_M1] || a->d.hess != PQP_HESSIAN_DENSE || symn <= symc);
  static const bool force_fused = std::getenv("PQP_FORCE_FUSED_KERNEL") != nullptr; // A/B hook: time the <1> instantiation on resident data
  const bool fused = a->ready != nullptr || a->fused_setup != 0 || force_fused;
  auto kern = fused ? ((a->lay.kind == 1) ? tilek::pqp_solve_kernel_fused : (a->lay.kind == 2) ? bigk::pqp_solve_kernel_fused : (fast ? fastk::pqp_solve_kernel_fused : genk::pqp_solve_kernel_fused))
                    : ((a->lay.kind == 1) ? tilek::pqp_solve_kernel : (a->lay.kind == 2) ? bigk::pqp_solve_kernel : (fast ? fastk::pqp_solve_kernel : genk::pqp_solve_kernel));
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  PQP_LAUNCH(kern, grid, NT, smem, stream, *a);
  return (int)cudaGetLastError();
}
