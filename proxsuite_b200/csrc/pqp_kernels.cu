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

#define NT PQP_NT
#define NW PQP_NW
#define FULL 0xffffffffu

// The solver body is compiled twice: `fastk` assumes that the vector arena and
// the two inverse blocks are in shared memory (the layout chosen whenever they
// fit), which lets ptxas emit LDS/STS with 32-bit addresses instead of generic
// LD/ST; `genk` makes no assumption (large problems spilling to global memory).
#define PQP_SM(p) __builtin_assume(__isShared(p))
namespace fastk {
#include "pqp_solver_body.inl"
}
namespace tilek {
#include "pqp_fast_body.inl"
}
#undef PQP_SM
#define PQP_SM(p) ((void)0)
namespace genk {
#include "pqp_solver_body.inl"
}
#undef PQP_SM

namespace {
using genk::warp_max;
using genk::warp_sum;
// ---------------------------------------------------------------------------
// Set-up kernel: model -> scaled copies, bound clamping, Ruiz equilibration.
// helpers.hpp:573-666, ruiz.hpp:31-311 (execute) and :425-511 (re-apply).
// One CTA per QP, operating on the L2-resident scaled arrays in place.
// ---------------------------------------------------------------------------
#define SETUP_CH 8 // columns per lane per chunk (32 * 8 = 256 columns)

struct SetupCtx
{
  int n, ne, ni, nc, box, hess;
  double *Hs, *As, *Cs, *gs, *bs, *us, *ls, *is, *delta;
  double *dcur;   // n + ne + nc
  double *colmax; // NW * 256
  double *rowmax; // ne + ni
  double *red;    // 64
};

// column / row infinity norms of one row-major matrix, accumulated into
// colmax (per-warp partials) and rowmax
__device__ void setup_norms(const SetupCtx& s, const double* M, int rows, int cb, int cw, double* cm /*regs*/, double* rowmax_out)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < rows; r += NW) {
    const double* row = M + (size_t)r * s.n;
    double rm = 0;
#pragma unroll
    for (int u = 0; u < SETUP_CH; ++u) {
      int j = cb + lane + 32 * u;
      if (j < cb + cw) {
        double v = fabs(row[j]);
        cm[u] = fmax(cm[u], v);
        rm = fmax(rm, v);
      }
    }
    if (rowmax_out) {
      rm = warp_max(rm);
      if (lane == 0) rowmax_out[r] = fmax(rowmax_out[r], rm);
    }
  }
}

__global__ void __launch_bounds__(NT) pqp_setup_kernel(PqpSetupArgs A)
{
  extern __shared__ double sm[];
  __shared__ SetupCtx s;
  const int q = A.first + blockIdx.x;
  const int n = A.d.n, ne = A.d.ne, ni = A.d.ni, nc = A.d.nc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nd = n + ne + nc;
  const double machine_eps = 2.220446049250313e-16;
  if (tid == 0) {
    s.n = n;
    s.ne = ne;
    s.ni = ni;
    s.nc = nc;
    s.box = A.d.box;
    s.hess = A.d.hess;
    s.Hs = A.p.Hs + (size_t)q * n * n;
    s.As = A.p.As + (size_t)q * ne * n;
    s.Cs = A.p.Cs + (size_t)q * ni * n;
    s.gs = A.p.gs + (size_t)q * n;
    s.bs = A.p.bs + (size_t)q * ne;
    s.us = A.p.us + (size_t)q * nc;
    s.ls = A.p.ls + (size_t)q * nc;
    s.is = A.p.is + (size_t)q * n;
    s.delta = A.p.delta + (size_t)q * nd;
    s.dcur = sm;
    s.colmax = sm + nd;
    s.rowmax = s.colmax + NW * 256;
    s.red = s.rowmax + ne + ni;
  }
  __syncthreads();
  const double* Hm = A.p.H + (size_t)q * n * n;
  const double* Am = A.p.A + (size_t)q * ne * n;
  const double* Cm = A.p.C + (size_t)q * ni * n;
  // scaled <- model (helpers.hpp:614-649)
  if (A.d.hess != PQP_HESSIAN_ZERO) {
    for (size_t i = tid; i < (size_t)n * n; i += NT) s.Hs[i] = Hm[i];
  } else {
    for (size_t i = tid; i < (size_t)n * n; i += NT) s.Hs[i] = 0.0;
  }
  for (size_t i = tid; i < (size_t)ne * n; i += NT) s.As[i] = Am[i];
  for (size_t i = tid; i < (size_t)ni * n; i += NT) s.Cs[i] = Cm[i];
  for (int j = tid; j < n; j += NT) s.gs[j] = A.p.g[(size_t)q * n + j];
  for (int j = tid; j < ne; j += NT) s.bs[j] = A.p.b[(size_t)q * ne + j];
  for (int j = tid; j < nc; j += NT) {
    double u = (j < ni) ? A.p.u[(size_t)q * ni + j] : A.p.u_box[(size_t)q * n + j - ni];
    double l = (j < ni) ? A.p.l[(size_t)q * ni + j] : A.p.l_box[(size_t)q * n + j - ni];
    s.us[j] = u <= 1e20 ? u : 1e20;
    s.ls[j] = l >= -1e20 ? l : -1e20;
  }
  // i_scaled always restarts from ones (the reference only does so when the
  // preconditioner is executed and lets it drift on re-application; see
  // DESIGN.md "deliberate deviations")
  for (int j = tid; j < n; j += NT) s.is[j] = 1.0;
  if (A.reset_scaling || A.execute) {
    for (int j = tid; j < nd; j += NT) s.delta[j] = 1.0;
    if (tid == 0) A.p.c[q] = 1.0;
  }
  __syncthreads();

  auto scale_AC = [&](const double* d) {
    for (int r = warp; r < ne + ni; r += NW) {
      double* row = (r < ne) ? s.As + (size_t)r * n : s.Cs + (size_t)(r - ne) * n;
      const double dr = d[n + r];
      for (int j = lane; j < n; j += 32) row[j] = dr * row[j] * d[j];
    }
  };
  auto scale_vecs = [&](const double* d) {
    for (int j = tid; j < n; j += NT) s.gs[j] *= d[j];
    for (int j = tid; j < ne; j += NT) s.bs[j] *= d[n + j];
    for (int j = tid; j < nc; j += NT) {
      s.us[j] *= d[n + ne + j];
      s.ls[j] *= d[n + ne + j];
    }
    if (A.d.box) {
      for (int j = tid; j < n; j += NT) {
        s.is[j] *= d[j];
        s.is[j] *= d[n + ne + ni + j];
      }
    }
  };

  if (!A.execute) {
    // ruiz.hpp:425-511: re-apply the stored scaling
    const double cq = A.p.c[q];
    for (int j = tid; j < nd; j += NT) s.dcur[j] = s.delta[j];
    __syncthreads();
    scale_AC(s.dcur);
    if (A.d.hess == PQP_HESSIAN_DENSE) {
      for (int r = warp; r < n; r += NW) {
        double* row = s.Hs + (size_t)r * n;
        const double dr = s.dcur[r];
        for (int j = lane; j < n; j += 32) row[j] = (dr * row[j] * s.dcur[j]) * cq;
      }
    } else if (A.d.hess == PQP_HESSIAN_DIAGONAL) {
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
    return;
  }

  // ruiz.hpp:31-311
  const PqpQpParams& prm = A.p.params[q];
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
    // --- norms of the current matrices
    for (int r = tid; r < ne + ni; r += NT) s.rowmax[r] = 0.0;
    __syncthreads();
    for (int cb = 0; cb < n; cb += 256) {
      const int cw = min(256, n - cb);
      double cm[SETUP_CH];
#pragma unroll
      for (int u = 0; u < SETUP_CH; ++u) cm[u] = 0.0;
      if (A.d.hess == PQP_HESSIAN_DENSE) setup_norms(s, s.Hs, n, cb, cw, cm, nullptr);
      setup_norms(s, s.As, ne, cb, cw, cm, s.rowmax);
      setup_norms(s, s.Cs, ni, cb, cw, cm, s.rowmax + ne);
#pragma unroll
      for (int u = 0; u < SETUP_CH; ++u) s.colmax[warp * 256 + lane + 32 * u] = cm[u];
      __syncthreads();
      for (int j = tid; j < cw; j += NT) {
        double v = 0;
        for (int w = 0; w < NW; ++w) v = fmax(v, s.colmax[w * 256 + j]);
        const int k = cb + j;
        if (A.d.hess == PQP_HESSIAN_DIAGONAL) v = fmax(v, fabs(s.Hs[(size_t)k * n + k]));
        if (A.d.box) v = fmax(v, s.is[k]);
        const double aux = sqrt(v);
        s.dcur[k] = (aux == 0.0) ? 1.0 : 1.0 / (aux + machine_eps);
      }
      __syncthreads();
    }
    if (for_infeasible) {
      for (int j = tid; j < ne + nc; j += NT) s.dcur[n + j] = 1.0;
    } else {
      for (int r = tid; r < ne + ni; r += NT) {
        const double aux = sqrt(s.rowmax[r]);
        s.dcur[n + r] = (aux == 0.0) ? 1.0 : 1.0 / (aux + machine_eps);
      }
      if (A.d.box) {
        for (int k = tid; k < n; k += NT) s.dcur[n + ne + ni + k] = 1.0 / sqrt(s.is[k] + machine_eps);
      }
    }
    __syncthreads();
    // --- scale
    scale_AC(s.dcur);
    scale_vecs(s.dcur);
    double gamma = 1.0;
    if (A.d.hess == PQP_HESSIAN_DENSE) {
      double colsum = 0;
      for (int cb = 0; cb < n; cb += 256) {
        const int cw = min(256, n - cb);
        double cm[SETUP_CH];
#pragma unroll
        for (int u = 0; u < SETUP_CH; ++u) cm[u] = 0.0;
        for (int r = warp; r < n; r += NW) {
          double* row = s.Hs + (size_t)r * n;
          const double dr = s.dcur[r];
#pragma unroll
          for (int u = 0; u < SETUP_CH; ++u) {
            int j = cb + lane + 32 * u;
            if (j < cb + cw) {
              double v = dr * row[j] * s.dcur[j];
              row[j] = v;
              cm[u] = fmax(cm[u], fabs(v));
            }
          }
        }
#pragma unroll
        for (int u = 0; u < SETUP_CH; ++u) s.colmax[warp * 256 + lane + 32 * u] = cm[u];
        __syncthreads();
        double part = 0;
        for (int j = tid; j < cw; j += NT) {
          double v = 0;
          for (int w = 0; w < NW; ++w) v = fmax(v, s.colmax[w * 256 + j]);
          part += v;
        }
        part = warp_sum(part);
        if (lane == 0) s.red[warp] = part;
        __syncthreads();
        for (int w = 0; w < NW; ++w) colsum += s.red[w];
        __syncthreads();
      }
      gamma = 1.0 / fmax(1.0, colsum / (double)n);
      // quirk 1: H itself is NOT multiplied by gamma here
    } else if (A.d.hess == PQP_HESSIAN_DIAGONAL) {
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
  if (tid == 0) A.p.c[q] = cacc;
}

} // namespace

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
  const int nd = a->d.n + a->d.ne + a->d.nc;
  size_t smem = sizeof(double) * (size_t)(nd + NW * 256 + a->d.ne + a->d.ni + 64);
  cudaError_t e = cudaFuncSetAttribute(pqp_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  pqp_setup_kernel<<<a->count, NT, smem, (cudaStream_t)stream>>>(*a);
  return (int)cudaGetLastError();
}

extern "C" int
pqp_launch_solve(const PqpSolveArgs* a, int grid, void* stream)
{
  size_t smem = sizeof(double) * (size_t)a->lay.smem_doubles + (size_t)a->lay.smem_int_bytes;
  // fast kernel: vectors and S^-1 in shared memory; P^-1 either there too or swept inside the S^-1 region
  const int64_t symn = (int64_t)a->d.n * (a->d.n + 1) / 2, symc = (int64_t)a->lay.si_cap * (a->lay.si_cap + 1) / 2;
  const bool fast = a->lay.in_smem[PA_VEC] && a->lay.in_smem[PA_MS] && (a->lay.in_smem[PA_M1] || a->d.hess != PQP_HESSIAN_DENSE || symn <= symc);
  auto kern = (a->lay.kind == 1) ? tilek::pqp_solve_kernel : (fast ? fastk::pqp_solve_kernel : genk::pqp_solve_kernel);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<grid, NT, smem, (cudaStream_t)stream>>>(*a);
  return (int)cudaGetLastError();
}
