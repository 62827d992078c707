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

namespace {

#define NT PQP_NT
#define NW PQP_NW
#define FULL 0xffffffffu

struct Ctx
{
  int n, ne, ni, nc, box, hess, cap;
  int ns; // current size of the dual block: ne + number of active inequalities
  // factor storage
  double *M1, *As, *Ms, *G, *Y;
  const double *Hs, *Cs;        // scaled matrices of this QP (global)
  const double *Hm, *Am, *Cm;   // model matrices (global, unscaled)
  // vectors
  double *x, *y, *z, *xp, *yp, *zp;
  double *dx, *ds, *dz;
  double *rx, *rs, *ex, *es;
  double *dual, *se, *rup, *si;
  double *hdx, *adx, *atdy, *cdx, *ctdz, *q;
  double *gs, *bs, *us, *ls, *is, *delta;
  double *b, *u, *l;
  double *d1inv, *dsv, *dsinv;
  double *t1, *t2, *t3, *s1, *s2, *s3, *s4;
  double *alphas, *grads, *scratch, *red;
  int *cons_slot, *slot_cons, *list1, *list2;
  unsigned char *act_up, *act_low;
  int *iscratch; // 2*NW + 8 ints
  double c_scale; // ruiz.c
};

__device__ __forceinline__ double nanmax(double a, double b)
{
  return (b > a || b != b) ? b : a;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// K sums followed by KM maxima reduced over the CTA; the result is returned to
// every thread (block-uniform control flow depends on it).
template<int KS, int KM>
__device__ void block_reduce(const Ctx& c, double* sums, double* maxs)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < KS; ++k) sums[k] = warp_sum(sums[k]);
#pragma unroll
  for (int k = 0; k < KM; ++k) maxs[k] = warp_max(maxs[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < KS; ++k) c.red[warp * (KS + KM) + k] = sums[k];
#pragma unroll
    for (int k = 0; k < KM; ++k) c.red[warp * (KS + KM) + KS + k] = maxs[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += c.red[w * (KS + KM) + k];
    sums[k] = s;
  }
#pragma unroll
  for (int k = 0; k < KM; ++k) {
    double m = c.red[KS + k];
#pragma unroll
    for (int w = 1; w < NW; ++w) m = nanmax(m, c.red[w * (KS + KM) + KS + k]);
    maxs[k] = m;
  }
  __syncthreads();
}
__device__ double block_max1(const Ctx& c, double v)
{
  double dummy[1] = { 0 };
  double m[1] = { v };
  block_reduce<0, 1>(c, dummy, m);
  return m[0];
}
__device__ double block_sum1(const Ctx& c, double v)
{
  double s[1] = { v };
  double dummy[1] = { 0 };
  block_reduce<1, 0>(c, s, dummy);
  return s[0];
}

// inclusive prefix sum over elements 0..cnt-1 (one per thread, cnt <= NT)
__device__ double block_scan_incl(const Ctx& c, double v)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) c.red[warp] = v;
  __syncthreads();
  double off = 0;
  for (int w = 0; w < warp; ++w) off += c.red[w];
  __syncthreads();
  return v + off;
}

// ordered stream compaction: list[k] = indices i in [0, count) with pred(i),
// ascending. Returns the number of entries (block-uniform).
template<class Pred>
__device__ int block_compact(const Ctx& c, int count, int* list, Pred pred)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int base = 0;
  for (int i0 = 0; i0 < count; i0 += NT) {
    int i = i0 + threadIdx.x;
    bool p = (i < count) && pred(i);
    unsigned m = __ballot_sync(FULL, p);
    if (lane == 0) c.iscratch[warp] = __popc(m);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += c.iscratch[w];
    int tot = 0;
    for (int w = 0; w < NW; ++w) tot += c.iscratch[w];
    if (p) list[off + __popc(m & ((1u << lane) - 1u))] = i;
    base += tot;
    __syncthreads();
  }
  return base;
}

__device__ __forceinline__ size_t tri_off(int i)
{
  return (size_t)i * (size_t)(i - 1) / 2;
}
__device__ __forceinline__ size_t gidx(int a, int b)
{
  int hi = a > b ? a : b, lo = a > b ? b : a;
  return (size_t)hi * (size_t)(hi + 1) / 2 + (size_t)lo;
}

// ---------------------------------------------------------------------------
// triangular mat-vecs on packed strict-lower unit-triangular M (row i has i
// entries at tri_off(i)).
//   tri_mv  : y = (I + M) x          [optionally y .*= scale]
//   tri_mv_t: y = sign * (I + M)^T x
// ---------------------------------------------------------------------------
template<int LPR>
__device__ void tri_mv_impl(const double* __restrict__ M, const double* __restrict__ x, double* __restrict__ y, int n, const double* __restrict__ scale)
{
  const int sub = threadIdx.x / LPR, sl = threadIdx.x % LPR;
  constexpr int NSUB = NT / LPR;
  // the trip count is uniform over the CTA: sub-groups of one warp own
  // different rows, and every lane must take part in the shuffles below
  for (int i0 = 0; i0 < n; i0 += NSUB) {
    const int i = i0 + sub;
    const bool valid = i < n;
    double acc = 0;
    if (valid) {
      const double* row = M + tri_off(i);
      double a0 = 0, a1 = 0;
      int j = sl;
      for (; j + LPR < i; j += 2 * LPR) {
        a0 += row[j] * x[j];
        a1 += row[j + LPR] * x[j + LPR];
      }
      if (j < i) a0 += row[j] * x[j];
      acc = a0 + a1;
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    if (valid && sl == 0) {
      double v = acc + x[i];
      y[i] = scale ? v * scale[i] : v;
    }
  }
  __syncthreads();
}
__device__ void tri_mv(const Ctx& c, const double* M, const double* x, double* y, int n, const double* scale)
{
  if (n <= 160)
    tri_mv_impl<8>(M, x, y, n, scale);
  else
    tri_mv_impl<32>(M, x, y, n, scale);
}

__device__ void tri_mv_t(const Ctx& c, const double* __restrict__ M, const double* __restrict__ x, double* __restrict__ y, int n, double sign)
{
  int ncol = (n + 31) & ~31;
  if (ncol > NT) ncol = NT;
  if (ncol < 32) ncol = 32;
  const int R = NT / ncol;
  const int jl = threadIdx.x % ncol, r = threadIdx.x / ncol;
  for (int jb = 0; jb < n; jb += ncol) {
    const int j = jb + jl;
    double a0 = 0, a1 = 0;
    if (r < R && j < n) {
      int i = j + 1 + r;
      for (; i + R < n; i += 2 * R) {
        a0 += M[tri_off(i) + j] * x[i];
        a1 += M[tri_off(i + R) + j] * x[i + R];
      }
      if (i < n) a0 += M[tri_off(i) + j] * x[i];
    }
    double acc = a0 + a1;
    if (R > 1) {
      if (r < R) c.scratch[r * ncol + jl] = acc;
      __syncthreads();
      if (r == 0 && j < n) {
        for (int rr = 1; rr < R; ++rr) acc += c.scratch[rr * ncol + jl];
      }
    }
    if (r == 0 && j < n) y[j] = sign * (acc + x[j]);
    if (R > 1 && jb + ncol < n) __syncthreads();
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// Row sources: 0 = rows of a plain matrix, 1 = dual slots (equality rows then
// active constraints), 2 = a list of constraint indices, 3 = all constraints.
// A constraint i < ni is row i of C_s; i >= ni is the box row i_s[k] e_k
// (solver.hpp:74-81, linesearch.hpp:725-731).
// ---------------------------------------------------------------------------
struct RowSrc
{
  const double* base;
  const int* list;
  int mode;
};
__device__ __forceinline__ const double* get_row(const Ctx& c, const RowSrc& rs, int r, int& bk, int& idx)
{
  bk = -1;
  idx = r;
  int cons;
  switch (rs.mode) {
    case 0:
      return rs.base + (size_t)r * c.n;
    case 1:
      if (r < c.ne) return c.As + (size_t)r * c.n;
      cons = c.slot_cons[r];
      break;
    case 2:
      cons = rs.list[r];
      idx = cons;
      break;
    default:
      cons = r;
      break;
  }
  if (cons < c.ni) return c.Cs + (size_t)cons * c.n;
  bk = cons - c.ni;
  return nullptr;
}

// out[idx(r)] = row_r . x   for r in [r0, r1)
__device__ void rows_dot(const Ctx& c, RowSrc rs, int r0, int r1, const double* __restrict__ x, double* __restrict__ out)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = c.n;
  for (int r = r0 + warp; r < r1; r += NW) {
    int bk, idx;
    const double* row = get_row(c, rs, r, bk, idx);
    double acc;
    if (row) {
      double a0 = 0, a1 = 0;
      int j = lane;
      for (; j + 32 < n; j += 64) {
        a0 += row[j] * x[j];
        a1 += row[j + 32] * x[j + 32];
      }
      if (j < n) a0 += row[j] * x[j];
      acc = warp_sum(a0 + a1);
    } else {
      acc = c.is[bk] * x[bk];
    }
    if (lane == 0) out[idx] = acc;
  }
  __syncthreads();
}

// out[j] = beta*add[j] + sign * sum_r coef[idx(r)] row_r[j]     (add may be null -> 0)
__device__ void rows_axpy_t(const Ctx& c, RowSrc rs, int r0, int r1, const double* __restrict__ coef, double* out, const double* add, double sign)
{
  const int n = c.n;
  int ncol = (n + 31) & ~31;
  if (ncol > NT) ncol = NT;
  const int R = NT / ncol;
  const int jl = threadIdx.x % ncol, rr0 = threadIdx.x / ncol;
  for (int jb = 0; jb < n; jb += ncol) {
    const int j = jb + jl;
    double acc = 0;
    if (rr0 < R && j < n) {
      for (int r = r0 + rr0; r < r1; r += R) {
        int bk, idx;
        const double* row = get_row(c, rs, r, bk, idx);
        if (row)
          acc += coef[idx] * row[j];
        else if (bk == j)
          acc += coef[idx] * c.is[bk];
      }
    }
    if (R > 1) {
      if (rr0 < R) c.scratch[rr0 * ncol + jl] = acc;
      __syncthreads();
      if (rr0 == 0 && j < n) {
        for (int k = 1; k < R; ++k) acc += c.scratch[k * ncol + jl];
      }
    }
    if (rr0 == 0 && j < n) out[j] = (add ? add[j] : 0.0) + sign * acc;
    if (R > 1 && jb + ncol < n) __syncthreads();
  }
  __syncthreads();
}

// y = P^-1 v  (P = Hs + rho I = L1 D1 L1^T, M1 = L1^-1). v may alias y.
__device__ void apply_Pinv(const Ctx& c, const double* v, double* y)
{
  if (c.hess != PQP_HESSIAN_DENSE) {
    for (int j = threadIdx.x; j < c.n; j += NT) y[j] = v[j] * c.d1inv[j];
    __syncthreads();
    return;
  }
  tri_mv(c, c.M1, v, c.t3, c.n, c.d1inv);
  tri_mv_t(c, c.M1, c.t3, y, c.n, 1.0);
}

// Bordering step: given the inverse factor of the leading i x i block
// (M, d, dinv), append row i of the symmetric matrix whose off-diagonal
// entries are a[0..i) and diagonal a_diag. Replaces the reference's
// factorize / insert_block_at for an appended row (ldlt.hpp:431-475).
__device__ void append_row(const Ctx& c, double* M, double* dv, double* dinv, const double* a, double a_diag, int i, double* u /*tmp i*/, double* l /*tmp i*/)
{
  double dnew = a_diag;
  if (i > 0) {
    tri_mv(c, M, a, u, i, nullptr); // u = L^-1 a
    double part = 0;
    for (int j = threadIdx.x; j < i; j += NT) {
      double lj = u[j] * dinv[j];
      l[j] = lj;
      part += u[j] * lj;
    }
    dnew -= block_sum1(c, part); // (also orders the writes to l)
    tri_mv_t(c, M, l, M + tri_off(i), i, -1.0); // new row of the inverse factor
  }
  if (threadIdx.x == 0) {
    if (dv) dv[i] = dnew;
    dinv[i] = 1.0 / dnew;
  }
  __syncthreads();
}

__device__ __forceinline__ int row_id(const Ctx& c, int s)
{
  return s < c.ne ? s : c.ne + c.slot_cons[s];
}

// Solve K [ox; os] = [b1; b2] with the current factors. In-place allowed
// (ox == b1, os == b2). Replaces Ldlt::solve_in_place (ldlt.hpp:767-782).
__device__ void solve_kkt(const Ctx& c, const double* b1, const double* b2, double* ox, double* os)
{
  const int ns = c.ns;
  if (ns == 0) {
    apply_Pinv(c, b1, ox);
    return;
  }
  apply_Pinv(c, b1, c.t1);
  RowSrc slots{ nullptr, nullptr, 1 };
  rows_dot(c, slots, 0, ns, c.t1, c.s1);
  for (int s = threadIdx.x; s < ns; s += NT) c.s1[s] -= b2[s];
  __syncthreads();
  tri_mv(c, c.Ms, c.s1, c.s2, ns, c.dsinv);
  tri_mv_t(c, c.Ms, c.s2, os, ns, 1.0);
  rows_axpy_t(c, slots, 0, ns, os, c.t2, b1, -1.0);
  apply_Pinv(c, c.t2, ox);
}

// Append dual slot `s == c.ns` (row already registered in slot_cons) with
// proximal parameter mu: y = P^-1 b_s, g = B y, bordering of Ms.
__device__ void insert_slot(Ctx& c, double mu)
{
  const int s = c.ns;
  RowSrc slots{ nullptr, nullptr, 1 };
  int bk, idx;
  const double* row = get_row(c, slots, s, bk, idx);
  if (!row) {
    for (int j = threadIdx.x; j < c.n; j += NT) c.t2[j] = (j == bk) ? c.is[bk] : 0.0;
    __syncthreads();
    row = c.t2;
  }
  apply_Pinv(c, row, c.t1);
  rows_dot(c, slots, 0, s + 1, c.t1, c.s3);
  const int ids = row_id(c, s);
  for (int j = threadIdx.x; j <= s; j += NT) c.G[gidx(ids, row_id(c, j))] = c.s3[j];
  append_row(c, c.Ms, c.dsv, c.dsinv, c.s3, c.s3[s] + mu, s, c.s1, c.s2);
  if (threadIdx.x == 0) c.ns = s + 1;
  __syncthreads();
}

// Remove dual slot k (k >= ne): rank-one modification of the trailing rows of
// Ms + compaction. Replaces Ldlt::delete_at (ldlt.hpp:340-387, modify.hpp:82-127).
__device__ void delete_slot(Ctx& c, int k)
{
  const int ns = c.ns;
  const int t = ns - 1 - k;
  double* p = c.s1;
  double* beta = c.s2;
  if (t > 0) {
    // gamma recurrence 1/alpha_{j+1} = 1/alpha_j + p_j^2 / d_j as a prefix sum
    for (int i0 = 0; i0 < t; i0 += NT) {
      int i = i0 + threadIdx.x;
      double pi = 0, di = 1, e = 0;
      if (i < t) {
        pi = -c.Ms[tri_off(k + 1 + i) + k];
        di = c.dsv[k + 1 + i];
        e = pi * pi / di;
      }
      double incl = block_scan_incl(c, e);
      double carry = (i0 == 0) ? 1.0 / c.dsv[k] : c.s4[0];
      double g1 = carry + incl;
      double g0 = g1 - e;
      if (i < t) {
        p[i] = pi;
        beta[i] = pi / (di * g1);
        double dn = di * g1 / g0;
        c.s3[i] = dn;
      }
      __syncthreads();
      if (i == min(i0 + NT, t) - 1) c.s4[0] = g1;
      __syncthreads();
    }
    // apply Ltilde^-1 to every column of the trailing rows
    for (int col = threadIdx.x; col < ns; col += NT) {
      if (col == k) continue;
      const double mk = (col < k) ? c.Ms[tri_off(k) + col] : 0.0;
      double s = 0;
      int i = (col > k + 1) ? (col - (k + 1)) : 0;
      for (; i < t; ++i) {
        const int row = k + 1 + i;
        double v = (col < row) ? c.Ms[tri_off(row) + col] : 1.0;
        v += p[i] * mk;
        const double yv = v - p[i] * s;
        s += beta[i] * yv;
        if (col < row) c.Ms[tri_off(row) + col] = yv;
      }
    }
    __syncthreads();
    // compaction of Ms: drop row k and column k (order preserving, in place)
    const size_t e0 = tri_off(k + 1), e1 = tri_off(ns);
    for (size_t eb = e0; eb < e1; eb += NT) {
      size_t e = eb + threadIdx.x;
      double v = 0;
      int row = 0, col = 0;
      bool valid = e < e1;
      if (valid) {
        row = (int)((1.0 + sqrt(1.0 + 8.0 * (double)e)) * 0.5);
        while (tri_off(row) > e) --row;
        while (tri_off(row + 1) <= e) ++row;
        col = (int)(e - tri_off(row));
        v = c.Ms[e];
      }
      __syncthreads();
      if (valid && col != k) c.Ms[tri_off(row - 1) + col - (col > k ? 1 : 0)] = v;
      __syncthreads();
    }
  }
  // shift d, dinv, slot_cons; fix cons_slot
  {
    const int cons_k = c.slot_cons[k];
    double dn = 0;
    int sc = 0;
    int i = k + threadIdx.x;
    // (t <= cap; loop in chunks to stay generic)
    for (int i0 = k; i0 < ns - 1; i0 += NT) {
      i = i0 + threadIdx.x;
      bool valid = i < ns - 1;
      if (valid) {
        dn = c.s3[i - k];
        sc = c.slot_cons[i + 1];
      }
      __syncthreads();
      if (valid) {
        c.dsv[i] = dn;
        c.dsinv[i] = 1.0 / dn;
        c.slot_cons[i] = sc;
        c.cons_slot[sc] = i;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      c.cons_slot[cons_k] = -1;
      c.ns = ns - 1;
    }
    __syncthreads();
  }
}

// Rebuild Ms from the cached Gram matrix with the given proximal parameters.
// Replaces Ldlt::diagonal_update_clobber_indices (ldlt.hpp:516-570) used by
// mu_update (solver.hpp:130-169).
__device__ void rebuild_Ms_from_G(Ctx& c, double mu_eq, double mu_in)
{
  const int ns = c.ns;
  for (int s = 0; s < ns; ++s) {
    const int ids = row_id(c, s);
    for (int j = threadIdx.x; j <= s; j += NT) c.s3[j] = c.G[gidx(ids, row_id(c, j))];
    __syncthreads();
    append_row(c, c.Ms, c.dsv, c.dsinv, c.s3, c.s3[s] + (s < c.ne ? mu_eq : mu_in), s, c.s1, c.s2);
  }
}

// Factorise P = Hs + rho I into (M1, d1inv) by successive bordering.
// Replaces the x-block part of Ldlt::factorize (ldlt.hpp:718-744).
__device__ void build_M1(Ctx& c, double rho)
{
  const int n = c.n;
  if (c.hess != PQP_HESSIAN_DENSE) {
    for (int j = threadIdx.x; j < n; j += NT) {
      double h = (c.hess == PQP_HESSIAN_DIAGONAL) ? c.Hs[(size_t)j * n + j] : 0.0;
      c.d1inv[j] = 1.0 / (h + rho);
    }
    __syncthreads();
    return;
  }
  for (int i = 0; i < n; ++i) {
    for (int j = threadIdx.x; j <= i; j += NT) c.t1[j] = c.Hs[(size_t)i * n + j];
    __syncthreads();
    append_row(c, c.M1, nullptr, c.d1inv, c.t1, c.t1[i] + rho, i, c.t2, c.t3);
  }
}

// (Re)build the dual block for the slots currently registered (0..ns_target):
// used for the first factorisation (equality rows only, helpers.hpp:241-285)
// and by refactorize (solver.hpp:40-87).
__device__ void build_dual_block(Ctx& c, int ns_target, double mu_eq, double mu_in)
{
  if (threadIdx.x == 0) c.ns = 0;
  __syncthreads();
  for (int s = 0; s < ns_target; ++s) insert_slot(c, s < c.ne ? mu_eq : mu_in);
}

struct Scal
{
  double rho, mu_eq, mu_in, mu_eq_inv, mu_in_inv, nu;
  long long iter, iter_ext, mu_updates;
  int status;
  double iterative_residual;
  bool factor_fresh; // !constraints_changed (solver.hpp:48)
};

// err = rhs - K dw, with the by-products the Newton loop reuses
// (solver.hpp:245-318; quirk 3 of SURVEY Appendix A). Returns |err|_inf.
__device__ double kkt_residual(const Ctx& c, const Scal& sc)
{
  const int n = c.n, ne = c.ne, ns = c.ns, ni = c.ni;
  RowSrc slots{ nullptr, nullptr, 1 };
  if (c.hess == PQP_HESSIAN_DENSE) {
    rows_dot(c, RowSrc{ c.Hs, nullptr, 0 }, 0, n, c.dx, c.hdx);
  } else {
    for (int j = threadIdx.x; j < n; j += NT) c.hdx[j] = (c.hess == PQP_HESSIAN_DIAGONAL) ? c.Hs[(size_t)j * n + j] * c.dx[j] : 0.0;
  }
  rows_dot(c, RowSrc{ c.As, nullptr, 0 }, 0, ne, c.dx, c.adx);
  rows_axpy_t(c, RowSrc{ c.As, nullptr, 0 }, 0, ne, c.ds, c.atdy, nullptr, 1.0);
  rows_dot(c, RowSrc{ nullptr, nullptr, 3 }, 0, c.nc, c.dx, c.cdx);
  (void)ni;
  rows_axpy_t(c, slots, ne, ns, c.ds, c.ctdz, nullptr, 1.0); // sum over active constraints dz_i c_i
  double m = 0;
  for (int j = threadIdx.x; j < n; j += NT) {
    double e = c.rx[j] - (c.hdx[j] + sc.rho * c.dx[j] + c.atdy[j] + c.ctdz[j]);
    c.ex[j] = e;
    m = nanmax(m, fabs(e));
  }
  for (int s = threadIdx.x; s < ns; s += NT) {
    double e;
    if (s < ne)
      e = c.rs[s] - (c.adx[s] - sc.mu_eq * c.ds[s]);
    else
      e = c.rs[s] - (c.cdx[c.slot_cons[s]] - sc.mu_in * c.ds[s]);
    c.es[s] = e;
    m = nanmax(m, fabs(e));
  }
  return block_max1(c, m);
}

// solver.hpp:40-87: rebuild everything from scratch (same active set)
__device__ void refactorize(Ctx& c, Scal& sc)
{
  if (sc.factor_fresh) return;
  const int ns_target = c.ns;
  __syncthreads();
  build_M1(c, sc.rho);
  build_dual_block(c, ns_target, sc.mu_eq, sc.mu_in);
  sc.factor_fresh = true;
}

// solver.hpp:408-541
__device__ void iterative_solve(Ctx& c, Scal& sc, const pqp_settings& S, double eps)
{
  for (int pass = 0; pass < 2; ++pass) {
    int it = 0, it_stab = 0;
    solve_kkt(c, c.rx, c.rs, c.dx, c.ds);
    double err = kkt_residual(c, sc);
    ++it;
    double prev = err;
    while (err >= eps) {
      if (it >= S.nb_iterative_refinement) break;
      ++it;
      solve_kkt(c, c.ex, c.es, c.ex, c.es);
      for (int j = threadIdx.x; j < c.n; j += NT) c.dx[j] += c.ex[j];
      for (int s = threadIdx.x; s < c.ns; s += NT) c.ds[s] += c.es[s];
      __syncthreads();
      err = kkt_residual(c, sc);
      if (err > prev)
        it_stab += 1;
      else
        it_stab = 0;
      if (it_stab == 2) break;
      prev = err;
    }
    sc.iterative_residual = err;
    if (pass == 0 && err >= fmax(eps, S.eps_refact) && !sc.factor_fresh) {
      refactorize(c, sc);
      continue;
    }
    break;
  }
  for (int j = threadIdx.x; j < c.n; j += NT) c.rx[j] = 0;
  for (int s = threadIdx.x; s < c.cap; s += NT) c.rs[s] = 0;
  __syncthreads();
}

// linesearch.hpp:551-786 with act[i] = act_up | act_low
__device__ void active_set_change(Ctx& c, Scal& sc)
{
  // deletions, from the last slot to the first
  int ndel = block_compact(c, c.ns - c.ne, c.list1, [&](int k) {
    int cons = c.slot_cons[c.ne + k];
    return !(c.act_up[cons] || c.act_low[cons]);
  });
  for (int k = ndel - 1; k >= 0; --k) delete_slot(c, c.ne + c.list1[k]);
  int nadd = block_compact(c, c.nc, c.list1, [&](int i) { return (c.act_up[i] || c.act_low[i]) && c.cons_slot[i] < 0; });
  for (int k = 0; k < nadd; ++k) {
    if (threadIdx.x == 0) {
      int cons = c.list1[k];
      c.slot_cons[c.ns] = cons;
      c.cons_slot[cons] = c.ns;
    }
    __syncthreads();
    insert_slot(c, sc.mu_in);
  }
  if (ndel > 0 || nadd > 0) sc.factor_fresh = false;
}

// unscaled global residual pieces -------------------------------------------------
struct Glob
{
  double pri_lhs, pri_eq_rhs0, pri_in_rhs0, pri_eq_lhs, pri_in_lhs;
  double dua_lhs, dua_rhs0, dua_rhs1, dua_rhs3, gap, rhs_gap;
};

// utils.hpp:166-252
__device__ void global_primal_residual(Ctx& c, const Scal& sc, const pqp_settings& S, Glob& g)
{
  const int n = c.n, ne = c.ne, ni = c.ni, nc = c.nc;
  rows_dot(c, RowSrc{ c.As, nullptr, 0 }, 0, ne, c.x, c.se);
  rows_dot(c, RowSrc{ nullptr, nullptr, 3 }, 0, ni, c.x, c.rup);
  double mx[5] = { 0, 0, 0, 0, 0 }; // eq_rhs0, in_rhs0, eq_lhs, in_lhs, |x| stuff
  double dummy[1] = { 0 };
  const double* de = c.delta + n;
  const double* di = c.delta + n + ne;
  const double* db = c.delta + n + ne + ni;
  for (int i = threadIdx.x; i < ne; i += NT) {
    double v = c.se[i] / de[i];
    mx[0] = nanmax(mx[0], fabs(v));
    v -= c.b[i];
    mx[2] = nanmax(mx[2], fabs(v));
    c.se[i] = v; // unscaled Ax - b, rescaled below
  }
  for (int i = threadIdx.x; i < nc; i += NT) {
    double v;
    if (i < ni) {
      v = c.rup[i] / di[i];
      mx[1] = nanmax(mx[1], fabs(v));
    } else {
      v = c.x[i - ni] * c.delta[i - ni]; // unscale_primal
    }
    c.rup[i] = v;
    double sv = fmax(v - c.u[i], 0.0) + fmin(v - c.l[i], 0.0);
    c.si[i] = sv;
    mx[3] = nanmax(mx[3], fabs(sv));
    if (i >= ni) {
      // quirk kept: active_part_z.tail = x(scaled) - si ; rhs_0 also takes |x| (scaled), utils.hpp:225-231
      mx[1] = nanmax(mx[1], fabs(c.x[i - ni] - sv));
      mx[1] = nanmax(mx[1], fabs(c.x[i - ni]));
    }
  }
  (void)db;
  block_reduce<0, 4>(c, dummy, mx);
  g.pri_eq_rhs0 = mx[0];
  g.pri_in_rhs0 = mx[1];
  g.pri_eq_lhs = mx[2];
  g.pri_in_lhs = mx[3];
  g.pri_lhs = fmax(mx[2], mx[3]);
  if (S.primal_infeasibility_solving && sc.status == PQP_PRIMAL_INFEASIBLE) {
    rows_axpy_t(c, RowSrc{ c.Am, nullptr, 0 }, 0, ne, c.se, c.t1, nullptr, 1.0);
    rows_axpy_t(c, RowSrc{ c.Cm, nullptr, 0 }, 0, ni, c.si, c.t1, c.t1, 1.0);
    double m = 0;
    for (int j = threadIdx.x; j < n; j += NT) m = nanmax(m, fabs(c.t1[j]));
    g.pri_lhs = block_max1(c, m);
  }
  for (int i = threadIdx.x; i < ne; i += NT) c.se[i] *= de[i];
  __syncthreads();
}

// utils.hpp:439-587
__device__ void global_dual_residual(Ctx& c, const Scal& sc, Glob& g)
{
  const int n = c.n, ne = c.ne, ni = c.ni, nc = c.nc;
  const double cs = c.c_scale;
  // Hx
  if (c.hess == PQP_HESSIAN_DENSE) {
    rows_dot(c, RowSrc{ c.Hs, nullptr, 0 }, 0, n, c.x, c.t1);
  } else {
    for (int j = threadIdx.x; j < n; j += NT) c.t1[j] = (c.hess == PQP_HESSIAN_DIAGONAL) ? c.Hs[(size_t)j * n + j] * c.x[j] : 0.0;
    __syncthreads();
  }
  rows_axpy_t(c, RowSrc{ c.As, nullptr, 0 }, 0, ne, c.y, c.t2, nullptr, 1.0);           // A^T y
  rows_axpy_t(c, RowSrc{ c.Cs, nullptr, 0 }, 0, ni, c.z, c.t3, nullptr, 1.0);           // C^T z_C
  double sm[6] = { 0, 0, 0, 0, 0, 0 }; // g.x, xHx, b.y, zu, zl, (unused)
  double mx[4] = { 0, 0, 0, 0 };       // rhs0, rhs1, rhs3, lhs
  const double inf_b = 1.3407807929942596e+154; // sqrt(DBL_MAX), helpers/common.hpp:20-24
  for (int j = threadIdx.x; j < n; j += NT) {
    const double dxc = c.delta[j] * cs;
    double hx = c.t1[j], aty = c.t2[j], ctz = c.t3[j];
    double zb = c.box ? c.z[ni + j] * c.is[j] : 0.0;
    double dr = c.gs[j] + hx + aty + ctz + zb;
    c.dual[j] = dr;
    const double hxu = hx / dxc;
    mx[0] = nanmax(mx[0], fabs(hxu));
    mx[1] = nanmax(mx[1], fabs(aty / dxc));
    mx[2] = nanmax(mx[2], fabs(ctz / dxc));
    if (c.box) mx[2] = nanmax(mx[2], fabs(zb / dxc));
    mx[3] = nanmax(mx[3], fabs(dr / dxc));
    const double xu = c.x[j] * c.delta[j];
    sm[0] += (c.gs[j] / dxc) * xu; // model.g = gs / (delta c)
    sm[1] += hxu * xu;
  }
  const double* de = c.delta + n;
  const double* di = c.delta + n + ne;
  for (int i = threadIdx.x; i < ne; i += NT) sm[2] += c.b[i] * (c.y[i] * de[i] / cs);
  for (int i = threadIdx.x; i < nc; i += NT) {
    double zu_ = c.z[i] * di[i] / cs; // delta laid out [x | eq | in | box]: di[i] covers box too
    if (c.act_up[i]) sm[3] += zu_ * fmin(c.u[i], inf_b);
    if (c.act_low[i]) sm[4] += zu_ * fmax(c.l[i], -inf_b);
  }
  block_reduce<5, 4>(c, sm, mx);
  g.dua_rhs0 = (c.hess == PQP_HESSIAN_ZERO) ? 0.0 : mx[0];
  g.dua_rhs1 = mx[1];
  g.dua_rhs3 = mx[2];
  g.dua_lhs = mx[3];
  double gap = sm[0];
  double rhs_gap = fabs(gap);
  if (c.hess != PQP_HESSIAN_ZERO) {
    gap += sm[1];
    rhs_gap = fmax(rhs_gap, fabs(sm[1]));
  }
  rhs_gap = fmax(rhs_gap, fabs(sm[2]));
  gap += sm[2];
  rhs_gap = fmax(rhs_gap, fabs(sm[3]));
  gap += sm[3];
  rhs_gap = fmax(rhs_gap, fabs(sm[4]));
  gap += sm[4];
  g.gap = gap;
  g.rhs_gap = rhs_gap;
  (void)sc;
}

// coefficients of phi'(alpha) = a alpha + b that do not depend on alpha
// (linesearch.hpp:85-119, 133-134, 159-160 for GPDAL; :213-255, 288-304 for PDAL)
struct LsBase
{
  double a0, b0;
};

__device__ LsBase ls_base(const Ctx& c, const Scal& sc, const pqp_settings& S)
{
  const int n = c.n, ne = c.ne, nc = c.nc;
  const bool gpdal = S.merit_function_type == PQP_MERIT_GPDAL;
  double sm[10] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
  double dummy[1] = { 0 };
  for (int j = threadIdx.x; j < n; j += NT) {
    double dxj = c.dx[j];
    sm[0] += dxj * c.hdx[j];
    sm[1] += dxj * dxj;
    sm[2] += c.x[j] * c.hdx[j];
    sm[3] += (sc.rho * (c.x[j] - c.xp[j]) + c.gs[j]) * dxj;
  }
  for (int i = threadIdx.x; i < ne; i += NT) {
    double ad = c.adx[i];
    double e = ad - c.ds[i] * sc.mu_eq;
    sm[4] += ad * ad;
    sm[5] += e * e;
    sm[6] += ad * (c.se[i] + c.y[i] * sc.mu_eq);
    sm[7] += e * c.se[i];
  }
  if (gpdal) {
    for (int i = threadIdx.x; i < nc; i += NT) {
      sm[8] += c.dz[i] * c.dz[i];
      sm[9] += c.dz[i] * c.z[i];
    }
  }
  block_reduce<10, 0>(c, sm, dummy);
  LsBase r;
  const double nu = gpdal ? 1.0 : sc.nu;
  r.a0 = sm[0] + sc.mu_eq_inv * sm[4] + sc.rho * sm[1] + sm[5] * sc.mu_eq_inv * nu;
  r.b0 = sm[2] + sm[3] + sc.mu_eq_inv * sm[6] + nu * sc.mu_eq_inv * sm[7];
  if (gpdal) {
    r.a0 += sc.mu_in * (1.0 - S.alpha_gpdal) * sm[8];
    r.b0 += sc.mu_in * (1.0 - S.alpha_gpdal) * sm[9];
  }
  return r;
}

// alpha-dependent part, evaluated by ONE thread over all constraints
// (linesearch.hpp:121-152 / 257-304)
__device__ __forceinline__ void ls_eval(const Ctx& c, const Scal& sc, const pqp_settings& S, const LsBase& base, double alpha, double& a, double& b)
{
  const bool gpdal = S.merit_function_type == PQP_MERIT_GPDAL;
  double sq = 0, dt = 0, sq2 = 0, dt2 = 0;
  for (int i = 0; i < c.nc; ++i) {
    const double cd = c.cdx[i], ru = c.rup[i], sl = c.si[i];
    const bool up = (ru + cd * alpha) > 0.0;
    const bool low = (sl + cd * alpha) < 0.0;
    const double cact = (up || low) ? cd : 0.0;
    const double apz = (up ? ru : 0.0) + (low ? sl : 0.0);
    sq += cact * cact;
    dt += apz * cact;
    if (!gpdal) {
      const double e = cact - c.dz[i] * sc.mu_in;
      const double f = apz - c.z[i] * sc.mu_in;
      sq2 += e * e;
      dt2 += e * f;
    }
  }
  if (gpdal) {
    a = base.a0 + sc.mu_in_inv * sq / S.alpha_gpdal;
    b = base.b0 + sc.mu_in_inv * dt / S.alpha_gpdal;
  } else {
    a = base.a0 + sc.mu_in_inv * sq + sc.nu * sc.mu_in_inv * sq2;
    b = base.b0 + sc.mu_in_inv * dt + sc.nu * sc.mu_in_inv * dt2;
  }
}

// Exact line search, linesearch.hpp:322-538. Breakpoints are evaluated in
// parallel (one thread each); phi' is non-decreasing, so "first breakpoint
// with phi' >= 0" / "last with phi' < 0" are a min / max reduction instead of
// the reference's sort + sequential scan.
__device__ double primal_dual_ls(Ctx& c, const Scal& sc, const pqp_settings& S)
{
  const double eps = 2.220446049250313e-16;
  const int nc = c.nc;
  LsBase base = ls_base(c, sc, S);
  if (threadIdx.x == 0) c.iscratch[2 * NW] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < nc; i += NT) {
    const double cd = c.cdx[i];
    if (cd != 0.0) {
      double a1 = -c.rup[i] / (cd + eps);
      if (a1 > eps) c.alphas[atomicAdd(&c.iscratch[2 * NW], 1)] = a1;
      double a2 = -c.si[i] / (cd + eps);
      if (a2 > eps) c.alphas[atomicAdd(&c.iscratch[2 * NW], 1)] = a2;
    }
  }
  __syncthreads();
  const int n_alpha = c.iscratch[2 * NW];
  // thread 0 of the last warp additionally evaluates alpha = 0
  double best_pos_alpha = INFINITY, best_pos_grad = 0, best_neg_alpha = 0, best_neg_grad = 0;
  for (int k = threadIdx.x; k < n_alpha + 1; k += NT) {
    const double al = (k < n_alpha) ? c.alphas[k] : 0.0;
    double a, b;
    ls_eval(c, sc, S, base, al, a, b);
    const double gr = a * al + b;
    if (k == n_alpha) {
      c.grads[0] = a;
      c.grads[1] = b; // phi'(0) pieces
    } else if (gr < 0.0) {
      if (al > best_neg_alpha) {
        best_neg_alpha = al;
        best_neg_grad = gr;
      }
    } else if (al < best_pos_alpha) {
      best_pos_alpha = al;
      best_pos_grad = gr;
    }
  }
  // reduce (alpha, grad) pairs: min over positives, max over negatives
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double pa = __shfl_xor_sync(FULL, best_pos_alpha, o), pg = __shfl_xor_sync(FULL, best_pos_grad, o);
    if (pa < best_pos_alpha) {
      best_pos_alpha = pa;
      best_pos_grad = pg;
    }
    double na = __shfl_xor_sync(FULL, best_neg_alpha, o), ng = __shfl_xor_sync(FULL, best_neg_grad, o);
    if (na > best_neg_alpha) {
      best_neg_alpha = na;
      best_neg_grad = ng;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    c.red[warp * 4 + 0] = best_pos_alpha;
    c.red[warp * 4 + 1] = best_pos_grad;
    c.red[warp * 4 + 2] = best_neg_alpha;
    c.red[warp * 4 + 3] = best_neg_grad;
  }
  __syncthreads();
  double alpha_first_pos = INFINITY, first_pos_grad = 0, alpha_last_neg = 0, last_neg_grad = 0;
  for (int w = 0; w < NW; ++w) {
    if (c.red[w * 4 + 0] < alpha_first_pos) {
      alpha_first_pos = c.red[w * 4 + 0];
      first_pos_grad = c.red[w * 4 + 1];
    }
    if (c.red[w * 4 + 2] > alpha_last_neg) {
      alpha_last_neg = c.red[w * 4 + 2];
      last_neg_grad = c.red[w * 4 + 3];
    }
  }
  const double a0 = c.grads[0], b0 = c.grads[1];
  __syncthreads();
  if (n_alpha == 0) return -b0 / a0;
  // the reference stops its scan at the first non-negative gradient, so
  // negatives beyond it are never seen (linesearch.hpp:460-467)
  if (alpha_last_neg > alpha_first_pos) {
    // not monotone to rounding: fall back to the breakpoint just below
    alpha_last_neg = 0;
  }
  if (alpha_last_neg == 0.0) last_neg_grad = b0; // phi'(0) = a*0 + b
  if (alpha_first_pos == INFINITY) {
    double a, b;
    ls_eval(c, sc, S, base, 2 * alpha_last_neg + 1, a, b);
    return -b / a;
  }
  return fabs(alpha_last_neg - last_neg_grad * (alpha_first_pos - alpha_last_neg) / (first_pos_grad - last_neg_grad));
}

__device__ __forceinline__ unsigned long long gtimer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ void dbg_write(const PqpSolveArgs& A, int q, int& pos, double a, double b, double c0, double d, double e, double f)
{
  if (A.dbg && q == A.dbg_qp && threadIdx.x == 0 && pos + 6 <= A.dbg_cap) {
    A.dbg[pos + 0] = a;
    A.dbg[pos + 1] = b;
    A.dbg[pos + 2] = c0;
    A.dbg[pos + 3] = d;
    A.dbg[pos + 4] = e;
    A.dbg[pos + 5] = f;
    pos += 6;
  }
}

// ---------------------------------------------------------------------------
// one QP, start to finish: dense/solver.hpp:1088-1843
// ---------------------------------------------------------------------------
__device__ void solve_one(Ctx& c, const PqpSolveArgs& A, int q)
{
  const PqpQpParams& prm = A.p.params[q];
  const pqp_settings& S = prm.s;
  const int n = c.n, ne = c.ne, ni = c.ni, nc = c.nc;
  const int tid = threadIdx.x;
  int dbg_pos = 0;

  // ---- stage the per-QP data in shared memory --------------------------------
  {
    const PqpBatchPtrs& P = A.p;
    const double* Asg = P.As + (size_t)q * ne * n;
    if (tid == 0) {
      c.Hs = P.Hs + (size_t)q * n * n;
      c.Cs = P.Cs + (size_t)q * ni * n;
      c.Hm = P.H + (size_t)q * n * n;
      c.Am = P.A + (size_t)q * ne * n;
      c.Cm = P.C + (size_t)q * ni * n;
      if (!A.lay.in_smem[PA_AS]) c.As = const_cast<double*>(Asg);
    }
    __syncthreads();
    if (A.lay.in_smem[PA_AS]) {
      for (int i = tid; i < ne * n; i += NT) c.As[i] = Asg[i];
    }
    for (int j = tid; j < n; j += NT) c.gs[j] = P.gs[(size_t)q * n + j];
    for (int j = tid; j < ne; j += NT) {
      c.bs[j] = P.bs[(size_t)q * ne + j];
      c.b[j] = P.b[(size_t)q * ne + j];
    }
    for (int j = tid; j < nc; j += NT) {
      c.us[j] = P.us[(size_t)q * nc + j];
      c.ls[j] = P.ls[(size_t)q * nc + j];
      if (j < ni) {
        c.u[j] = P.u[(size_t)q * ni + j];
        c.l[j] = P.l[(size_t)q * ni + j];
      } else {
        c.u[j] = P.u_box[(size_t)q * n + j - ni];
        c.l[j] = P.l_box[(size_t)q * n + j - ni];
      }
      c.cons_slot[j] = -1;
      c.act_up[j] = 0;
      c.act_low[j] = 0;
    }
    if (c.box) {
      for (int j = tid; j < n; j += NT) c.is[j] = P.is[(size_t)q * n + j];
    }
    for (int j = tid; j < n + ne + nc; j += NT) c.delta[j] = P.delta[(size_t)q * (n + ne + nc) + j];
    if (tid == 0) {
      c.c_scale = P.c[q];
      c.ns = 0;
    }
    __syncthreads();
  }
  const double cs = c.c_scale;
  const double* dlx = c.delta;
  const double* dle = c.delta + n;
  const double* dli = c.delta + n + ne; // covers box entries too ([in | box] contiguous)

  Scal sc;
  sc.rho = prm.rho;
  sc.mu_eq = prm.mu_eq;
  sc.mu_in = prm.mu_in;
  sc.mu_eq_inv = 1.0 / sc.mu_eq;
  sc.mu_in_inv = 1.0 / sc.mu_in;
  sc.nu = 1.0;
  sc.iter = 0;
  sc.iter_ext = 0;
  sc.mu_updates = 0;
  sc.status = PQP_MAX_ITER_REACHED;
  sc.iterative_residual = 0;
  sc.factor_fresh = true;

  // ---- initial iterate (solver.hpp:1125-1377) --------------------------------
  if (prm.start_mode == PQP_START_WARM || prm.start_mode == PQP_START_WARM_KEEP) {
    for (int j = tid; j < n; j += NT) c.x[j] = A.p.x[(size_t)q * n + j] / dlx[j];
    for (int j = tid; j < ne; j += NT) c.y[j] = A.p.y[(size_t)q * ne + j] / dle[j] * cs;
    for (int j = tid; j < nc; j += NT) c.z[j] = A.p.z[(size_t)q * nc + j] / dli[j] * cs;
  } else {
    for (int j = tid; j < n; j += NT) c.x[j] = 0;
    for (int j = tid; j < ne; j += NT) c.y[j] = 0;
    for (int j = tid; j < nc; j += NT) c.z[j] = 0;
  }
  for (int j = tid; j < n; j += NT) {
    c.rx[j] = 0;
    c.dx[j] = 0;
  }
  for (int j = tid; j < c.cap; j += NT) {
    c.rs[j] = 0;
    c.ds[j] = 0;
  }
  for (int j = tid; j < ne; j += NT) c.se[j] = 0;
  for (int j = tid; j < nc; j += NT) {
    c.si[j] = 0;
    c.dz[j] = 0;
  }
  __syncthreads();

  // ---- first factorisation (helpers.hpp:241-285) -----------------------------
  build_M1(c, sc.rho);
  build_dual_block(c, ne, sc.mu_eq, sc.mu_in);

  if (prm.start_mode == PQP_START_EQ_GUESS) {
    // helpers.hpp:201-228
    for (int j = tid; j < n; j += NT) c.rx[j] = -c.gs[j];
    for (int j = tid; j < ne; j += NT) c.rs[j] = c.bs[j];
    __syncthreads();
    iterative_solve(c, sc, S, 1.0);
    for (int j = tid; j < n; j += NT) {
      c.x[j] = c.dx[j];
      c.dx[j] = 0;
    }
    for (int j = tid; j < ne; j += NT) {
      c.y[j] = c.ds[j];
      c.ds[j] = 0;
    }
    __syncthreads();
  } else if (prm.start_mode == PQP_START_WARM || prm.start_mode == PQP_START_WARM_KEEP) {
    // active set := { i : z_i != 0 } (solver.hpp:1300-1309)
    for (int i = tid; i < nc; i += NT) {
      c.act_up[i] = (c.z[i] != 0.0);
      c.act_low[i] = 0;
    }
    __syncthreads();
    active_set_change(c, sc);
    for (int i = tid; i < nc; i += NT) c.act_up[i] = 0;
    __syncthreads();
  }

  double bcl_eta_ext_init = pow(0.1, S.alpha_bcl);
  double bcl_eta_ext = bcl_eta_ext_init;
  double bcl_eta_in = 1.0;
  const double eps_in_min = fmin(S.eps_abs, 1e-9);
  double scaled_eps = S.eps_abs;
  Glob g;
  g.pri_lhs = g.pri_eq_rhs0 = g.pri_in_rhs0 = g.pri_eq_lhs = g.pri_in_lhs = 0;
  g.dua_lhs = g.dua_rhs0 = g.dua_rhs1 = g.dua_rhs3 = g.gap = g.rhs_gap = 0;
  const double dual_rhs2 = [&]() {
    double m = 0;
    for (int j = tid; j < n; j += NT) m = nanmax(m, fabs(c.gs[j] / (dlx[j] * cs)));
    return block_max1(c, m);
  }(); // |model.g|_inf (helpers.hpp:651)
  double info_pri = 0, info_dua = 0, info_gap = 0;
  bool infeasible_exit = false;
  bool expired = false; // watchdog (debug aid, off by default)
  const unsigned long long t_start = A.watchdog_ns ? gtimer_ns() : 0ull;

  for (long long iter = 0; iter < S.max_iter; ++iter) {
    global_primal_residual(c, sc, S, g);
    global_dual_residual(c, sc, g);
    double primal_feasibility_lhs = g.pri_lhs;
    double dual_feasibility_lhs = g.dua_lhs;
    info_pri = g.pri_lhs;
    info_dua = g.dua_lhs;
    info_gap = g.gap;
    dbg_write(A, q, dbg_pos, (double)iter, g.pri_lhs, g.dua_lhs, sc.mu_in, (double)(c.ns - ne), (double)sc.iter);

    double new_mu_in = sc.mu_in, new_mu_eq = sc.mu_eq, new_mu_in_inv = sc.mu_in_inv, new_mu_eq_inv = sc.mu_eq_inv;
    double rhs_pri = scaled_eps;
    if (S.eps_rel != 0) rhs_pri += S.eps_rel * fmax(g.pri_eq_rhs0, g.pri_in_rhs0);
    bool is_primal_feasible = primal_feasibility_lhs <= rhs_pri;
    double rhs_dua = S.eps_abs;
    if (S.eps_rel != 0) rhs_dua += S.eps_rel * fmax(fmax(g.dua_rhs3, g.dua_rhs0), fmax(g.dua_rhs1, dual_rhs2));
    bool is_dual_feasible = dual_feasibility_lhs <= rhs_dua;
    if (is_primal_feasible && is_dual_feasible) {
      if (S.check_duality_gap) {
        if (fabs(g.gap) <= S.eps_duality_gap_abs + S.eps_duality_gap_rel * g.rhs_gap) {
          sc.status = (S.primal_infeasibility_solving && sc.status == PQP_PRIMAL_INFEASIBLE) ? PQP_SOLVED_CLOSEST_PRIMAL_FEASIBLE : PQP_SOLVED;
          break;
        }
      } else {
        sc.status = PQP_SOLVED;
        break;
      }
    }
    sc.iter_ext += 1;
    // x_prev..; shifted residuals (solver.hpp:1517-1559)
    for (int j = tid; j < n; j += NT) c.xp[j] = c.x[j];
    for (int j = tid; j < ne; j += NT) c.yp[j] = c.y[j];
    const double ag = (S.merit_function_type == PQP_MERIT_GPDAL) ? S.alpha_gpdal : 1.0;
    for (int i = tid; i < nc; i += NT) {
      const double zi = c.z[i];
      c.zp[i] = zi;
      double v = c.rup[i] * dli[i]; // scaled C x (box: scaled x-bound residual)
      v += zi * sc.mu_in;
      if (S.merit_function_type == PQP_MERIT_GPDAL) v += (S.alpha_gpdal - 1.0) * sc.mu_in * zi;
      c.rup[i] = v - c.us[i];
      c.si[i] = v - c.ls[i];
    }
    __syncthreads();

    // ---- inner loop: primal_dual_newton_semi_smooth (solver.hpp:884-1077) ----
    {
      const double eps_int = bcl_eta_in;
      for (long long it_in = 0; it_in <= S.max_iter_in; ++it_in) {
        if (it_in == S.max_iter_in) {
          sc.iter += S.max_iter_in + 1;
          break;
        }
        if (A.watchdog_ns) {
          if (tid == 0) c.iscratch[2 * NW + 1] = (gtimer_ns() - t_start > A.watchdog_ns) ? 1 : 0;
          __syncthreads();
          expired = c.iscratch[2 * NW + 1] != 0;
          __syncthreads();
          if (expired) break;
        }
        // -- Newton step (solver.hpp:756-869)
        for (int i = tid; i < nc; i += NT) {
          c.act_up[i] = c.rup[i] >= 0.0;
          c.act_low[i] = c.si[i] <= 0.0;
        }
        __syncthreads();
        active_set_change(c, sc);
        // q = sum over inactive constraints with z_i != 0 of z_i c_i
        int nq = block_compact(c, nc, c.list2, [&](int i) { return c.cons_slot[i] < 0 && c.z[i] != 0.0; });
        if (nq > 0) {
          rows_axpy_t(c, RowSrc{ nullptr, c.list2, 2 }, 0, nq, c.z, c.q, nullptr, 1.0);
        } else {
          for (int j = tid; j < n; j += NT) c.q[j] = 0;
        }
        for (int j = tid; j < n; j += NT) c.rx[j] = -c.dual[j] + c.q[j];
        for (int s = tid; s < c.ns; s += NT) {
          if (s < ne) {
            c.rs[s] = -c.se[s];
          } else {
            const int i = c.slot_cons[s];
            if (c.act_up[i])
              c.rs[s] = -c.rup[i] + c.z[i] * sc.mu_in * ag;
            else
              c.rs[s] = -c.si[i] + c.z[i] * sc.mu_in * ag;
          }
        }
        __syncthreads();
        iterative_solve(c, sc, S, eps_int);
        // un-permute dz; Cdx, CTdz (solver.hpp:860-967)
        for (int i = tid; i < nc; i += NT) {
          const int s = c.cons_slot[i];
          const double dzi = (s >= 0) ? c.ds[s] : -c.z[i];
          c.dz[i] = dzi;
          if (S.merit_function_type == PQP_MERIT_GPDAL) c.cdx[i] += (S.alpha_gpdal - 1.0) * sc.mu_in * dzi;
        }
        for (int j = tid; j < n; j += NT) c.ctdz[j] -= c.q[j];
        __syncthreads();
        double alpha = 1.0;
        if (ni > 0 || c.box) alpha = primal_dual_ls(c, sc, S);
        // |alpha dw|_inf
        {
          double m = 0;
          for (int j = tid; j < n; j += NT) m = nanmax(m, fabs(c.dx[j]));
          for (int j = tid; j < ne; j += NT) m = nanmax(m, fabs(c.ds[j]));
          for (int i = tid; i < nc; i += NT) m = nanmax(m, fabs(c.dz[i]));
          m = block_max1(c, m);
          if (m * fabs(alpha) < 1e-11 && it_in > 0) {
            sc.iter += it_in + 1;
            break;
          }
        }
        // iterate update + inner residual + infeasibility tests, fused
        double sm[6] = { 0, 0, 0, 0, 0, 0 }; // lb1 (primal inf), gdx
        double mx[10] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        // mx: 0 err_in | 1 |dy|u 2 |dz|u 3 |ATdy+CTdz|u 4 |dy|s,|dz|s any nonzero | 5 |dx|u 6 |Adx|u 7 |Hdx|u 8 first_cond violation 9 spare
        for (int j = tid; j < n; j += NT) {
          const double dxj = c.dx[j];
          c.x[j] += alpha * dxj;
          double dr = c.dual[j] + alpha * (sc.rho * dxj + ((c.hess == PQP_HESSIAN_ZERO) ? 0.0 : c.hdx[j]) + c.atdy[j] + c.ctdz[j]);
          c.dual[j] = dr;
          mx[0] = nanmax(mx[0], fabs(dr));
          const double dxc = dlx[j] * cs;
          mx[3] = nanmax(mx[3], fabs(c.atdy[j] / dxc + c.ctdz[j] / dxc));
          const double dxu = dxj * dlx[j];
          mx[5] = nanmax(mx[5], fabs(dxu));
          mx[7] = nanmax(mx[7], fabs(c.hdx[j] / dxc));
          sm[1] += dxj * c.gs[j];
        }
        for (int i = tid; i < ne; i += NT) {
          const double dyi = c.ds[i];
          double sev = c.se[i] + alpha * (c.adx[i] - sc.mu_eq * dyi);
          c.se[i] = sev;
          c.y[i] += alpha * dyi;
          mx[0] = nanmax(mx[0], fabs(sev));
          mx[4] = nanmax(mx[4], fabs(dyi));
          sm[0] += dyi * c.bs[i];
          mx[1] = nanmax(mx[1], fabs(dyi * dle[i] / cs));
          mx[6] = nanmax(mx[6], fabs(c.adx[i] / dle[i]));
        }
        for (int i = tid; i < nc; i += NT) {
          const double dzi = c.dz[i], cd = c.cdx[i];
          const double ru = c.rup[i] + alpha * cd;
          const double sl = c.si[i] + alpha * cd;
          const double zi = c.z[i] + alpha * dzi;
          c.rup[i] = ru;
          c.si[i] = sl;
          c.z[i] = zi;
          const double apz = fmax(ru, 0.0) + fmin(sl, 0.0) - ag * zi * sc.mu_in;
          mx[0] = nanmax(mx[0], fabs(apz));
          mx[4] = nanmax(mx[4], fabs(dzi));
          sm[0] += fmax(dzi, 0.0) * c.us[i] - fmin(dzi, 0.0) * c.ls[i];
          mx[2] = nanmax(mx[2], fabs(dzi * dli[i] / cs));
        }
        block_reduce<2, 8>(c, sm, mx);
        const double err_in = mx[0];
        if (it_in % S.frequence_infeasibility_check == 0 || S.primal_infeasibility_solving) {
          // utils.hpp:271-324
          bool is_primal_infeasible = false;
          if (mx[4] != 0.0) {
            const double upper = S.eps_primal_inf * fmax(mx[1], mx[2]);
            is_primal_infeasible = mx[3] <= upper && sm[0] <= -upper;
          }
          // utils.hpp:345-419
          bool is_dual_infeasible = false;
          {
            double bound = mx[5] * S.eps_dual_inf;
            double viol = 0;
            for (int i = tid; i < nc; i += NT) {
              const double v = c.cdx[i] / dli[i]; // unscaled (box entries use delta_box)
              bool ok = true;
              if (c.us[i] <= 1e20 && c.ls[i] >= -1e20)
                ok = v <= bound && v >= -bound;
              else if (c.us[i] > 1e20)
                ok = v >= -bound;
              else if (c.ls[i] < -1e20)
                ok = v <= bound;
              if (!ok) viol = 1.0;
            }
            viol = block_max1(c, viol);
            bool first_cond = mx[6] <= bound && viol == 0.0;
            bound *= cs;
            bool second = mx[7] <= bound && sm[1] <= -bound;
            is_dual_infeasible = first_cond && second && mx[5] != 0.0;
          }
          if (is_primal_infeasible) {
            sc.status = PQP_PRIMAL_INFEASIBLE;
            if (!S.primal_infeasibility_solving) {
              sc.iter += it_in + 1;
              break;
            }
          } else if (is_dual_infeasible) {
            sc.status = PQP_DUAL_INFEASIBLE;
            sc.iter += it_in + 1;
            break;
          }
        }
        if (err_in <= eps_int) {
          sc.iter += it_in + 1;
          break;
        }
      }
    }
    if (expired) break;
    if ((sc.status == PQP_PRIMAL_INFEASIBLE && !S.primal_infeasibility_solving) || sc.status == PQP_DUAL_INFEASIBLE) {
      // certificate of infeasibility: the (already unscaled, quirk 4) step
      for (int j = tid; j < n; j += NT) c.x[j] = c.dx[j] * dlx[j];
      for (int j = tid; j < ne; j += NT) c.y[j] = c.ds[j] * dle[j] / cs;
      for (int i = tid; i < nc; i += NT) c.z[i] = c.dz[i] * dli[i] / cs;
      __syncthreads();
      infeasible_exit = true;
      break;
    }
    if (scaled_eps == S.eps_abs && S.primal_infeasibility_solving && sc.status == PQP_PRIMAL_INFEASIBLE) {
      // solver.hpp:1581-1595
      for (int j = tid; j < c.cap; j += NT) c.s1[j] = 1.0;
      __syncthreads();
      rows_axpy_t(c, RowSrc{ c.Am, nullptr, 0 }, 0, ne, c.s1, c.t1, nullptr, 1.0);
      rows_axpy_t(c, RowSrc{ c.Cm, nullptr, 0 }, 0, ni, c.s1, c.t1, c.t1, 1.0);
      double m = 0;
      for (int j = tid; j < n; j += NT) m = nanmax(m, fabs(c.t1[j] + (c.box ? c.is[j] : 0.0)));
      scaled_eps = block_max1(c, m) * S.eps_abs;
    }
    global_primal_residual(c, sc, S, g);
    double primal_feasibility_lhs_new = g.pri_lhs;
    is_primal_feasible = primal_feasibility_lhs_new <= (scaled_eps + S.eps_rel * fmax(g.pri_eq_rhs0, g.pri_in_rhs0));
    info_pri = primal_feasibility_lhs_new;
    if (is_primal_feasible) {
      global_dual_residual(c, sc, g);
      info_dua = g.dua_lhs;
      info_gap = g.gap;
      is_dual_feasible = g.dua_lhs <= (S.eps_abs + S.eps_rel * fmax(fmax(g.dua_rhs3, g.dua_rhs0), fmax(g.dua_rhs1, dual_rhs2)));
      if (is_dual_feasible) {
        bool gap_ok = !S.check_duality_gap || fabs(g.gap) <= S.eps_duality_gap_abs + S.eps_duality_gap_rel * g.rhs_gap;
        if (gap_ok) sc.status = (S.primal_infeasibility_solving && sc.status == PQP_PRIMAL_INFEASIBLE) ? PQP_SOLVED_CLOSEST_PRIMAL_FEASIBLE : PQP_SOLVED;
      }
    }
    if (S.bcl_update) {
      // solver.hpp:566-614
      if (primal_feasibility_lhs_new <= bcl_eta_ext || sc.iter > S.safe_guard) {
        bcl_eta_ext *= pow(sc.mu_in, S.beta_bcl);
        bcl_eta_in = fmax(bcl_eta_in * sc.mu_in, eps_in_min);
      } else {
        for (int j = tid; j < ne; j += NT) c.y[j] = c.yp[j];
        for (int i = tid; i < nc; i += NT) c.z[i] = c.zp[i];
        __syncthreads();
        new_mu_in = fmax(sc.mu_in * S.mu_update_factor, S.mu_min_in);
        new_mu_eq = fmax(sc.mu_eq * S.mu_update_factor, S.mu_min_eq);
        new_mu_in_inv = fmin(sc.mu_in_inv * S.mu_update_inv_factor, S.mu_max_in_inv);
        new_mu_eq_inv = fmin(sc.mu_eq_inv * S.mu_update_inv_factor, S.mu_max_eq_inv);
        bcl_eta_ext = bcl_eta_ext_init * pow(new_mu_in, S.alpha_bcl);
        bcl_eta_in = fmax(new_mu_in, eps_in_min);
      }
    } else {
      // solver.hpp:639-677
      bcl_eta_in = fmax(bcl_eta_in * 0.1, eps_in_min);
      if (!(primal_feasibility_lhs_new <= 0.95 * primal_feasibility_lhs)) {
        new_mu_in = fmax(sc.mu_in * S.mu_update_factor, S.mu_min_in);
        new_mu_eq = fmax(sc.mu_eq * S.mu_update_factor, S.mu_min_eq);
        new_mu_in_inv = fmin(sc.mu_in_inv * S.mu_update_inv_factor, S.mu_max_in_inv);
        new_mu_eq_inv = fmin(sc.mu_eq_inv * S.mu_update_inv_factor, S.mu_max_eq_inv);
      }
    }
    global_dual_residual(c, sc, g);
    const double dual_feasibility_lhs_new = g.dua_lhs;
    info_dua = g.dua_lhs;
    info_gap = g.gap;
    if (primal_feasibility_lhs_new >= primal_feasibility_lhs && dual_feasibility_lhs_new >= dual_feasibility_lhs && sc.mu_in <= 1e-5) {
      new_mu_in = S.cold_reset_mu_in;
      new_mu_eq = S.cold_reset_mu_eq;
      new_mu_in_inv = S.cold_reset_mu_in_inv;
      new_mu_eq_inv = S.cold_reset_mu_eq_inv;
    }
    if (sc.mu_in != new_mu_in || sc.mu_eq != new_mu_eq) {
      ++sc.mu_updates;
      if (c.ns > 0) {
        rebuild_Ms_from_G(c, new_mu_eq, new_mu_in);
        sc.factor_fresh = false;
      }
    }
    sc.mu_eq = new_mu_eq;
    sc.mu_in = new_mu_in;
    sc.mu_eq_inv = new_mu_eq_inv;
    sc.mu_in_inv = new_mu_in_inv;
  }

  // ---- unscale and write back (solver.hpp:1749-1836) -------------------------
  double* xo = A.p.x + (size_t)q * n;
  double* yo = A.p.y + (size_t)q * ne;
  double* zo = A.p.z + (size_t)q * nc;
  double* seo = A.p.se + (size_t)q * ne;
  double* sio = A.p.si + (size_t)q * nc;
  const bool unscale_s = S.primal_infeasibility_solving && sc.status == PQP_PRIMAL_INFEASIBLE;
  for (int j = tid; j < n; j += NT) {
    const double xu = c.x[j] * dlx[j];
    c.t1[j] = xu;
    xo[j] = xu;
  }
  for (int j = tid; j < ne; j += NT) {
    yo[j] = c.y[j] * dle[j] / cs;
    seo[j] = unscale_s ? c.se[j] / dle[j] : c.se[j];
  }
  for (int i = tid; i < nc; i += NT) {
    zo[i] = c.z[i] * dli[i] / cs;
    sio[i] = unscale_s ? c.si[i] / dli[i] : c.si[i];
  }
  __syncthreads();
  (void)infeasible_exit;
  // objective 0.5 x^T H x + g^T x from the model (solver.hpp:1769-1781)
  double obj;
  {
    if (c.hess == PQP_HESSIAN_DENSE) {
      rows_dot(c, RowSrc{ c.Hm, nullptr, 0 }, 0, n, c.t1, c.t2);
    } else {
      for (int j = tid; j < n; j += NT) c.t2[j] = c.Hm[(size_t)j * n + j] * c.t1[j];
      __syncthreads();
    }
    double part = 0;
    const double* gm = A.p.g + (size_t)q * n;
    for (int j = tid; j < n; j += NT) part += c.t1[j] * (0.5 * c.t2[j] + gm[j]);
    obj = block_sum1(c, part);
  }
  if (tid == 0) {
    double* I = A.p.info + (size_t)q * PQP_INFO_DOUBLES;
    I[0] = sc.mu_eq;
    I[1] = sc.mu_eq_inv;
    I[2] = sc.mu_in;
    I[3] = sc.mu_in_inv;
    I[4] = sc.rho;
    I[5] = sc.nu;
    I[6] = (double)sc.iter;
    I[7] = (double)sc.iter_ext;
    I[8] = (double)sc.mu_updates;
    I[9] = 0.0;
    I[10] = (double)sc.status;
    I[11] = 0;
    I[12] = 0;
    I[13] = 0;
    I[14] = obj;
    I[15] = info_pri;
    I[16] = info_dua;
    I[17] = info_gap;
    I[18] = sc.iterative_residual;
    I[19] = S.default_H_eigenvalue_estimate;
  }
  __syncthreads();
}

extern __shared__ double smem_dyn[];

__global__ void __launch_bounds__(NT, 1) pqp_solve_kernel(PqpSolveArgs A)
{
  __shared__ Ctx c;
  __shared__ int cur_q;
  const PqpLayout& L = A.lay;
  if (threadIdx.x == 0) {
    double* ws = A.ws + (size_t)blockIdx.x * (size_t)L.ws_doubles;
    auto place = [&](int id) -> double* { return (L.in_smem[id] ? smem_dyn : ws) + L.off[id]; };
    c.n = A.d.n;
    c.ne = A.d.ne;
    c.ni = A.d.ni;
    c.nc = A.d.nc;
    c.box = A.d.box;
    c.hess = A.d.hess;
    c.cap = A.d.cap;
    c.ns = 0;
    c.M1 = place(PA_M1);
    c.As = place(PA_AS);
    c.Ms = place(PA_MS);
    c.G = place(PA_G);
    c.Y = place(PA_Y);
    double* v = place(PA_VEC);
    c.x = v + L.voff[V_X];
    c.y = v + L.voff[V_Y];
    c.z = v + L.voff[V_Z];
    c.xp = v + L.voff[V_XP];
    c.yp = v + L.voff[V_YP];
    c.zp = v + L.voff[V_ZP];
    c.dx = v + L.voff[V_DX];
    c.ds = v + L.voff[V_DS];
    c.dz = v + L.voff[V_DZ];
    c.rx = v + L.voff[V_RX];
    c.rs = v + L.voff[V_RS];
    c.ex = v + L.voff[V_EX];
    c.es = v + L.voff[V_ES];
    c.dual = v + L.voff[V_DUAL];
    c.se = v + L.voff[V_SE];
    c.rup = v + L.voff[V_RUP];
    c.si = v + L.voff[V_SI];
    c.hdx = v + L.voff[V_HDX];
    c.adx = v + L.voff[V_ADX];
    c.atdy = v + L.voff[V_ATDY];
    c.cdx = v + L.voff[V_CDX];
    c.ctdz = v + L.voff[V_CTDZ];
    c.q = v + L.voff[V_Q];
    c.gs = v + L.voff[V_GS];
    c.bs = v + L.voff[V_BS];
    c.us = v + L.voff[V_US];
    c.ls = v + L.voff[V_LS];
    c.is = v + L.voff[V_IS];
    c.delta = v + L.voff[V_DELTA];
    c.b = v + L.voff[V_B];
    c.u = v + L.voff[V_U];
    c.l = v + L.voff[V_L];
    c.d1inv = v + L.voff[V_D1INV];
    c.dsv = v + L.voff[V_DSV];
    c.dsinv = v + L.voff[V_DSINV];
    c.t1 = v + L.voff[V_T1];
    c.t2 = v + L.voff[V_T2];
    c.t3 = v + L.voff[V_T3];
    c.s1 = v + L.voff[V_S1];
    c.s2 = v + L.voff[V_S2];
    c.s3 = v + L.voff[V_S3];
    c.s4 = v + L.voff[V_S4];
    c.alphas = v + L.voff[V_ALPHAS];
    c.grads = v + L.voff[V_GRADS];
    c.scratch = v + L.voff[V_SCRATCH];
    c.red = v + L.voff[V_RED];
    int* ib = reinterpret_cast<int*>(smem_dyn + L.smem_doubles);
    c.cons_slot = ib;
    c.slot_cons = c.cons_slot + A.d.nc;
    c.list1 = c.slot_cons + A.d.cap;
    c.list2 = c.list1 + (A.d.nc > A.d.cap ? A.d.nc : A.d.cap);
    c.iscratch = c.list2 + A.d.nc;
    c.act_up = reinterpret_cast<unsigned char*>(c.iscratch + 2 * NW + 8);
    c.act_low = c.act_up + A.d.nc;
  }
  __syncthreads();
  double* As_home = c.As;
  while (true) {
    if (threadIdx.x == 0) cur_q = atomicAdd(A.counter, 1);
    __syncthreads();
    const int q = cur_q;
    __syncthreads();
    if (q >= A.batch) break;
    if (!A.p.params[q].active) continue;
    if (threadIdx.x == 0) c.As = As_home;
    __syncthreads();
    solve_one(c, A, q);
  }
}

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
  cudaError_t e = cudaFuncSetAttribute(pqp_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  pqp_solve_kernel<<<grid, NT, smem, (cudaStream_t)stream>>>(*a);
  return (int)cudaGetLastError();
}
