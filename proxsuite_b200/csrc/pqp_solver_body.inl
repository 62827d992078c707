// Device code of the solve kernel; included twice by pqp_kernels.cu (fast: every
// factor array and vector lives in shared memory and is accessed with LDS/STS;
// generic: any placement).


struct Ctx
{
  int n, ne, ni, nc, box, hess, cap;
  int ns; // current size of the dual block: ne + number of active inequalities
  // factor storage
  double *Pi, *As, *Si, *G, *Y;
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
  long long* prof; // per-phase cycle counters (shared memory) or NULL
  int vec_smem;    // 1: the vector arena is in shared memory
  int pi_smem;     // 1: P^-1 lives in shared memory (else global, read through L2)
  int si_cap;      // largest dual-block size the S^-1 storage can hold
  int uv_ld;       // leading dimension of the 8 sweep panel vectors kept in `scratch`
  int overflow;    // set when an insertion would exceed si_cap (QP is retried by the generic kernel)
};

// local (register) copies of the vector pointers with the address-space hint
#define PQP_VECS(c)   \
  double* const v_x = c.x; PQP_SM(v_x); (void)v_x;   \
  double* const v_y = c.y; PQP_SM(v_y); (void)v_y;   \
  double* const v_z = c.z; PQP_SM(v_z); (void)v_z;   \
  double* const v_xp = c.xp; PQP_SM(v_xp); (void)v_xp;   \
  double* const v_yp = c.yp; PQP_SM(v_yp); (void)v_yp;   \
  double* const v_zp = c.zp; PQP_SM(v_zp); (void)v_zp;   \
  double* const v_dx = c.dx; PQP_SM(v_dx); (void)v_dx;   \
  double* const v_ds = c.ds; PQP_SM(v_ds); (void)v_ds;   \
  double* const v_dz = c.dz; PQP_SM(v_dz); (void)v_dz;   \
  double* const v_rx = c.rx; PQP_SM(v_rx); (void)v_rx;   \
  double* const v_rs = c.rs; PQP_SM(v_rs); (void)v_rs;   \
  double* const v_ex = c.ex; PQP_SM(v_ex); (void)v_ex;   \
  double* const v_es = c.es; PQP_SM(v_es); (void)v_es;   \
  double* const v_dual = c.dual; PQP_SM(v_dual); (void)v_dual;   \
  double* const v_se = c.se; PQP_SM(v_se); (void)v_se;   \
  double* const v_rup = c.rup; PQP_SM(v_rup); (void)v_rup;   \
  double* const v_si = c.si; PQP_SM(v_si); (void)v_si;   \
  double* const v_hdx = c.hdx; PQP_SM(v_hdx); (void)v_hdx;   \
  double* const v_adx = c.adx; PQP_SM(v_adx); (void)v_adx;   \
  double* const v_atdy = c.atdy; PQP_SM(v_atdy); (void)v_atdy;   \
  double* const v_cdx = c.cdx; PQP_SM(v_cdx); (void)v_cdx;   \
  double* const v_ctdz = c.ctdz; PQP_SM(v_ctdz); (void)v_ctdz;   \
  double* const v_q = c.q; PQP_SM(v_q); (void)v_q;   \
  double* const v_gs = c.gs; PQP_SM(v_gs); (void)v_gs;   \
  double* const v_bs = c.bs; PQP_SM(v_bs); (void)v_bs;   \
  double* const v_us = c.us; PQP_SM(v_us); (void)v_us;   \
  double* const v_ls = c.ls; PQP_SM(v_ls); (void)v_ls;   \
  double* const v_is = c.is; PQP_SM(v_is); (void)v_is;   \
  double* const v_delta = c.delta; PQP_SM(v_delta); (void)v_delta;   \
  double* const v_b = c.b; PQP_SM(v_b); (void)v_b;   \
  double* const v_u = c.u; PQP_SM(v_u); (void)v_u;   \
  double* const v_l = c.l; PQP_SM(v_l); (void)v_l;   \
  double* const v_d1inv = c.d1inv; PQP_SM(v_d1inv); (void)v_d1inv;   \
  double* const v_t1 = c.t1; PQP_SM(v_t1); (void)v_t1;   \
  double* const v_t2 = c.t2; PQP_SM(v_t2); (void)v_t2;   \
  double* const v_t3 = c.t3; PQP_SM(v_t3); (void)v_t3;   \
  double* const v_s1 = c.s1; PQP_SM(v_s1); (void)v_s1;   \
  double* const v_s2 = c.s2; PQP_SM(v_s2); (void)v_s2;   \
  double* const v_s3 = c.s3; PQP_SM(v_s3); (void)v_s3;   \
  double* const v_s4 = c.s4; PQP_SM(v_s4); (void)v_s4;   \
  double* const v_alphas = c.alphas; PQP_SM(v_alphas); (void)v_alphas;   \
  double* const v_grads = c.grads; PQP_SM(v_grads); (void)v_grads;   \
  double* const v_scratch = c.scratch; PQP_SM(v_scratch); (void)v_scratch;   \
  double* const v_red = c.red; PQP_SM(v_red); (void)v_red;   \
  (void)0

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
  PQP_VECS(c);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < KS; ++k) sums[k] = warp_sum(sums[k]);
#pragma unroll
  for (int k = 0; k < KM; ++k) maxs[k] = warp_max(maxs[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < KS; ++k) v_red[warp * (KS + KM) + k] = sums[k];
#pragma unroll
    for (int k = 0; k < KM; ++k) v_red[warp * (KS + KM) + KS + k] = maxs[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += v_red[w * (KS + KM) + k];
    sums[k] = s;
  }
#pragma unroll
  for (int k = 0; k < KM; ++k) {
    double m = v_red[KS + k];
#pragma unroll
    for (int w = 1; w < NW; ++w) m = nanmax(m, v_red[w * (KS + KM) + KS + k]);
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
  PQP_VECS(c);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) v_red[warp] = v;
  __syncthreads();
  double off = 0;
  for (int w = 0; w < warp; ++w) off += v_red[w];
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

__device__ __forceinline__ int sym_off(int i)
{
  return (i * (i + 1)) >> 1; // packed lower WITH diagonal: row i has i+1 entries
}
__device__ __forceinline__ size_t gidx(int a, int b)
{
  int hi = a > b ? a : b, lo = a > b ? b : a;
  return (size_t)hi * (size_t)(hi + 1) / 2 + (size_t)lo;
}

// Reduce RR per-row partial sums across the 32 lanes with 1 + log2 steps per
// group instead of 5 shuffles per row. On return lane (32/RR)*r holds the sum
// of row r in d[0].
template<int RR>
__device__ __forceinline__ void reduce_rows(double (&d)[RR], int lane)
{
  int width = 16;
#pragma unroll
  for (int cnt = RR; cnt > 1; cnt >>= 1) {
    const bool hi = (lane & width) != 0;
#pragma unroll
    for (int k = 0; k < cnt / 2; ++k) {
      const double send = hi ? d[k] : d[k + cnt / 2];
      const double keep = hi ? d[k + cnt / 2] : d[k];
      d[k] = keep + __shfl_xor_sync(FULL, send, width);
    }
    width >>= 1;
  }
  for (; width >= 1; width >>= 1) d[0] += __shfl_xor_sync(FULL, d[0], width);
}

// ---------------------------------------------------------------------------
// Packed symmetric primitives (lower triangle with diagonal, row i at
// sym_off(i)). Rows are processed in blocks of 32: block b (rows 32b..32b+31)
// has b full 32-column chunks plus the diagonal chunk, so chunk counts are
// compile-time and only the diagonal chunk needs a per-lane guard. Warp w owns
// rows 32b + w + NW*r (r < RPB) of every block; all loads of a block are
// issued before the arithmetic and the stores.
// ---------------------------------------------------------------------------
#define RPB (32 / NW) // rows per warp per 32-row block

// y = T x. x and y must not alias. Uses c.scratch (NW x 32*NG doubles).
template<int NG, bool TSM>
__device__ void sym_mv_fast(const Ctx& c, const double* __restrict__ T, const double* __restrict__ x, double* __restrict__ y, int n)
{
  constexpr int NC = 32 * NG;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* const scr = c.scratch;
  if (TSM) PQP_SM(T);
  PQP_SM(x);
  PQP_SM(y);
  PQP_SM(scr);
  double acc[NG], xl[NG];
#pragma unroll
  for (int cc = 0; cc < NG; ++cc) {
    const int j = lane + 32 * cc;
    acc[cc] = 0.0;
    xl[cc] = (j < n) ? x[j] : 0.0;
  }
#pragma unroll
  for (int b = 0; b < NG; ++b) {
    if (32 * b < n) {
      double a[RPB][NG];
      double xi[RPB], d[RPB];
#pragma unroll
      for (int r = 0; r < RPB; ++r) {
        const int i = 32 * b + warp + NW * r;
        const bool ok = i < n;
        const double* row = T + sym_off(ok ? i : 0) + lane;
        xi[r] = ok ? x[i] : 0.0;
#pragma unroll
        for (int cc = 0; cc < b; ++cc) a[r][cc] = ok ? row[32 * cc] : 0.0;
        a[r][b] = (ok && lane + 32 * b <= i) ? row[32 * b] : 0.0;
      }
#pragma unroll
      for (int r = 0; r < RPB; ++r) {
        const int i = 32 * b + warp + NW * r;
        double dd = 0.0;
#pragma unroll
        for (int cc = 0; cc < b; ++cc) {
          dd += a[r][cc] * xl[cc];
          acc[cc] += a[r][cc] * xi[r];
        }
        dd += a[r][b] * xl[b];
        if (lane + 32 * b < i) acc[b] += a[r][b] * xi[r]; // the diagonal element only feeds the row sum
        d[r] = dd;
      }
      reduce_rows<RPB>(d, lane);
      if ((lane & (32 / RPB - 1)) == 0) {
        const int i = 32 * b + warp + NW * (lane / (32 / RPB));
        if (i < n) y[i] = d[0];
      }
    }
  }
#pragma unroll
  for (int cc = 0; cc < NG; ++cc) scr[warp * NC + lane + 32 * cc] = acc[cc];
  __syncthreads();
  _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) {
    double sacc = y[j];
#pragma unroll
    for (int w = 0; w < NW; ++w) sacc += scr[w * NC + j];
    y[j] = sacc;
  }
  __syncthreads();
}

// any n (shapes beyond the register-resident fast paths: cfg 4 / cfg 5 dual blocks). Packed lower storage:
//   y_i = sum_{j <= i} T[i][j] x_j  (row part: a warp owns row i, coalesced loads, one warp reduction)
//       + sum_{i' > i} T[i'][i] x_i' (column part, AXPY form: lane-stationary accumulators for a group of 256 columns)
// Columns are processed in groups of 256 so that the accumulators stay in registers; the partial vectors of the NW
// warps are combined through c.scratch (NW x n doubles). Row i belongs to the same lane-0 thread in every group
// (256 is a multiple of NW), so the row sums accumulate in a fixed order: results do not depend on scheduling.
__device__ void sym_mv_generic(const Ctx& c, const double* __restrict__ T, const double* __restrict__ x, double* __restrict__ y, int n)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* const scr = c.scratch;
  _Pragma("unroll 1") for (int g0 = 0; g0 < n; g0 += 256) {
    double acc[8], xl[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      const int j = g0 + lane + 32 * cc;
      acc[cc] = 0.0;
      xl[cc] = (j < n) ? x[j] : 0.0;
    }
    _Pragma("unroll 1") for (int i = g0 + warp; i < n; i += NW) {
      const double* row = T + (size_t)i * (size_t)(i + 1) / 2;
      const double xi = x[i];
      double a[8];
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        const int j = g0 + lane + 32 * cc;
        a[cc] = (j <= i) ? row[j] : 0.0;
      }
      double d = 0.0;
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        const int j = g0 + lane + 32 * cc;
        d = fma(a[cc], xl[cc], d);
        if (j < i) acc[cc] = fma(a[cc], xi, acc[cc]);
      }
      d = warp_sum(d);
      if (lane == 0) y[i] = (g0 == 0) ? d : y[i] + d;
    }
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      const int j = g0 + lane + 32 * cc;
      if (j < n) scr[(size_t)warp * n + j] = acc[cc];
    }
  }
  __syncthreads();
  _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) {
    double sacc = y[j];
#pragma unroll
    for (int w = 0; w < NW; ++w) sacc += scr[(size_t)w * n + j];
    y[j] = sacc;
  }
  __syncthreads();
}

__device__ __noinline__ void sym_mv(const Ctx& c, const double* T, const double* x, double* y, int n, bool t_smem)
{
  if (t_smem) {
    if (n <= 128)
      sym_mv_fast<4, true>(c, T, x, y, n);
    else if (n <= 160)
      sym_mv_fast<5, true>(c, T, x, y, n);
    else if (n <= 256)
      sym_mv_fast<8, true>(c, T, x, y, n);
    else
      sym_mv_generic(c, T, x, y, n);
  } else {
    if (n <= 128)
      sym_mv_fast<4, false>(c, T, x, y, n);
    else if (n <= 256)
      sym_mv_fast<8, false>(c, T, x, y, n);
    else
      sym_mv_generic(c, T, x, y, n);
  }
}

// T[i][j] += u_i * v_j on the packed lower triangle (j <= i < n). If kfix >= 0
// the element (kfix, kfix) is overwritten with dfix by the lane that owns it
// (used by the sweep operator).
template<int NG>
__device__ void sym_rank1_uv(double* __restrict__ T, const double* __restrict__ u, const double* __restrict__ v, int n, int kfix, double dfix)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  PQP_SM(T);
  PQP_SM(u);
  PQP_SM(v);
  double vl[NG];
#pragma unroll
  for (int cc = 0; cc < NG; ++cc) {
    const int j = lane + 32 * cc;
    vl[cc] = (j < n) ? v[j] : 0.0;
  }
#pragma unroll
  for (int b = 0; b < NG; ++b) {
    if (32 * b < n) {
      double a[RPB][NG];
      double ui[RPB];
      double* rowp[RPB];
#pragma unroll
      for (int r = 0; r < RPB; ++r) {
        const int i = 32 * b + warp + NW * r;
        const bool ok = i < n;
        rowp[r] = T + sym_off(ok ? i : 0) + lane;
        ui[r] = ok ? u[i] : 0.0;
#pragma unroll
        for (int cc = 0; cc < b; ++cc) a[r][cc] = ok ? rowp[r][32 * cc] : 0.0;
        a[r][b] = (ok && lane + 32 * b <= i) ? rowp[r][32 * b] : 0.0;
      }
#pragma unroll
      for (int r = 0; r < RPB; ++r) {
        const int i = 32 * b + warp + NW * r;
        if (i < n) {
#pragma unroll
          for (int cc = 0; cc < b; ++cc) rowp[r][32 * cc] = a[r][cc] + ui[r] * vl[cc];
          const int j = lane + 32 * b;
          if (j <= i) rowp[r][32 * b] = (i == kfix && j == kfix) ? dfix : a[r][b] + ui[r] * vl[b];
        }
      }
    }
  }
  __syncthreads();
}
__device__ void sym_rank1_uv_generic(double* __restrict__ T, const double* __restrict__ u, const double* __restrict__ v, int n, int kfix, double dfix)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < n; i += NW) {
    double* row = T + (size_t)i * (size_t)(i + 1) / 2;
    const double ui = u[i];
    for (int j = lane; j <= i; j += 32) row[j] = (i == kfix && j == kfix) ? dfix : row[j] + ui * v[j];
  }
  __syncthreads();
}
__device__ __noinline__ void sym_rank1(double* T, const double* u, const double* v, int n, int kfix, double dfix)
{
  if (n <= 128)
    sym_rank1_uv<4>(T, u, v, n, kfix, dfix);
  else if (n <= 160)
    sym_rank1_uv<5>(T, u, v, n, kfix, dfix);
  else if (n <= 256)
    sym_rank1_uv<8>(T, u, v, n, kfix, dfix);
  else
    sym_rank1_uv_generic(T, u, v, n, kfix, dfix);
}

// T[i][j] += sum_{k<4} U_k[i] * V_k[j] on the packed lower triangle (j <= i < n),
// accumulated k = 0..3 in order (the chain four scalar sweeps would produce).
// U_k = U + k*ldv, V_k = V + k*ldv. Rows are handled two at a time to keep the
// register footprint below the two-CTAs-per-SM budget.
template<int NG>
__device__ void sym_rank4_uv(double* __restrict__ T, const double* __restrict__ U, const double* __restrict__ V, int ldv, int n)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  PQP_SM(T);
  PQP_SM(U);
  PQP_SM(V);
  double vl[4][NG];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int cc = 0; cc < NG; ++cc) {
      const int j = lane + 32 * cc;
      vl[k][cc] = (j < n) ? V[k * ldv + j] : 0.0;
    }
  }
#pragma unroll
  for (int b = 0; b < NG; ++b) {
    if (32 * b < n) {
#pragma unroll
      for (int rr = 0; rr < RPB; rr += 2) {
        double a[2][NG];
        double u[2][4];
        double* rowp[2];
        int ii[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          ii[r] = 32 * b + warp + NW * (rr + r);
          if (ii[r] < n) {
            rowp[r] = T + sym_off(ii[r]) + lane;
#pragma unroll
            for (int k = 0; k < 4; ++k) u[r][k] = U[k * ldv + ii[r]];
#pragma unroll
            for (int cc = 0; cc < b; ++cc) a[r][cc] = rowp[r][32 * cc];
            if (lane + 32 * b <= ii[r]) a[r][b] = rowp[r][32 * b];
          }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          if (ii[r] < n) {
#pragma unroll
            for (int cc = 0; cc < b; ++cc)
              rowp[r][32 * cc] = fma(u[r][3], vl[3][cc], fma(u[r][2], vl[2][cc], fma(u[r][1], vl[1][cc], fma(u[r][0], vl[0][cc], a[r][cc]))));
            const int j = lane + 32 * b;
            if (j <= ii[r]) {
              rowp[r][32 * b] = fma(u[r][3], vl[3][b], fma(u[r][2], vl[2][b], fma(u[r][1], vl[1][b], fma(u[r][0], vl[0][b], a[r][b]))));
            }
          }
        }
      }
    }
  }
  __syncthreads();
}
__device__ void sym_rank4_uv_generic(double* __restrict__ T, const double* __restrict__ U, const double* __restrict__ V, int ldv, int n)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < n; i += NW) {
    double* row = T + (size_t)i * (size_t)(i + 1) / 2;
    const double u0 = U[i], u1 = U[ldv + i], u2 = U[2 * ldv + i], u3 = U[3 * ldv + i];
    for (int j = lane; j <= i; j += 32) {
      row[j] = fma(u3, V[3 * ldv + j], fma(u2, V[2 * ldv + j], fma(u1, V[ldv + j], fma(u0, V[j], row[j]))));
    }
  }
  __syncthreads();
}
__device__ __noinline__ void sym_rank4(double* T, const double* U, const double* V, int ldv, int n)
{
  if (n <= 128)
    sym_rank4_uv<4>(T, U, V, ldv, n);
  else if (n <= 160)
    sym_rank4_uv<5>(T, U, V, ldv, n);
  else if (n <= 256)
    sym_rank4_uv<8>(T, U, V, ldv, n);
  else
    sym_rank4_uv_generic(T, U, V, ldv, n);
}

// In-place inverse of an SPD matrix in packed storage by BLOCKED symmetric
// Gauss-Jordan sweeps (Goodnight's sweep operator, four pivots per pass over
// the triangle). One scalar sweep on pivot k maps
//   T_kk -> -1/d,  T_kj -> T_kj/d,  T_ij -> T_ij - T_ik T_kj / d   (d = T_kk).
// Four consecutive sweeps touch an entry outside the pivot rows/columns K only
// through   T_ij += sum_{k in K} U_ik V_kj,   U_ik = T^(k)_ik (column k just
// before its own sweep), V_kj = -U_jk / d_k, and U^(k) depends only on row i of
// the n x 4 panel T[:, K] plus the 4 x 4 pivot block. So every thread sweeps
// its own panel row in registers (the 4 x 4 block is swept redundantly by all
// threads), writes the finished K rows/columns straight back, stores U and V
// (zero on K), and ONE rank-4 pass applies the rest with the same chained FMAs
// the four scalar sweeps would have used: results match the scalar sweeps bit
// for bit, pivot rows are scaled exactly (no cancellation for large pivots).
// After all blocks the array holds -T^-1. `uv` is scratch for 8 vectors of
// length ldv >= n. Replaces Ldlt::factorize for the blocks this path inverts
// (linalg/dense/ldlt.hpp:718-744, factorize.hpp:91-148).
__device__ __noinline__ void sym_sweep_invert(double* __restrict__ T, double* __restrict__ uv, int ldv, int n)
{
  PQP_SM(T);
  PQP_SM(uv);
  double* const U = uv;
  double* const V = uv + 4 * ldv;
  for (int k0 = 0; k0 < n; k0 += 4) {
    const int kb = min(4, n - k0);
    double a0[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int bq = 0; bq < 4; ++bq) {
        const int hi = k0 + (a > bq ? a : bq), lo = k0 + (a > bq ? bq : a);
        a0[a][bq] = (a < kb && bq < kb) ? T[sym_off(hi) + lo] : ((a == bq) ? 1.0 : 0.0);
      }
    }
    __syncthreads(); // every thread holds the pivot block before anyone overwrites it
    _Pragma("unroll 1") for (int i = threadIdx.x; i < n; i += NT) {
      double a[4][4], p[4], ui[4], vi[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int r = 0; r < 4; ++r) a[q][r] = a0[q][r];
      }
      const int ai = i - k0; // position of this row inside the pivot block when 0 <= ai < kb
      const bool inK = (ai >= 0) && (ai < kb);
      // rows of the pivot block take their panel from the copy every thread holds (a0): their entries (col, i),
      // i < col, are being overwritten by the thread that owns row col in this very loop
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const int col = k0 + l;
        double v = 0.0;
        if (l < kb) {
          if (inK) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              if (ai == a) v = a0[a][l];
            }
          } else {
            v = (i >= col) ? T[sym_off(i) + col] : T[sym_off(col) + i];
          }
        }
        p[l] = v;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < kb) {
          const double inv = 1.0 / a[k][k];
          double vK[4];
#pragma unroll
          for (int l = 0; l < 4; ++l) vK[l] = -a[k][l] * inv;
          const double pk = p[k];
          if (ai == k) {
            ui[k] = 0.0;
            vi[k] = 0.0;
#pragma unroll
            for (int l = 0; l < 4; ++l) p[l] = (l == k) ? -inv : p[l] * inv;
          } else {
            ui[k] = inK ? 0.0 : pk;
            vi[k] = inK ? 0.0 : -pk * inv;
#pragma unroll
            for (int l = 0; l < 4; ++l) p[l] = (l == k) ? pk * inv : fma(pk, vK[l], p[l]);
          }
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            if (m != k) {
              const double amk = a[m][k];
#pragma unroll
              for (int l = 0; l < 4; ++l) a[m][l] = (l == k) ? amk * inv : fma(amk, vK[l], a[m][l]);
            }
          }
#pragma unroll
          for (int l = 0; l < 4; ++l) a[k][l] = (l == k) ? -inv : a[k][l] * inv;
        } else {
          ui[k] = 0.0;
          vi[k] = 0.0;
        }
      }
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const int col = k0 + l;
        if (l < kb) {
          if (i >= col)
            T[sym_off(i) + col] = p[l];
          else if (!inK)
            T[sym_off(col) + i] = p[l];
        }
        U[l * ldv + i] = ui[l];
        V[l * ldv + i] = vi[l];
      }
    }
    __syncthreads();
    sym_rank4(T, U, V, ldv, n);
  }
  const int tot = sym_off(n);
  _Pragma("unroll 1") for (int e = threadIdx.x; e < tot; e += NT) T[e] = -T[e];
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

// Fused streaming pass over a set of rows (the HBM/L2-facing primitive):
//   out_dot[idx(r)]  = row_r . x                                       (if x != null)
//   out_axpy[j]      = add[j] + sign * sum_r coef[idx(r)] row_r[j]     (if coef != null)
// Warp w owns rows r0 + w + NW*k and handles them four at a time: the eight
// 16-byte loads of a group are issued together, the four row sums share one
// transpose-reduction, and every matrix element is loaded once for both
// products. n must be even and <= 128 (two double2 per lane).
template<int MODE>
__device__ void mat_pass_fast(const Ctx& c, RowSrc rs, int r0, int r1, const double* __restrict__ x, double* __restrict__ out_dot, const double* __restrict__ coef, double* out_axpy, const double* add, double sign)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = c.n, n2 = c.n >> 1;
  const bool DOT = x != nullptr, AXPY = coef != nullptr;
  double* const scr = c.scratch;
  const double* const isv = c.is;
  PQP_SM(scr);
  PQP_SM(isv);
  if (DOT) {
    PQP_SM(x);
    PQP_SM(out_dot);
  }
  if (AXPY) {
    PQP_SM(coef);
    PQP_SM(out_axpy);
  }
  const bool l1 = lane + 32 < n2; // second double2 of the row belongs to this lane
  double2 x0 = make_double2(0.0, 0.0), x1 = x0, acc0 = x0, acc1 = x0;
  if (DOT) {
    if (lane < n2) x0 = reinterpret_cast<const double2*>(x)[lane];
    if (l1) x1 = reinterpret_cast<const double2*>(x)[lane + 32];
  }
  RowSrc one = rs;
  one.mode = MODE;
  for (int rb = r0 + warp; rb < r1; rb += 4 * NW) {
    const double* rowp[4];
    int bk[4], idx[4];
    double2 v0[4], v1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = rb + u * NW;
      rowp[u] = nullptr;
      bk[u] = -1;
      idx[u] = -1;
      if (r < r1) rowp[u] = get_row(c, one, r, bk[u], idx[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v0[u] = (rowp[u] && lane < n2) ? reinterpret_cast<const double2*>(rowp[u])[lane] : make_double2(0.0, 0.0);
      v1[u] = (rowp[u] && l1) ? reinterpret_cast<const double2*>(rowp[u])[lane + 32] : make_double2(0.0, 0.0);
    }
    if (DOT) {
      double d[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) d[u] = (v0[u].x * x0.x + v0[u].y * x0.y) + (v1[u].x * x1.x + v1[u].y * x1.y);
      reduce_rows<4>(d, lane);
      if ((lane & 7) == 0) {
        const int u = lane >> 3;
        // (rowp / idx are warp-uniform per u; select without dynamic indexing)
        const int id = (u == 0) ? idx[0] : (u == 1) ? idx[1] : (u == 2) ? idx[2] : idx[3];
        const int bb = (u == 0) ? bk[0] : (u == 1) ? bk[1] : (u == 2) ? bk[2] : bk[3];
        if (id >= 0) out_dot[id] = (bb < 0) ? d[0] : isv[bb] * x[bb];
      }
    }
    if (AXPY) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (idx[u] >= 0) {
          const double cf = coef[idx[u]];
          if (bk[u] < 0) {
            acc0.x += cf * v0[u].x;
            acc0.y += cf * v0[u].y;
            acc1.x += cf * v1[u].x;
            acc1.y += cf * v1[u].y;
          } else {
            const int jv = bk[u] >> 1;
            const double add_v = cf * isv[bk[u]];
            if (jv == lane) {
              if (bk[u] & 1) acc0.y += add_v; else acc0.x += add_v;
            } else if (jv == lane + 32) {
              if (bk[u] & 1) acc1.y += add_v; else acc1.x += add_v;
            }
          }
        }
      }
    }
  }
  if (AXPY) {
    if (lane < n2) reinterpret_cast<double2*>(scr)[warp * n2 + lane] = acc0;
    if (l1) reinterpret_cast<double2*>(scr)[warp * n2 + lane + 32] = acc1;
    __syncthreads();
    _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) {
      double sacc = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) sacc += scr[w * n + j];
      out_axpy[j] = (add ? add[j] : 0.0) + sign * sacc;
    }
  }
  __syncthreads();
}

// generic fallbacks (n > 256)
__device__ void rows_dot(const Ctx& c, RowSrc rs, int r0, int r1, const double* __restrict__ x, double* __restrict__ out)
{
  PQP_VECS(c);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = c.n;
  // four rows of a warp in flight (one row at a time is one L2 round trip per 32 columns of every row)
  _Pragma("unroll 1") for (int rb = r0 + warp; rb < r1; rb += 4 * NW) {
    const double* rowp[4];
    int bk[4], idx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = rb + u * NW;
      rowp[u] = nullptr;
      bk[u] = -1;
      idx[u] = -1;
      if (r < r1) rowp[u] = get_row(c, rs, r, bk[u], idx[u]);
    }
    double d[4] = { 0.0, 0.0, 0.0, 0.0 };
    _Pragma("unroll 2") for (int j = lane; j < n; j += 32) {
      const double xj = x[j];
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = rowp[u] ? rowp[u][j] : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) d[u] = fma(v[u], xj, d[u]);
    }
    reduce_rows<4>(d, lane);
    if ((lane & 7) == 0) {
      const int u = lane >> 3;
      const int id = (u == 0) ? idx[0] : (u == 1) ? idx[1] : (u == 2) ? idx[2] : idx[3];
      const int bb = (u == 0) ? bk[0] : (u == 1) ? bk[1] : (u == 2) ? bk[2] : bk[3];
      if (id >= 0) out[id] = (bb < 0) ? d[0] : v_is[bb] * x[bb];
    }
  }
  __syncthreads();
}
// out[j] = add[j] + sign * sum_r coef[idx(r)] row_r[j], any n: warp w owns rows r0 + w + NW k (two in flight), a lane
// owns the columns lane + 32 cc of a 256-column group (register accumulators), the NW partial vectors are combined
// through c.scratch (NW x n doubles). (The first version gave every thread one column and walked all rows
// sequentially: one dependent L2 round trip per row - cfg 5 spent seconds per QP there.)
__device__ void rows_axpy_t(const Ctx& c, RowSrc rs, int r0, int r1, const double* __restrict__ coef, double* out, const double* add, double sign)
{
  PQP_VECS(c);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = c.n;
  double* const scr = c.scratch;
  _Pragma("unroll 1") for (int g0 = 0; g0 < n; g0 += 256) {
    double acc[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) acc[cc] = 0.0;
    _Pragma("unroll 1") for (int rb = r0 + warp; rb < r1; rb += 2 * NW) {
      const double* rowp[2];
      int bk[2], idx[2];
      double cf[2];
      double v[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = rb + u * NW;
        rowp[u] = nullptr;
        bk[u] = -1;
        idx[u] = -1;
        cf[u] = 0.0;
        if (r < r1) {
          rowp[u] = get_row(c, rs, r, bk[u], idx[u]);
          cf[u] = coef[idx[u]];
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const int j = g0 + lane + 32 * cc;
          v[u][cc] = (rowp[u] && j < n) ? rowp[u][j] : 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (idx[u] >= 0) {
          if (bk[u] < 0) {
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) acc[cc] = fma(cf[u], v[u][cc], acc[cc]);
          } else { // box row i_s[k] e_k
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              if (bk[u] == g0 + lane + 32 * cc) acc[cc] = fma(cf[u], v_is[bk[u]], acc[cc]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      const int j = g0 + lane + 32 * cc;
      if (j < n) scr[(size_t)warp * n + j] = acc[cc];
    }
  }
  __syncthreads();
  _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) {
    double sacc = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) sacc += scr[(size_t)w * n + j];
    out[j] = (add ? add[j] : 0.0) + sign * sacc;
  }
  __syncthreads();
}

__device__ __noinline__ void mat_pass(const Ctx& c, RowSrc rs, int r0, int r1, const double* x, double* out_dot, const double* coef, double* out_axpy, const double* add, double sign)
{
  const int n = c.n;
  if ((n & 1) == 0 && n <= 128 && c.vec_smem) {
    switch (rs.mode) {
      case 0: mat_pass_fast<0>(c, rs, r0, r1, x, out_dot, coef, out_axpy, add, sign); break;
      case 1: mat_pass_fast<1>(c, rs, r0, r1, x, out_dot, coef, out_axpy, add, sign); break;
      case 2: mat_pass_fast<2>(c, rs, r0, r1, x, out_dot, coef, out_axpy, add, sign); break;
      default: mat_pass_fast<3>(c, rs, r0, r1, x, out_dot, coef, out_axpy, add, sign); break;
    }
  } else {
    if (x) rows_dot(c, rs, r0, r1, x, out_dot);
    if (coef) rows_axpy_t(c, rs, r0, r1, coef, out_axpy, add, sign);
  }
}

// y = P^-1 v  (P = Hs + rho I, Pi = P^-1 explicit, packed). v must not alias y.
__device__ void apply_Pinv(const Ctx& c, const double* v, double* y)
{
  PQP_VECS(c);
  if (c.hess != PQP_HESSIAN_DENSE) {
    _Pragma("unroll 1") for (int j = threadIdx.x; j < c.n; j += NT) y[j] = v[j] * v_d1inv[j];
    __syncthreads();
    return;
  }
  sym_mv(c, c.Pi, v, y, c.n, c.pi_smem != 0);
}

__device__ __forceinline__ int row_id(const Ctx& c, int s)
{
  return s < c.ne ? s : c.ne + c.slot_cons[s];
}

// Solve K [ox; os] = [b1; b2],  K = [P B^T; B -Dlt], with the explicit block
// inverses:  t = P^-1 b1;  lam = S^-1 (B t - b2);  x = P^-1 (b1 - B^T lam).
// In place allowed (ox == b1, os == b2). Replaces Ldlt::solve_in_place
// (ldlt.hpp:767-782).
__device__ __noinline__ void solve_kkt(const Ctx& c, const double* b1, const double* b2, double* ox, double* os)
{
  PQP_VECS(c);
  const int ns = c.ns;
  if (ns == 0) {
    apply_Pinv(c, b1, v_t1);
    _Pragma("unroll 1") for (int j = threadIdx.x; j < c.n; j += NT) ox[j] = v_t1[j];
    __syncthreads();
    return;
  }
  apply_Pinv(c, b1, v_t1);
  RowSrc slots{ nullptr, nullptr, 1 };
  mat_pass(c, slots, 0, ns, v_t1, v_s1, nullptr, nullptr, nullptr, 1.0);
  _Pragma("unroll 1") for (int s = threadIdx.x; s < ns; s += NT) v_s1[s] -= b2[s];
  __syncthreads();
  sym_mv(c, c.Si, v_s1, os, ns, true);
  mat_pass(c, slots, 0, ns, nullptr, nullptr, os, v_t2, b1, -1.0);
  apply_Pinv(c, v_t2, ox);
}

// Gram row of dual slot s against slots 0..s:  y = P^-1 b_s (-> t1),
// s3[j] = b_j . y, stored in G by row id.
__device__ void gram_row(Ctx& c, int s)
{
  PQP_VECS(c);
  RowSrc slots{ nullptr, nullptr, 1 };
  int bk, idx;
  const double* row = get_row(c, slots, s, bk, idx);
  // stage the row in shared memory (rows of C_s live in global memory)
  _Pragma("unroll 1") for (int j = threadIdx.x; j < c.n; j += NT) v_t2[j] = row ? row[j] : ((j == bk) ? v_is[bk] : 0.0);
  __syncthreads();
  row = v_t2;
  apply_Pinv(c, row, v_t1);
  mat_pass(c, slots, 0, s + 1, v_t1, v_s3, nullptr, nullptr, nullptr, 1.0);
  const int ids = row_id(c, s);
  _Pragma("unroll 1") for (int j = threadIdx.x; j <= s; j += NT) c.G[gidx(ids, row_id(c, j))] = v_s3[j];
}

// Append dual slot s == c.ns (already registered in slot_cons) with proximal
// parameter mu: bordering of the explicit inverse
//   w = S^-1 g, delta = (b.P^-1 b + mu) - g.w,
//   S^-1 <- [S^-1 + w w^T/delta, -w/delta; -w^T/delta, 1/delta].
// Replaces Ldlt::insert_block_at (ldlt.hpp:431-475, modify.hpp:131-264).
__device__ __noinline__ void rebuild_Si_from_G(Ctx& c, double mu_eq, double mu_in);
__device__ __noinline__ void insert_slot(Ctx& c, double mu, double mu_eq)
{
  PQP_VECS(c);
  const int s = c.ns;
  if (s + 1 > c.si_cap) { // does not fit the shared-memory S^-1: hand the QP to the generic kernel
    if (threadIdx.x == 0) c.overflow = 1;
    __syncthreads();
    return;
  }
  gram_row(c, s);
  __syncthreads();
  double delta = v_s3[s] + mu;
  if (s > 0) {
    sym_mv(c, c.Si, v_s3, v_s1, s, true);
    double part = 0;
    _Pragma("unroll 1") for (int j = threadIdx.x; j < s; j += NT) part += v_s3[j] * v_s1[j];
    delta -= block_sum1(c, part);
    // delta = (b.P^-1 b + mu) - g.w is the Schur complement of the new slot in S; it cancels when the new row is nearly
    // dependent on the active ones (terms ~ |b|^2 / rho against mu). A non-positive or fully cancelled value means
    // the bordering formula has no accuracy left: register the slot and re-form S^-1 from the Gram matrix instead
    // (one sweep inversion; block-uniform: delta comes out of a block reduction).
    if (!(delta > 1e-13 * (v_s3[s] + mu))) {
      __syncthreads();
      if (threadIdx.x == 0) c.ns = s + 1;
      __syncthreads();
      rebuild_Si_from_G(c, mu_eq, mu);
      return;
    }
    const double dinv = 1.0 / delta;
    double* row = c.Si + sym_off(s);
    _Pragma("unroll 1") for (int j = threadIdx.x; j < s; j += NT) {
      const double wj = v_s1[j] * dinv;
      v_s2[j] = wj;
      row[j] = -wj;
    }
    __syncthreads();
    sym_rank1(c.Si, v_s1, v_s2, s, -1, 0.0);
  }
  if (threadIdx.x == 0) {
    c.Si[sym_off(s) + s] = 1.0 / delta;
    c.ns = s + 1;
  }
  __syncthreads();
}

// Remove dual slot k (k >= ne): Schur complement of the explicit inverse,
//   S'^-1 = T - q q^T / T_kk   (T = S^-1 without row/column k, q = column k),
// then the last slot takes the place of k. Replaces Ldlt::delete_at (ldlt.hpp:340-387).
__device__ __noinline__ void delete_slot(Ctx& c, int k)
{
  PQP_VECS(c);
  const int ns = c.ns;
  double* T = c.Si;
  _Pragma("unroll 1") for (int i = threadIdx.x; i < ns; i += NT) v_s1[i] = (i >= k) ? T[sym_off(i) + k] : T[sym_off(k) + i];
  __syncthreads();
  const double sinv = -1.0 / v_s1[k];
  __syncthreads();
  _Pragma("unroll 1") for (int i = threadIdx.x; i < ns; i += NT) {
    const double q = (i == k) ? 0.0 : v_s1[i]; // row / column k are dropped below
    v_s2[i] = q;
    v_s3[i] = q * sinv;
  }
  __syncthreads();
  sym_rank1(T, v_s2, v_s3, ns, -1, 0.0);
  // drop row / column k: the LAST slot takes the freed position (slot order carries no meaning; the first version
  // compacted the packed triangle in place, order preserving: O(ns^2) element moves behind two barriers per 256
  // elements - the dominant cost of a deletion for the large shapes)
  const int L = ns - 1;
  if (k != L) {
    const double* rowL = T + sym_off(L);
    _Pragma("unroll 1") for (int i = threadIdx.x; i <= L; i += NT) v_s1[i] = rowL[i];
    __syncthreads();
    _Pragma("unroll 1") for (int i = threadIdx.x; i < L; i += NT) {
      if (i == k)
        T[sym_off(k) + k] = v_s1[L];
      else if (i > k)
        T[sym_off(i) + k] = v_s1[i];
      else
        T[sym_off(k) + i] = v_s1[i];
    }
  }
  if (threadIdx.x == 0) {
    const int cons_k = c.slot_cons[k], cons_L = c.slot_cons[L];
    if (k != L) {
      c.slot_cons[k] = cons_L;
      c.cons_slot[cons_L] = k;
    }
    c.cons_slot[cons_k] = -1;
    c.ns = L;
  }
  __syncthreads();
}

// S^-1 from the cached Gram matrix with the given proximal parameters
// (S = Dlt + G): gather + sweep inversion. Replaces
// Ldlt::diagonal_update_clobber_indices (ldlt.hpp:516-570) used by mu_update
// (solver.hpp:130-169).
__device__ __noinline__ void rebuild_Si_from_G(Ctx& c, double mu_eq, double mu_in)
{
  PQP_VECS(c);
  const int ns = c.ns;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int s = warp; s < ns; s += NW) {
    const int ids = row_id(c, s);
    double* row = c.Si + sym_off(s);
    for (int j = lane; j <= s; j += 32) row[j] = c.G[gidx(ids, row_id(c, j))] + ((j == s) ? (s < c.ne ? mu_eq : mu_in) : 0.0);
  }
  __syncthreads();
  sym_sweep_invert(c.Si, v_scratch, c.uv_ld, ns);
}

// P^-1 = (Hs + rho I)^-1, explicit. Replaces the x-block part of
// Ldlt::factorize (ldlt.hpp:718-744).
__device__ __noinline__ void build_Pi(Ctx& c, double rho)
{
  PQP_VECS(c);
  const int n = c.n;
  if (c.hess != PQP_HESSIAN_DENSE) {
    _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) {
      double h = (c.hess == PQP_HESSIAN_DIAGONAL) ? c.Hs[(size_t)j * n + j] : 0.0;
      v_d1inv[j] = 1.0 / (h + rho);
    }
    __syncthreads();
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // When P^-1 is kept in global memory (compact shared-memory layout, two CTAs
  // per SM) the sweeps still run in shared memory: the S^-1 region is free at
  // this point (the dual block is always rebuilt afterwards).
  double* const work = c.pi_smem ? c.Pi : c.Si;
  for (int i = warp; i < n; i += NW) {
    double* row = work + sym_off(i);
    const double* h = c.Hs + (size_t)i * n;
    for (int j = lane; j <= i; j += 32) row[j] = h[j] + ((j == i) ? rho : 0.0);
  }
  __syncthreads();
  sym_sweep_invert(work, v_scratch, c.uv_ld, n);
  if (!c.pi_smem) {
    const int tot = sym_off(n);
    _Pragma("unroll 1") for (int e = threadIdx.x; e < tot; e += NT) c.Pi[e] = work[e];
    __syncthreads();
  }
}

// (Re)build the dual block for the slots 0..ns_target-1 currently registered:
// Gram rows, then one sweep inversion. Used for the first factorisation
// (equality rows only, helpers.hpp:241-285) and by refactorize (solver.hpp:40-87).
__device__ __noinline__ void build_dual_block(Ctx& c, int ns_target, double mu_eq, double mu_in)
{
  for (int s = 0; s < ns_target; ++s) {
    gram_row(c, s);
    __syncthreads();
  }
  if (threadIdx.x == 0) c.ns = ns_target;
  __syncthreads();
  if (ns_target > 0) rebuild_Si_from_G(c, mu_eq, mu_in);
}

// optional per-phase cycle accounting (thread 0 only; enabled when args.prof != NULL)
enum ProfPhase { PH_STAGE = 0, PH_M1, PH_EQ, PH_INSERT, PH_DELETE, PH_SOLVE, PH_RESID, PH_LS, PH_MU, PH_GLOBAL, PH_NEWTON_MISC, PH_TOTAL, PH_COUNT };
#define PROF_T0() (c.prof ? clock64() : 0ll)
#define PROF_ADD(ph, t0)                                                                                                                                                                                                                                       \
  do {                                                                                                                                                                                                                                                         \
    if (c.prof && threadIdx.x == 0) c.prof[ph] += clock64() - (t0);                                                                                                                                                                                           \
  } while (0)

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
__device__ __noinline__ double kkt_residual(const Ctx& c, const Scal& sc)
{
  PQP_VECS(c);
  const int n = c.n, ne = c.ne, ns = c.ns;
  // per-constraint coefficient of the transposed product: dz of active rows, 0 otherwise
  _Pragma("unroll 1") for (int i = threadIdx.x; i < c.nc; i += NT) {
    const int s = c.cons_slot[i];
    v_dz[i] = (s >= 0) ? v_ds[s] : 0.0;
  }
  if (c.hess != PQP_HESSIAN_DENSE) {
    _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) v_hdx[j] = (c.hess == PQP_HESSIAN_DIAGONAL) ? c.Hs[(size_t)j * n + j] * v_dx[j] : 0.0;
  }
  __syncthreads();
  if (c.hess == PQP_HESSIAN_DENSE) mat_pass(c, RowSrc{ c.Hs, nullptr, 0 }, 0, n, v_dx, v_hdx, nullptr, nullptr, nullptr, 1.0);
  // one pass over A: A dx and A^T dy ; one pass over C (+box): C dx and C_J^T dz_J
  mat_pass(c, RowSrc{ c.As, nullptr, 0 }, 0, ne, v_dx, v_adx, v_ds, v_atdy, nullptr, 1.0);
  mat_pass(c, RowSrc{ nullptr, nullptr, 3 }, 0, c.nc, v_dx, v_cdx, v_dz, v_ctdz, nullptr, 1.0);
  double m = 0;
  _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) {
    double e = v_rx[j] - (v_hdx[j] + sc.rho * v_dx[j] + v_atdy[j] + v_ctdz[j]);
    v_ex[j] = e;
    m = nanmax(m, fabs(e));
  }
  _Pragma("unroll 1") for (int s = threadIdx.x; s < ns; s += NT) {
    double e;
    if (s < ne)
      e = v_rs[s] - (v_adx[s] - sc.mu_eq * v_ds[s]);
    else
      e = v_rs[s] - (v_cdx[c.slot_cons[s]] - sc.mu_in * v_ds[s]);
    v_es[s] = e;
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
  build_Pi(c, sc.rho);
  build_dual_block(c, ns_target, sc.mu_eq, sc.mu_in);
  sc.factor_fresh = true;
}

#ifndef PQP_TIGHT_REFINE
#define PQP_TIGHT_REFINE 1e-10
#endif
// solver.hpp:408-541
__device__ __noinline__ void iterative_solve(Ctx& c, Scal& sc, const pqp_settings& S, double eps)
{
  PQP_VECS(c);
  for (int pass = 0; pass < 2; ++pass) {
    int it = 0, it_stab = 0;
    long long tp = PROF_T0();
    solve_kkt(c, v_rx, v_rs, v_dx, v_ds);
    PROF_ADD(PH_SOLVE, tp);
    tp = PROF_T0();
    double err = kkt_residual(c, sc);
    PROF_ADD(PH_RESID, tp);
    ++it;
    double prev = err;
    // Diagonal / zero Hessians: P^-1 = diag(1 / (H_ii + rho)) reaches 1 / rho where H_ii = 0, the Schur complement
    // is then badly conditioned and the explicit S^-1 carries ~1e-8 relative error. The reference's LDL^T does not
    // have that loss, so its first solve is already accurate; here the refinement (one solve + one residual per
    // round, error shrinking ~1e-8 per round) is driven down to PQP_TIGHT_REFINE instead of stopping at the inner
    // tolerance, which keeps the Newton steps as good as the reference's (3-6x fewer Newton iterations on the
    // diagonal-Hessian benchmark family, measured on the emulator).
    const double eps_ref = (c.hess != PQP_HESSIAN_DENSE) ? fmin(eps, PQP_TIGHT_REFINE) : eps;
    while (err >= eps_ref) {
      if (it >= S.nb_iterative_refinement) break;
      ++it;
      tp = PROF_T0();
      solve_kkt(c, v_ex, v_es, v_ex, v_es);
      _Pragma("unroll 1") for (int j = threadIdx.x; j < c.n; j += NT) v_dx[j] += v_ex[j];
      _Pragma("unroll 1") for (int s = threadIdx.x; s < c.ns; s += NT) v_ds[s] += v_es[s];
      __syncthreads();
      PROF_ADD(PH_SOLVE, tp);
      tp = PROF_T0();
      err = kkt_residual(c, sc);
      PROF_ADD(PH_RESID, tp);
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
  _Pragma("unroll 1") for (int j = threadIdx.x; j < c.n; j += NT) v_rx[j] = 0;
  _Pragma("unroll 1") for (int s = threadIdx.x; s < c.cap; s += NT) v_rs[s] = 0;
  __syncthreads();
}

// linesearch.hpp:551-786 with act[i] = act_up | act_low
__device__ __noinline__ void active_set_change(Ctx& c, Scal& sc)
{
  // deletions, from the last slot to the first
  int ndel = block_compact(c, c.ns - c.ne, c.list1, [&](int k) {
    int cons = c.slot_cons[c.ne + k];
    return !(c.act_up[cons] || c.act_low[cons]);
  });
  long long tp = PROF_T0();
  for (int k = ndel - 1; k >= 0; --k) delete_slot(c, c.ne + c.list1[k]);
  PROF_ADD(PH_DELETE, tp);
  tp = PROF_T0();
  int nadd = block_compact(c, c.nc, c.list1, [&](int i) { return (c.act_up[i] || c.act_low[i]) && c.cons_slot[i] < 0; });
  for (int k = 0; k < nadd; ++k) {
    if (threadIdx.x == 0) {
      int cons = c.list1[k];
      c.slot_cons[c.ns] = cons;
      c.cons_slot[cons] = c.ns;
    }
    __syncthreads();
    insert_slot(c, sc.mu_in, sc.mu_eq);
  }
  PROF_ADD(PH_INSERT, tp);
  if (ndel > 0 || nadd > 0) sc.factor_fresh = false;
}

// unscaled global residual pieces -------------------------------------------------
struct Glob
{
  double pri_lhs, pri_eq_rhs0, pri_in_rhs0, pri_eq_lhs, pri_in_lhs;
  double dua_lhs, dua_rhs0, dua_rhs1, dua_rhs3, gap, rhs_gap;
};

// utils.hpp:166-252
// Streaming passes shared by the global residuals: H x -> t1, A x -> se,
// A^T y -> t2, C x -> rup, C^T z_C -> t3 (one pass per matrix).
__device__ __noinline__ void global_passes(Ctx& c, bool primal, bool dual)
{
  PQP_VECS(c);
  const int n = c.n, ne = c.ne, ni = c.ni;
  if (dual) {
    if (c.hess == PQP_HESSIAN_DENSE) {
      mat_pass(c, RowSrc{ c.Hs, nullptr, 0 }, 0, n, v_x, v_t1, nullptr, nullptr, nullptr, 1.0);
    } else {
      _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) v_t1[j] = (c.hess == PQP_HESSIAN_DIAGONAL) ? c.Hs[(size_t)j * n + j] * v_x[j] : 0.0;
      __syncthreads();
    }
  }
  mat_pass(c, RowSrc{ c.As, nullptr, 0 }, 0, ne, primal ? v_x : nullptr, v_se, dual ? v_y : nullptr, v_t2, nullptr, 1.0);
  mat_pass(c, RowSrc{ c.Cs, nullptr, 0 }, 0, ni, primal ? v_x : nullptr, v_rup, dual ? v_z : nullptr, v_t3, nullptr, 1.0);
}

__device__ __noinline__ void global_primal_residual(Ctx& c, const Scal& sc, const pqp_settings& S, Glob& g)
{
  PQP_VECS(c);
  const int n = c.n, ne = c.ne, ni = c.ni, nc = c.nc;
  double mx[5] = { 0, 0, 0, 0, 0 }; // eq_rhs0, in_rhs0, eq_lhs, in_lhs, |x| stuff
  double dummy[1] = { 0 };
  const double* de = v_delta + n;
  const double* di = v_delta + n + ne;
  const double* db = v_delta + n + ne + ni;
  _Pragma("unroll 1") for (int i = threadIdx.x; i < ne; i += NT) {
    double v = v_se[i] / de[i];
    mx[0] = nanmax(mx[0], fabs(v));
    v -= v_b[i];
    mx[2] = nanmax(mx[2], fabs(v));
    v_se[i] = v; // unscaled Ax - b, rescaled below
  }
  _Pragma("unroll 1") for (int i = threadIdx.x; i < nc; i += NT) {
    double v;
    if (i < ni) {
      v = v_rup[i] / di[i];
      mx[1] = nanmax(mx[1], fabs(v));
    } else {
      v = v_x[i - ni] * v_delta[i - ni]; // unscale_primal
    }
    v_rup[i] = v;
    double sv = fmax(v - v_u[i], 0.0) + fmin(v - v_l[i], 0.0);
    v_si[i] = sv;
    mx[3] = nanmax(mx[3], fabs(sv));
    if (i >= ni) {
      // quirk kept: active_part_z.tail = x(scaled) - si ; rhs_0 also takes |x| (scaled), utils.hpp:225-231
      mx[1] = nanmax(mx[1], fabs(v_x[i - ni] - sv));
      mx[1] = nanmax(mx[1], fabs(v_x[i - ni]));
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
    // (v_ex is free between Newton steps; t1..t3 may hold H x, A^T y, C^T z)
    mat_pass(c, RowSrc{ c.Am, nullptr, 0 }, 0, ne, nullptr, nullptr, v_se, v_ex, nullptr, 1.0);
    mat_pass(c, RowSrc{ c.Cm, nullptr, 0 }, 0, ni, nullptr, nullptr, v_si, v_ex, v_ex, 1.0);
    double m = 0;
    _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) m = nanmax(m, fabs(v_ex[j]));
    g.pri_lhs = block_max1(c, m);
  }
  _Pragma("unroll 1") for (int i = threadIdx.x; i < ne; i += NT) v_se[i] *= de[i];
  __syncthreads();
}

// utils.hpp:439-587
__device__ __noinline__ void global_dual_residual(Ctx& c, const Scal& sc, Glob& g)
{
  PQP_VECS(c);
  const int n = c.n, ne = c.ne, ni = c.ni, nc = c.nc;
  const double cs = c.c_scale;
  // t1 = H x, t2 = A^T y, t3 = C^T z_C were produced by global_passes()
  double sm[6] = { 0, 0, 0, 0, 0, 0 }; // g.x, xHx, b.y, zu, zl, (unused)
  double mx[4] = { 0, 0, 0, 0 };       // rhs0, rhs1, rhs3, lhs
  const double inf_b = 1.3407807929942596e+154; // sqrt(DBL_MAX), helpers/common.hpp:20-24
  _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) {
    const double dxc = v_delta[j] * cs;
    double hx = v_t1[j], aty = v_t2[j], ctz = v_t3[j];
    double zb = c.box ? v_z[ni + j] * v_is[j] : 0.0;
    double dr = v_gs[j] + hx + aty + ctz + zb;
    v_dual[j] = dr;
    const double hxu = hx / dxc;
    mx[0] = nanmax(mx[0], fabs(hxu));
    mx[1] = nanmax(mx[1], fabs(aty / dxc));
    mx[2] = nanmax(mx[2], fabs(ctz / dxc));
    if (c.box) mx[2] = nanmax(mx[2], fabs(zb / dxc));
    mx[3] = nanmax(mx[3], fabs(dr / dxc));
    const double xu = v_x[j] * v_delta[j];
    sm[0] += (v_gs[j] / dxc) * xu; // model.g = gs / (delta c)
    sm[1] += hxu * xu;
  }
  const double* de = v_delta + n;
  const double* di = v_delta + n + ne;
  _Pragma("unroll 1") for (int i = threadIdx.x; i < ne; i += NT) sm[2] += v_b[i] * (v_y[i] * de[i] / cs);
  _Pragma("unroll 1") for (int i = threadIdx.x; i < nc; i += NT) {
    double zu_ = v_z[i] * di[i] / cs; // delta laid out [x | eq | in | box]: di[i] covers box too
    if (c.act_up[i]) sm[3] += zu_ * fmin(v_u[i], inf_b);
    if (c.act_low[i]) sm[4] += zu_ * fmax(v_l[i], -inf_b);
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

__device__ __noinline__ LsBase ls_base(const Ctx& c, const Scal& sc, const pqp_settings& S)
{
  PQP_VECS(c);
  const int n = c.n, ne = c.ne, nc = c.nc;
  const bool gpdal = S.merit_function_type == PQP_MERIT_GPDAL;
  double sm[10] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
  double dummy[1] = { 0 };
  _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) {
    double dxj = v_dx[j];
    sm[0] += dxj * v_hdx[j];
    sm[1] += dxj * dxj;
    sm[2] += v_x[j] * v_hdx[j];
    sm[3] += (sc.rho * (v_x[j] - v_xp[j]) + v_gs[j]) * dxj;
  }
  _Pragma("unroll 1") for (int i = threadIdx.x; i < ne; i += NT) {
    double ad = v_adx[i];
    double e = ad - v_ds[i] * sc.mu_eq;
    sm[4] += ad * ad;
    sm[5] += e * e;
    sm[6] += ad * (v_se[i] + v_y[i] * sc.mu_eq);
    sm[7] += e * v_se[i];
  }
  if (gpdal) {
    _Pragma("unroll 1") for (int i = threadIdx.x; i < nc; i += NT) {
      sm[8] += v_dz[i] * v_dz[i];
      sm[9] += v_dz[i] * v_z[i];
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
  PQP_VECS(c);
  const bool gpdal = S.merit_function_type == PQP_MERIT_GPDAL;
  double sq = 0, dt = 0, sq2 = 0, dt2 = 0;
  for (int i = 0; i < c.nc; ++i) {
    const double cd = v_cdx[i], ru = v_rup[i], sl = v_si[i];
    const bool up = (ru + cd * alpha) > 0.0;
    const bool low = (sl + cd * alpha) < 0.0;
    const double cact = (up || low) ? cd : 0.0;
    const double apz = (up ? ru : 0.0) + (low ? sl : 0.0);
    sq += cact * cact;
    dt += apz * cact;
    if (!gpdal) {
      const double e = cact - v_dz[i] * sc.mu_in;
      const double f = apz - v_z[i] * sc.mu_in;
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
__device__ __noinline__ double primal_dual_ls(Ctx& c, const Scal& sc, const pqp_settings& S)
{
  PQP_VECS(c);
  const double eps = 2.220446049250313e-16;
  const int nc = c.nc;
  LsBase base = ls_base(c, sc, S);
  if (threadIdx.x == 0) c.iscratch[2 * NW] = 0;
  __syncthreads();
  _Pragma("unroll 1") for (int i = threadIdx.x; i < nc; i += NT) {
    const double cd = v_cdx[i];
    if (cd != 0.0) {
      double a1 = -v_rup[i] / (cd + eps);
      if (a1 > eps) v_alphas[atomicAdd(&c.iscratch[2 * NW], 1)] = a1;
      double a2 = -v_si[i] / (cd + eps);
      if (a2 > eps) v_alphas[atomicAdd(&c.iscratch[2 * NW], 1)] = a2;
    }
  }
  __syncthreads();
  const int n_alpha = c.iscratch[2 * NW];
  // thread 0 of the last warp additionally evaluates alpha = 0
  double best_pos_alpha = INFINITY, best_pos_grad = 0, best_neg_alpha = 0, best_neg_grad = 0;
  _Pragma("unroll 1") for (int k = threadIdx.x; k < n_alpha + 1; k += NT) {
    const double al = (k < n_alpha) ? v_alphas[k] : 0.0;
    double a, b;
    ls_eval(c, sc, S, base, al, a, b);
    const double gr = a * al + b;
    if (k == n_alpha) {
      v_grads[0] = a;
      v_grads[1] = b; // phi'(0) pieces
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
    v_red[warp * 4 + 0] = best_pos_alpha;
    v_red[warp * 4 + 1] = best_pos_grad;
    v_red[warp * 4 + 2] = best_neg_alpha;
    v_red[warp * 4 + 3] = best_neg_grad;
  }
  __syncthreads();
  double alpha_first_pos = INFINITY, first_pos_grad = 0, alpha_last_neg = 0, last_neg_grad = 0;
  for (int w = 0; w < NW; ++w) {
    if (v_red[w * 4 + 0] < alpha_first_pos) {
      alpha_first_pos = v_red[w * 4 + 0];
      first_pos_grad = v_red[w * 4 + 1];
    }
    if (v_red[w * 4 + 2] > alpha_last_neg) {
      alpha_last_neg = v_red[w * 4 + 2];
      last_neg_grad = v_red[w * 4 + 3];
    }
  }
  const double a0 = v_grads[0], b0 = v_grads[1];
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
#ifdef PQP_CPU_EMU
  return emu::globaltimer();
#else
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
#endif
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
  PQP_VECS(c);
  const PqpQpParams& prm = A.p.params[q];
  const pqp_settings& S = prm.s;
  const int n = c.n, ne = c.ne, ni = c.ni, nc = c.nc;
  const int tid = threadIdx.x;
  int dbg_pos = 0;
  const long long t_qp = PROF_T0();
  long long tph = t_qp;

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
      _Pragma("unroll 1") for (int i = tid; i < ne * n; i += NT) c.As[i] = Asg[i];
    }
    _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_gs[j] = P.gs[(size_t)q * n + j];
    _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) {
      v_bs[j] = P.bs[(size_t)q * ne + j];
      v_b[j] = P.b[(size_t)q * ne + j];
    }
    _Pragma("unroll 1") for (int j = tid; j < nc; j += NT) {
      v_us[j] = P.us[(size_t)q * nc + j];
      v_ls[j] = P.ls[(size_t)q * nc + j];
      if (j < ni) {
        v_u[j] = P.u[(size_t)q * ni + j];
        v_l[j] = P.l[(size_t)q * ni + j];
      } else {
        v_u[j] = P.u_box[(size_t)q * n + j - ni];
        v_l[j] = P.l_box[(size_t)q * n + j - ni];
      }
      c.cons_slot[j] = -1;
      c.act_up[j] = 0;
      c.act_low[j] = 0;
    }
    if (c.box) {
      _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_is[j] = P.is[(size_t)q * n + j];
    }
    _Pragma("unroll 1") for (int j = tid; j < n + ne + nc; j += NT) v_delta[j] = P.delta[(size_t)q * (n + ne + nc) + j];
    if (tid == 0) {
      c.c_scale = P.c[q];
      c.ns = 0;
      c.overflow = 0;
    }
    __syncthreads();
  }
  const double cs = c.c_scale;
  const double* dlx = v_delta;
  const double* dle = v_delta + n;
  const double* dli = v_delta + n + ne; // covers box entries too ([in | box] contiguous)

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
    _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_x[j] = A.p.x[(size_t)q * n + j] / dlx[j];
    _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) v_y[j] = A.p.y[(size_t)q * ne + j] / dle[j] * cs;
    _Pragma("unroll 1") for (int j = tid; j < nc; j += NT) v_z[j] = A.p.z[(size_t)q * nc + j] / dli[j] * cs;
  } else {
    _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_x[j] = 0;
    _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) v_y[j] = 0;
    _Pragma("unroll 1") for (int j = tid; j < nc; j += NT) v_z[j] = 0;
  }
  _Pragma("unroll 1") for (int j = tid; j < n; j += NT) {
    v_rx[j] = 0;
    v_dx[j] = 0;
  }
  _Pragma("unroll 1") for (int j = tid; j < c.cap; j += NT) {
    v_rs[j] = 0;
    v_ds[j] = 0;
  }
  _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) v_se[j] = 0;
  _Pragma("unroll 1") for (int j = tid; j < nc; j += NT) {
    v_si[j] = 0;
    v_dz[j] = 0;
  }
  __syncthreads();

  PROF_ADD(PH_STAGE, tph);
  // ---- first factorisation (helpers.hpp:241-285) -----------------------------
  tph = PROF_T0();
  build_Pi(c, sc.rho);
  PROF_ADD(PH_M1, tph);
  tph = PROF_T0();
  build_dual_block(c, ne, sc.mu_eq, sc.mu_in);
  PROF_ADD(PH_EQ, tph);

  if (prm.start_mode == PQP_START_EQ_GUESS) {
    // helpers.hpp:201-228
    _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_rx[j] = -v_gs[j];
    _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) v_rs[j] = v_bs[j];
    __syncthreads();
    iterative_solve(c, sc, S, 1.0);
    _Pragma("unroll 1") for (int j = tid; j < n; j += NT) {
      v_x[j] = v_dx[j];
      v_dx[j] = 0;
    }
    _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) {
      v_y[j] = v_ds[j];
      v_ds[j] = 0;
    }
    __syncthreads();
  } else if (prm.start_mode == PQP_START_WARM || prm.start_mode == PQP_START_WARM_KEEP) {
    // active set := { i : z_i != 0 } (solver.hpp:1300-1309)
    _Pragma("unroll 1") for (int i = tid; i < nc; i += NT) {
      c.act_up[i] = (v_z[i] != 0.0);
      c.act_low[i] = 0;
    }
    __syncthreads();
    active_set_change(c, sc);
    _Pragma("unroll 1") for (int i = tid; i < nc; i += NT) c.act_up[i] = 0;
    __syncthreads();
  }
  const bool overflow_at_start = c.overflow != 0;

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
    _Pragma("unroll 1") for (int j = tid; j < n; j += NT) m = nanmax(m, fabs(v_gs[j] / (dlx[j] * cs)));
    return block_max1(c, m);
  }(); // |model.g|_inf (helpers.hpp:651)
  double info_pri = 0, info_dua = 0, info_gap = 0;
  bool infeasible_exit = false;
  bool expired = false; // watchdog (debug aid, off by default)
  const unsigned long long t_start = A.watchdog_ns ? gtimer_ns() : 0ull;

  bool residuals_fresh = false;
  for (long long iter = 0; iter < S.max_iter && !overflow_at_start; ++iter) {
    // The reference recomputes both global residuals here; from the second
    // outer iteration on they were already evaluated for exactly this
    // (x, y, z) at the end of the previous iteration, so the values are reused
    // (identical numbers, one streaming pass over H, A, C saved).
    tph = PROF_T0();
    if (!residuals_fresh) {
      global_passes(c, true, true);
      global_primal_residual(c, sc, S, g);
      global_dual_residual(c, sc, g);
    }
    PROF_ADD(PH_GLOBAL, tph);
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
    _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_xp[j] = v_x[j];
    _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) v_yp[j] = v_y[j];
    const double ag = (S.merit_function_type == PQP_MERIT_GPDAL) ? S.alpha_gpdal : 1.0;
    _Pragma("unroll 1") for (int i = tid; i < nc; i += NT) {
      const double zi = v_z[i];
      v_zp[i] = zi;
      double v = v_rup[i] * dli[i]; // scaled C x (box: scaled x-bound residual)
      v += zi * sc.mu_in;
      if (S.merit_function_type == PQP_MERIT_GPDAL) v += (S.alpha_gpdal - 1.0) * sc.mu_in * zi;
      v_rup[i] = v - v_us[i];
      v_si[i] = v - v_ls[i];
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
        _Pragma("unroll 1") for (int i = tid; i < nc; i += NT) {
          c.act_up[i] = v_rup[i] >= 0.0;
          c.act_low[i] = v_si[i] <= 0.0;
        }
        __syncthreads();
        active_set_change(c, sc);
        if (c.overflow) { // S^-1 capacity exceeded: the QP is re-solved by the generic kernel
          expired = true;
          break;
        }
        // q = sum over inactive constraints with z_i != 0 of z_i c_i
        int nq = block_compact(c, nc, c.list2, [&](int i) { return c.cons_slot[i] < 0 && v_z[i] != 0.0; });
        if (nq > 0) {
          mat_pass(c, RowSrc{ nullptr, c.list2, 2 }, 0, nq, nullptr, nullptr, v_z, v_q, nullptr, 1.0);
        } else {
          _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_q[j] = 0;
        }
        _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_rx[j] = -v_dual[j] + v_q[j];
        _Pragma("unroll 1") for (int s = tid; s < c.ns; s += NT) {
          if (s < ne) {
            v_rs[s] = -v_se[s];
          } else {
            const int i = c.slot_cons[s];
            if (c.act_up[i])
              v_rs[s] = -v_rup[i] + v_z[i] * sc.mu_in * ag;
            else
              v_rs[s] = -v_si[i] + v_z[i] * sc.mu_in * ag;
          }
        }
        __syncthreads();
        iterative_solve(c, sc, S, eps_int);
        // un-permute dz; Cdx, CTdz (solver.hpp:860-967)
        _Pragma("unroll 1") for (int i = tid; i < nc; i += NT) {
          const int s = c.cons_slot[i];
          const double dzi = (s >= 0) ? v_ds[s] : -v_z[i];
          v_dz[i] = dzi;
          if (S.merit_function_type == PQP_MERIT_GPDAL) v_cdx[i] += (S.alpha_gpdal - 1.0) * sc.mu_in * dzi;
        }
        _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_ctdz[j] -= v_q[j];
        __syncthreads();
        double alpha = 1.0;
        tph = PROF_T0();
        if (ni > 0 || c.box) alpha = primal_dual_ls(c, sc, S);
        PROF_ADD(PH_LS, tph);
        // |alpha dw|_inf
        {
          double m = 0;
          _Pragma("unroll 1") for (int j = tid; j < n; j += NT) m = nanmax(m, fabs(v_dx[j]));
          _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) m = nanmax(m, fabs(v_ds[j]));
          _Pragma("unroll 1") for (int i = tid; i < nc; i += NT) m = nanmax(m, fabs(v_dz[i]));
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
        _Pragma("unroll 1") for (int j = tid; j < n; j += NT) {
          const double dxj = v_dx[j];
          v_x[j] += alpha * dxj;
          double dr = v_dual[j] + alpha * (sc.rho * dxj + ((c.hess == PQP_HESSIAN_ZERO) ? 0.0 : v_hdx[j]) + v_atdy[j] + v_ctdz[j]);
          v_dual[j] = dr;
          mx[0] = nanmax(mx[0], fabs(dr));
          const double dxc = dlx[j] * cs;
          mx[3] = nanmax(mx[3], fabs(v_atdy[j] / dxc + v_ctdz[j] / dxc));
          const double dxu = dxj * dlx[j];
          mx[5] = nanmax(mx[5], fabs(dxu));
          mx[7] = nanmax(mx[7], fabs(v_hdx[j] / dxc));
          sm[1] += dxj * v_gs[j];
        }
        _Pragma("unroll 1") for (int i = tid; i < ne; i += NT) {
          const double dyi = v_ds[i];
          double sev = v_se[i] + alpha * (v_adx[i] - sc.mu_eq * dyi);
          v_se[i] = sev;
          v_y[i] += alpha * dyi;
          mx[0] = nanmax(mx[0], fabs(sev));
          mx[4] = nanmax(mx[4], fabs(dyi));
          sm[0] += dyi * v_bs[i];
          mx[1] = nanmax(mx[1], fabs(dyi * dle[i] / cs));
          mx[6] = nanmax(mx[6], fabs(v_adx[i] / dle[i]));
        }
        _Pragma("unroll 1") for (int i = tid; i < nc; i += NT) {
          const double dzi = v_dz[i], cd = v_cdx[i];
          const double ru = v_rup[i] + alpha * cd;
          const double sl = v_si[i] + alpha * cd;
          const double zi = v_z[i] + alpha * dzi;
          v_rup[i] = ru;
          v_si[i] = sl;
          v_z[i] = zi;
          const double apz = fmax(ru, 0.0) + fmin(sl, 0.0) - ag * zi * sc.mu_in;
          mx[0] = nanmax(mx[0], fabs(apz));
          mx[4] = nanmax(mx[4], fabs(dzi));
          sm[0] += fmax(dzi, 0.0) * v_us[i] - fmin(dzi, 0.0) * v_ls[i];
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
            _Pragma("unroll 1") for (int i = tid; i < nc; i += NT) {
              const double v = v_cdx[i] / dli[i]; // unscaled (box entries use delta_box)
              bool ok = true;
              if (v_us[i] <= 1e20 && v_ls[i] >= -1e20)
                ok = v <= bound && v >= -bound;
              else if (v_us[i] > 1e20)
                ok = v >= -bound;
              else if (v_ls[i] < -1e20)
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
      _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_x[j] = v_dx[j] * dlx[j];
      _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) v_y[j] = v_ds[j] * dle[j] / cs;
      _Pragma("unroll 1") for (int i = tid; i < nc; i += NT) v_z[i] = v_dz[i] * dli[i] / cs;
      __syncthreads();
      infeasible_exit = true;
      break;
    }
    if (scaled_eps == S.eps_abs && S.primal_infeasibility_solving && sc.status == PQP_PRIMAL_INFEASIBLE) {
      // solver.hpp:1581-1595
      _Pragma("unroll 1") for (int j = tid; j < c.cap; j += NT) v_s1[j] = 1.0;
      __syncthreads();
      mat_pass(c, RowSrc{ c.Am, nullptr, 0 }, 0, ne, nullptr, nullptr, v_s1, v_t1, nullptr, 1.0);
      mat_pass(c, RowSrc{ c.Cm, nullptr, 0 }, 0, ni, nullptr, nullptr, v_s1, v_t1, v_t1, 1.0);
      double m = 0;
      _Pragma("unroll 1") for (int j = tid; j < n; j += NT) m = nanmax(m, fabs(v_t1[j] + (c.box ? v_is[j] : 0.0)));
      scaled_eps = block_max1(c, m) * S.eps_abs;
    }
    tph = PROF_T0();
    global_passes(c, true, false);
    global_primal_residual(c, sc, S, g);
    PROF_ADD(PH_GLOBAL, tph);
    bool dual_done = false; // dual residual already evaluated for the current (x, y, z)
    double primal_feasibility_lhs_new = g.pri_lhs;
    is_primal_feasible = primal_feasibility_lhs_new <= (scaled_eps + S.eps_rel * fmax(g.pri_eq_rhs0, g.pri_in_rhs0));
    info_pri = primal_feasibility_lhs_new;
    if (is_primal_feasible) {
      tph = PROF_T0();
      global_passes(c, false, true);
      global_dual_residual(c, sc, g);
      PROF_ADD(PH_GLOBAL, tph);
      dual_done = true;
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
        _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) v_y[j] = v_yp[j];
        _Pragma("unroll 1") for (int i = tid; i < nc; i += NT) v_z[i] = v_zp[i];
        __syncthreads();
        dual_done = false;
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
    tph = PROF_T0();
    if (!dual_done) {
      global_passes(c, false, true);
      global_dual_residual(c, sc, g);
    }
    PROF_ADD(PH_GLOBAL, tph);
    // (not in closest-feasible mode: there the primal residual depends on info.status, which the end of this
    // iteration may just have changed from PRIMAL_INFEASIBLE to SOLVED_CLOSEST_PRIMAL_FEASIBLE, utils.hpp:241-248)
    residuals_fresh = !S.primal_infeasibility_solving;
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
        tph = PROF_T0();
        rebuild_Si_from_G(c, new_mu_eq, new_mu_in);
        PROF_ADD(PH_MU, tph);
        sc.factor_fresh = false;
      }
    }
    sc.mu_eq = new_mu_eq;
    sc.mu_in = new_mu_in;
    sc.mu_eq_inv = new_mu_eq_inv;
    sc.mu_in_inv = new_mu_in_inv;
  }

  if (c.overflow) {
    // leave x, y, z untouched (warm starts must see the caller's values again)
    if (tid == 0) A.p.info[(size_t)q * PQP_INFO_DOUBLES + 10] = 99.0; // internal: retry with the generic kernel
    __syncthreads();
    return;
  }
  // ---- unscale and write back (solver.hpp:1749-1836) -------------------------
  double* xo = A.p.x + (size_t)q * n;
  double* yo = A.p.y + (size_t)q * ne;
  double* zo = A.p.z + (size_t)q * nc;
  double* seo = A.p.se + (size_t)q * ne;
  double* sio = A.p.si + (size_t)q * nc;
  const bool unscale_s = S.primal_infeasibility_solving && sc.status == PQP_PRIMAL_INFEASIBLE;
  _Pragma("unroll 1") for (int j = tid; j < n; j += NT) {
    const double xu = v_x[j] * dlx[j];
    v_t1[j] = xu;
    xo[j] = xu;
  }
  _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) {
    yo[j] = v_y[j] * dle[j] / cs;
    seo[j] = unscale_s ? v_se[j] / dle[j] : v_se[j];
  }
  _Pragma("unroll 1") for (int i = tid; i < nc; i += NT) {
    zo[i] = v_z[i] * dli[i] / cs;
    sio[i] = unscale_s ? v_si[i] / dli[i] : v_si[i];
  }
  __syncthreads();
  (void)infeasible_exit;
  // objective 0.5 x^T H x + g^T x from the model (solver.hpp:1769-1781)
  double obj;
  {
    if (c.hess == PQP_HESSIAN_DENSE) {
      mat_pass(c, RowSrc{ c.Hm, nullptr, 0 }, 0, n, v_t1, v_t2, nullptr, nullptr, nullptr, 1.0);
    } else {
      _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_t2[j] = c.Hm[(size_t)j * n + j] * v_t1[j];
      __syncthreads();
    }
    double part = 0;
    const double* gm = A.p.g + (size_t)q * n;
    _Pragma("unroll 1") for (int j = tid; j < n; j += NT) part += v_t1[j] * (0.5 * v_t2[j] + gm[j]);
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
  PROF_ADD(PH_TOTAL, t_qp);
  __syncthreads();
}

#ifdef PQP_WITH_BACKWARD
// ---------------------------------------------------------------------------
// QPLayer backward of one solved QP: dense/compute_ECJ.hpp:29-190
// (compute_backward, compute_backward_loss_ESG) with backward_data.hpp:27-129.
// One more solve of the regularised KKT system (rho_new, mu_new) with the
// active set AT THE SOLUTION, right-hand side = - scaled loss derivative, then
// the outer products that are the Jacobian-vector products w.r.t. H, g, A, b,
// C, u, l. Same block elimination as the forward path: P^-1 for rho_new, the
// dual block for (equalities + active rows) with mu_new, iterative refinement.
// Deviation (documented in DESIGN.md): a non-zero dL/dz of an ACTIVE row is
// scaled once by that row's Ruiz factor; the reference rescales the whole
// block inside its loop over the rows (compute_ECJ.hpp:105-117) — identical
// for dL/dz = 0, which is what the QP layer passes unless the loss depends on
// the multipliers.
// ---------------------------------------------------------------------------
__device__ void backward_one(Ctx& c, const PqpSolveArgs& A, const PqpBackwardArgs& K, int q)
{
  PQP_VECS(c);
  const PqpQpParams& prm = A.p.params[q];
  const pqp_settings& S = prm.s;
  const int n = c.n, ne = c.ne, ni = c.ni, nc = c.nc; // no box constraints on this path: nc == ni
  const int tid = threadIdx.x;
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
      _Pragma("unroll 1") for (int i = tid; i < ne * n; i += NT) c.As[i] = Asg[i];
    }
    _Pragma("unroll 1") for (int j = tid; j < nc; j += NT) {
      v_u[j] = P.u[(size_t)q * ni + j];
      v_l[j] = P.l[(size_t)q * ni + j];
      c.cons_slot[j] = -1;
      c.act_up[j] = 0;
      c.act_low[j] = 0;
    }
    _Pragma("unroll 1") for (int j = tid; j < n + ne + nc; j += NT) v_delta[j] = P.delta[(size_t)q * (n + ne + nc) + j];
    // the solution, in the model's units
    _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_x[j] = P.x[(size_t)q * n + j];
    _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) v_y[j] = P.y[(size_t)q * ne + j];
    _Pragma("unroll 1") for (int j = tid; j < nc; j += NT) v_z[j] = P.z[(size_t)q * nc + j];
    _Pragma("unroll 1") for (int j = tid; j < n; j += NT) {
      v_rx[j] = 0;
      v_dx[j] = 0;
    }
    _Pragma("unroll 1") for (int j = tid; j < c.cap; j += NT) {
      v_rs[j] = 0;
      v_ds[j] = 0;
    }
    _Pragma("unroll 1") for (int j = tid; j < nc; j += NT) v_dz[j] = 0;
    if (tid == 0) {
      c.c_scale = P.c[q];
      c.ns = 0;
      c.overflow = 0;
    }
    __syncthreads();
  }
  const double cs = c.c_scale;
  const double* dlx = v_delta;
  const double* dle = v_delta + n;
  const double* dli = v_delta + n + ne;

  Scal sc;
  sc.rho = K.rho_new; // compute_ECJ.hpp:66-68
  sc.mu_eq = K.mu_new;
  sc.mu_in = K.mu_new;
  sc.mu_eq_inv = 1.0 / sc.mu_eq;
  sc.mu_in_inv = 1.0 / sc.mu_in;
  sc.nu = 1.0;
  sc.iter = 0;
  sc.iter_ext = 0;
  sc.mu_updates = 0;
  sc.status = PQP_SOLVED;
  sc.iterative_residual = 0;
  sc.factor_fresh = true;

  // active set at the solution (compute_ECJ.hpp:52-61): C x + z against u and l, unscaled
  if (ni > 0) mat_pass(c, RowSrc{ c.Cm, nullptr, 0 }, 0, ni, v_x, v_cdx, nullptr, nullptr, nullptr, 1.0);
  __syncthreads();
  _Pragma("unroll 1") for (int i = tid; i < ni; i += NT) {
    const double ctz = v_cdx[i] + v_z[i];
    c.act_up[i] = (ctz - v_u[i]) >= 0.0;
    c.act_low[i] = (ctz - v_l[i]) <= 0.0;
  }
  __syncthreads();
  // factorisation from scratch with the new proximal parameters, whole active set inserted (:74-91)
  build_Pi(c, sc.rho);
  build_dual_block(c, ne, sc.mu_eq, sc.mu_in);
  active_set_change(c, sc);
  sc.factor_fresh = true; // constraints_changed = false: no refactorisation inside the refinement
  // rhs = - loss derivative, scaled block by block (:93-118)
  const double* ld = K.loss_derivative + (size_t)q * (n + ne + ni);
  _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_rx[j] = (-ld[j]) * (dlx[j] * cs);
  _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) v_rs[j] = (-ld[n + j]) * dle[j];
  _Pragma("unroll 1") for (int i = tid; i < ni; i += NT) {
    const int s = c.cons_slot[i];
    if (s >= 0) v_rs[s] = (-ld[n + ne + i]) * dli[i];
  }
  __syncthreads();
  iterative_solve(c, sc, S, K.eps);
  // unpermute, unscale (compute_ECJ.hpp:131-154): dx -> t1, dy -> s1, dz -> dz
  _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_t1[j] = v_dx[j] * dlx[j];
  _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) v_s1[j] = v_ds[j] * dle[j] / cs;
  _Pragma("unroll 1") for (int i = tid; i < ni; i += NT) {
    const int s = c.cons_slot[i];
    const double raw = (s >= 0) ? v_ds[s] : ld[n + ne + i];
    v_dz[i] = raw * dli[i] / cs;
  }
  __syncthreads();
  // jacobian-vector products (compute_ECJ.hpp:156-187), one warp per output row
  const int lane = tid & 31, warp = tid >> 5;
  double* oH = K.dL_dH + (size_t)q * n * n;
  double* oA = K.dL_dA + (size_t)q * ne * n;
  double* oC = K.dL_dC + (size_t)q * ni * n;
  _Pragma("unroll 1") for (int i = warp; i < n; i += NW) {
    const double dxi = v_t1[i], xi = v_x[i];
    _Pragma("unroll 1") for (int j = lane; j < n; j += 32) oH[(size_t)i * n + j] = 0.5 * (dxi * v_x[j] + xi * v_t1[j]);
  }
  _Pragma("unroll 1") for (int i = warp; i < ne; i += NW) {
    const double dyi = v_s1[i], yi = v_y[i];
    _Pragma("unroll 1") for (int j = lane; j < n; j += 32) oA[(size_t)i * n + j] = dyi * v_x[j] + yi * v_t1[j];
  }
  _Pragma("unroll 1") for (int i = warp; i < ni; i += NW) {
    const double dzi = v_dz[i], zi = v_z[i];
    _Pragma("unroll 1") for (int j = lane; j < n; j += 32) oC[(size_t)i * n + j] = dzi * v_x[j] + zi * v_t1[j];
  }
  _Pragma("unroll 1") for (int j = tid; j < n; j += NT) K.dL_dg[(size_t)q * n + j] = v_t1[j];
  _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) K.dL_db[(size_t)q * ne + j] = -v_s1[j];
  _Pragma("unroll 1") for (int i = tid; i < ni; i += NT) {
    K.dL_du[(size_t)q * ni + i] = c.act_up[i] ? -v_dz[i] : 0.0;
    K.dL_dl[(size_t)q * ni + i] = c.act_low[i] ? -v_dz[i] : 0.0;
  }
  if (tid == 0) A.p.info[(size_t)q * PQP_INFO_DOUBLES + 18] = sc.iterative_residual;
  __syncthreads();
}
#endif // PQP_WITH_BACKWARD

#ifdef PQP_CPU_EMU
static double* const smem_dyn = emu::dyn_smem;
#else
extern __shared__ __align__(16) double smem_dyn[];
#endif

// FUSED == 1: the feed gate + set-up of the end-to-end path are compiled in (a separate instantiation keeps the
// register allocation of the plain solve kernel untouched); FUSED == 2: the QPLayer backward pass per QP
template<int FUSED>
__device__ __forceinline__ void solve_kernel_body(const PqpSolveArgs& A, const PqpBackwardArgs* K = nullptr)
{
  __shared__ Ctx c;
  __shared__ int cur_q;
  __shared__ setupk::FeedArgs feed_args;
  if (FUSED == 1 && threadIdx.x == 0) {
    feed_args.d = A.d;
    feed_args.p = A.p;
    feed_args.ready = A.ready;
    feed_args.batch = A.batch;
    feed_args.fused_setup = A.fused_setup;
    feed_args.feed_margin = A.feed_margin;
  }
  __shared__ long long prof_sh[PH_COUNT];
  const PqpLayout& L = A.lay;
  if (threadIdx.x == 0) {
    for (int k = 0; k < PH_COUNT; ++k) prof_sh[k] = 0;
    c.prof = A.prof ? prof_sh : nullptr;
    c.vec_smem = L.in_smem[PA_VEC];
    c.pi_smem = L.in_smem[PA_M1];
    c.si_cap = L.si_cap;
    c.uv_ld = ((A.d.n > A.d.cap ? A.d.n : A.d.cap) + 2) & ~1;
    c.overflow = 0;
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
    c.Pi = place(PA_M1);
    c.As = place(PA_AS);
    c.Si = place(PA_MS);
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
    const int cq = cur_q;
    const int q = A.first + cq; // this launch owns the QPs [first, first + batch)
    __syncthreads();
    if (cq >= A.batch) break;
    if (FUSED == 1 && (A.ready || A.fused_setup)) {
      if (!setupk::feed_and_setup(&feed_args, cq, q, smem_dyn)) continue;
    }
    if (!A.p.params[q].active) continue;
    if (threadIdx.x == 0) c.As = As_home;
    __syncthreads();
#ifdef PQP_WITH_BACKWARD
    if (FUSED == 2) {
      backward_one(c, A, *K, q);
      continue;
    }
#endif
    solve_one(c, A, q);
  }
  if (A.prof && threadIdx.x == 0) {
    for (int k = 0; k < PH_COUNT; ++k) atomicAdd(reinterpret_cast<unsigned long long*>(A.prof) + k, (unsigned long long)prof_sh[k]);
  }
}

// Plain launch: parameters by value, exactly the kernel the device-resident path has always run.
__global__ void __launch_bounds__(NT, PQP_MIN_CTAS) pqp_solve_kernel(PqpSolveArgs A)
{
  solve_kernel_body<0>(A);
}
// Fused feed: __grid_constant__ lets the non-inlined set-up read the parameters in place.
__global__ void __launch_bounds__(NT, PQP_MIN_CTAS) pqp_solve_kernel_fused(const __grid_constant__ PqpSolveArgs A)
{
  solve_kernel_body<1>(A);
}

#ifdef PQP_WITH_BACKWARD
// QPLayer backward of every active QP of the launch (same persistent work queue as the solve)
__global__ void __launch_bounds__(NT, PQP_MIN_CTAS) pqp_backward_kernel(const __grid_constant__ PqpSolveArgs A, const __grid_constant__ PqpBackwardArgs K)
{
  solve_kernel_body<2>(A, &K);
}
#endif
