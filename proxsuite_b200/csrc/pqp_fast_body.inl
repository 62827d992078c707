// Device code of the TILE kernel: the specialisation of the solve kernel for dense
// Hessians without box constraints, n even and <= 128, dual block <= 128 (the
// headline shapes). Same algorithm and driver as pqp_solver_body.inl; different
// linear-algebra layer: every matrix-vector product is an AXPY-form streaming
// pass (no cross-lane reductions), S^-1 lives in shared memory in 32 x 32 tile
// storage, P^-1 / Bt / G in the per-CTA L2 workspace. See DESIGN.md section 3.


struct Ctx
{
  int n, ne, ni, nc, box, hess, cap;
  int ns; // current size of the dual block: ne + number of active inequalities
  // factor storage
  double *Pi, *As, *Si, *G, *Y;
  double* Bt;      // [A_s; C_s]^T, n x ldb (L2 workspace)
  double* kt;      // ne + ni products of one Bt pass / dense coefficient vector of a B^T product (shared memory)
  double* kt2;     // second dense coefficient vector (global dual residual)
  double* W;       // P^-1 B^T, n x ldb (L2 workspace; only needed to form G)
  int ldb, ldn;    // leading dimensions of Bt and Pi (even)
  int m;           // rows of B = [A_s; C_s; box rows]: ne + nc
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
  int si_valid;    // 0: the dual block has not been formed yet (deferred to the first active-set change)
  int overflow;    // set when an insertion would exceed si_cap (QP is retried by the generic kernel)
  // BIG variant, last-resort fallback: explicit inverse of the WHOLE KKT matrix (see kkt_factor)
  int pf;          // BIG variant: L2 prefetch distance of the streaming passes (warp iterations ahead, 0 = off)
  int kkt_mode;    // 1: the dual-block path stagnated on this QP; solves go through K^-1 (stored where S^-1 was)
  int kkt_dirty;   // K^-1 must be re-formed before the next solve (active set / mu changed)
  double* kws;     // workspace of the fallback (the W region: free once G is built)
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

#ifndef PQP_REDUX_MAX
#define PQP_REDUX_MAX 1
#endif
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
// NaN-propagating maximum over the warp of values with the sign bit CLEAR (every caller reduces |.| values or 0 / 1
// flags): for such doubles the order of the bit patterns is the order of the values, with every NaN above +inf, so the
// maximum is the 64-bit unsigned maximum of the patterns - two redux.sync.max.u32 (high words, then the low words of the
// lanes that hold the winning high word) instead of a five-step shuffle butterfly with two compares and two selects
// per step. Returns the same value as the butterfly of nanmax() (a NaN if any lane holds one, else the largest).
__device__ __forceinline__ double warp_max(double v)
{
#if PQP_REDUX_MAX
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned hmax = __reduce_max_sync(FULL, hi);
  const unsigned lmax = __reduce_max_sync(FULL, (hi == hmax) ? lo : 0u);
  return __hiloint2double((int)hmax, (int)lmax);
#else
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
#endif
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
#if PQP_REDUX_MAX
    maxs[k] = warp_max(v_red[(lane & (NW - 1)) * (KS + KM) + KS + k]); // (every warp value four times over the lanes)
#else
    double m = v_red[KS + k];
#pragma unroll
    for (int w = 1; w < NW; ++w) m = nanmax(m, v_red[w * (KS + KM) + KS + k]);
    maxs[k] = m;
#endif
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

#define RPB (32 / NW) // rows per warp per 32-row tile

// ---------------------------------------------------------------------------
// AXPY-form streaming pass (the only matrix-vector primitive of this kernel):
//   out[j] = add[j] + sign * sum_{k < nrows} coef[ck] * row_k[j],   j < ncols
// row_k = base0 + k*ld for k < split, else base1 + list[k]*ld (list == null:
// base1 + (k - split)*ld); ck = k, or list[k] when byid. Warp w owns rows
// w, w+NW, ...; a lane owns column PAIRS (16-byte loads, no shuffles, no
// per-row predicates); the NW partial vectors are combined through shared
// memory. Every product of the iteration is put in this form by keeping the
// transposed copy Bt = [A_s; C_s]^T next to the row-major matrices and by
// using the symmetry of H_s and P^-1. Rows must be 16-byte aligned (ld even).
// ---------------------------------------------------------------------------
template<int NCH>
__device__ void axpy_pass_t(const Ctx& c, const double* __restrict__ base0, int split, const double* __restrict__ base1, int ld, const int* __restrict__ list, int byid, int nrows, const double* __restrict__ coef, int ncols, double* out, const double* add, double sign, double* raw)
{
  constexpr int UNR = (NCH <= 2) ? 8 : 4; // rows in flight per warp (L2 latency is the bound)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* const scr = c.scratch;
  PQP_SM(scr);
  PQP_SM(coef);
  PQP_SM(out);
  const int np = (ncols + 1) >> 1;
  bool pv[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) pv[ch] = lane + 32 * ch < np;
  double2 acc[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) acc[ch] = make_double2(0.0, 0.0);
  auto rowptr = [&](int k, double& cf) -> const double2* {
    int id = k;
    const double* b = base0;
    if (k >= split) {
      b = base1;
      id = list ? list[k] : k - split;
    }
    cf = coef[(byid && k >= split) ? id : k];
    return reinterpret_cast<const double2*>(b + (size_t)id * (size_t)ld) + lane;
  };
  int k = warp;
  _Pragma("unroll 1") for (; k + (UNR - 1) * NW < nrows; k += UNR * NW) {
    const double2* rp[UNR];
    double cf[UNR];
    double2 v[UNR][NCH];
#pragma unroll
    for (int u = 0; u < UNR; ++u) rp[u] = rowptr(k + u * NW, cf[u]);
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) v[u][ch] = pv[ch] ? rp[u][32 * ch] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        acc[ch].x = fma(cf[u], v[u][ch].x, acc[ch].x);
        acc[ch].y = fma(cf[u], v[u][ch].y, acc[ch].y);
      }
    }
  }
  // remainder (fewer than UNR rows of this warp): ONE more batch with the loads of all its rows in flight together
  // (a row-at-a-time tail costs one L2 round trip per row: up to five in a 100-row pass, against two for the batches)
  if (k < nrows) {
    const double2* rp[UNR];
    double cf[UNR];
    double2 v[UNR][NCH];
    bool rv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      rv[u] = k + u * NW < nrows;
      cf[u] = 0.0;
      rp[u] = nullptr;
      if (rv[u]) rp[u] = rowptr(k + u * NW, cf[u]);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) v[u][ch] = (rv[u] && pv[ch]) ? rp[u][32 * ch] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (rv[u]) { // (skipping keeps the sum bit-identical to the row-at-a-time form: no 0 * x terms are added)
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          acc[ch].x = fma(cf[u], v[u][ch].x, acc[ch].x);
          acc[ch].y = fma(cf[u], v[u][ch].y, acc[ch].y);
        }
      }
    }
  }
  double2* const scr2 = reinterpret_cast<double2*>(scr);
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    if (pv[ch]) scr2[warp * np + lane + 32 * ch] = acc[ch];
  }
  __syncthreads();
  _Pragma("unroll 1") for (int j = threadIdx.x; j < ncols; j += NT) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += scr[w * 2 * np + j];
    out[j] = (add ? add[j] : 0.0) + sign * s;
    if (raw) raw[j] = s; // the plain sum, for callers that need both
  }
  __syncthreads();
}

#ifdef PQP_BIG
// Software prefetch into L2 of `count` doubles at p (global memory): lane l of nl cooperating lanes takes every nl-th
// 128-byte line. The large shapes stream per-CTA workspaces that live in HBM (cfg 5: 27 MB x 296 CTAs); a pass keeps
// at most 128-256 bytes per thread in flight and stalls a full HBM round trip per batch of rows - with the rows of
// the iteration after next requested here, the loads find them in L2.
__device__ __forceinline__ void pf_l2_span(const double* p, int count, int l, int nl)
{
#ifndef PQP_CPU_EMU
  const char* const b = reinterpret_cast<const char*>(p);
  const int bytes = count * 8;
  _Pragma("unroll 1") for (int o = l * 128; o < bytes; o += nl * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(__cvta_generic_to_global(b + o)));
#else
  (void)p; (void)count; (void)l; (void)nl;
#endif
}
// BIG variant (any number of columns): the same pass, columns handled in groups of 256 (register accumulators for
// one group at a time), rows predicated instead of peeled. Partial vectors: c.scratch, NW x 2*np doubles.
__device__ __noinline__ void axpy_pass_big(const Ctx& c, const double* __restrict__ base0, int split, const double* __restrict__ base1, int ld, const int* __restrict__ list, int byid, int nrows, const double* __restrict__ coef, int ncols, double* out, const double* add, double sign, double* raw)
{
  constexpr int NCH = 4, UNR = 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* const scr = c.scratch;
  const int np = (ncols + 1) >> 1;
  double2* const scr2 = reinterpret_cast<double2*>(scr);
  _Pragma("unroll 1") for (int g0 = 0; g0 < np; g0 += 32 * NCH) {
    bool pv[NCH];
    double2 acc[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      pv[ch] = g0 + lane + 32 * ch < np;
      acc[ch] = make_double2(0.0, 0.0);
    }
    _Pragma("unroll 1") for (int k = warp; k < nrows; k += UNR * NW) {
      if (c.pf) { // row u = lane / 8 of the batch `pf` iterations ahead, eight lanes per row
        const int kk = k + c.pf * UNR * NW + (lane >> 3) * NW;
        if (kk < nrows) {
          int id = kk;
          const double* b = base0;
          if (kk >= split) {
            b = base1;
            id = list ? list[kk] : kk - split;
          }
          pf_l2_span(b + (size_t)id * (size_t)ld + 2 * g0, min(256, ncols - 2 * g0), lane & 7, 8);
        }
      }
      const double2* rp[UNR];
      double cf[UNR];
      bool rv[UNR];
      double2 v[UNR][NCH];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int kk = k + u * NW;
        rv[u] = kk < nrows;
        cf[u] = 0.0;
        rp[u] = nullptr;
        if (rv[u]) {
          int id = kk;
          const double* b = base0;
          if (kk >= split) {
            b = base1;
            id = list ? list[kk] : kk - split;
          }
          cf[u] = coef[(byid && kk >= split) ? id : kk];
          rp[u] = reinterpret_cast<const double2*>(b + (size_t)id * (size_t)ld) + g0 + lane;
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) v[u][ch] = (rv[u] && pv[ch]) ? rp[u][32 * ch] : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (rv[u]) {
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            acc[ch].x = fma(cf[u], v[u][ch].x, acc[ch].x);
            acc[ch].y = fma(cf[u], v[u][ch].y, acc[ch].y);
          }
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      if (pv[ch]) scr2[warp * np + g0 + lane + 32 * ch] = acc[ch];
    }
  }
  __syncthreads();
  _Pragma("unroll 1") for (int j = threadIdx.x; j < ncols; j += NT) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += scr[w * 2 * np + j];
    out[j] = (add ? add[j] : 0.0) + sign * s;
    if (raw) raw[j] = s;
  }
  __syncthreads();
}
#endif

__device__ __noinline__ void axpy_pass2(const Ctx& c, const double* base0, int split, const double* base1, int ld, const int* list, int byid, int nrows, const double* coef, int ncols, double* out, const double* add, double sign, double* raw = nullptr)
{
#ifdef PQP_BIG
  axpy_pass_big(c, base0, split, base1, ld, list, byid, nrows, coef, ncols, out, add, sign, raw);
  return;
#endif
  const int np = (ncols + 1) >> 1;
  if (np <= 32)
    axpy_pass_t<1>(c, base0, split, base1, ld, list, byid, nrows, coef, ncols, out, add, sign, raw);
  else if (np <= 64)
    axpy_pass_t<2>(c, base0, split, base1, ld, list, byid, nrows, coef, ncols, out, add, sign, raw);
  else if (np <= 96)
    axpy_pass_t<3>(c, base0, split, base1, ld, list, byid, nrows, coef, ncols, out, add, sign, raw);
  else
    axpy_pass_t<4>(c, base0, split, base1, ld, list, byid, nrows, coef, ncols, out, add, sign, raw);
}
// all rows from one matrix
__device__ __forceinline__ void axpy_pass(const Ctx& c, const double* base, int ld, int nrows, const double* coef, int ncols, double* out, const double* add, double sign)
{
  axpy_pass2(c, base, nrows, base, ld, nullptr, 0, nrows, coef, ncols, out, add, sign);
}

// Products with B^T through the SAME transposed copy Bt (row j of Bt = column j of B = [A_s; C_s]):
//   out1[j] = add[j] + sign * (Bt[j,:] . coef1),  raw1[j] = the plain dot,  out2[j] = Bt[j,:] . coef2
// coef vectors are dense over the ne + ni constraint rows (zero where a row does not take part).
// A warp owns four rows at a time; the four row sums share one transpose-reduction. Using Bt for
// both B v (AXPY form) and B^T lam (this form) keeps A_s / C_s out of the L2 working set of the
// iteration (they are read once, to build Bt).
template<int NCH, bool TWO>
__device__ void bt_dot_t(const Ctx& c, const double* __restrict__ coef1, const double* __restrict__ coef2, double* out1, const double* add, double sign, double* raw1, double* out2)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = c.n, np = c.ldb >> 1;
  PQP_SM(coef1);
  PQP_SM(out1);
  if (TWO) {
    PQP_SM(coef2);
    PQP_SM(out2);
  }
  bool pv[NCH];
  double2 c1[NCH], c2[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    pv[ch] = lane + 32 * ch < np;
    c1[ch] = pv[ch] ? reinterpret_cast<const double2*>(coef1)[lane + 32 * ch] : make_double2(0.0, 0.0);
    c2[ch] = (TWO && pv[ch]) ? reinterpret_cast<const double2*>(coef2)[lane + 32 * ch] : make_double2(0.0, 0.0);
  }
  _Pragma("unroll 1") for (int jb = warp; jb < n; jb += 4 * NW) {
    double2 v[4][NCH];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = jb + u * NW;
      const double2* rp = reinterpret_cast<const double2*>(c.Bt + (size_t)(j < n ? j : 0) * c.ldb) + lane;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) v[u][ch] = (j < n && pv[ch]) ? rp[32 * ch] : make_double2(0.0, 0.0);
    }
    double d1[4], d2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double a1 = 0.0, a2 = 0.0;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        a1 = fma(v[u][ch].x, c1[ch].x, a1);
        a1 = fma(v[u][ch].y, c1[ch].y, a1);
        if (TWO) {
          a2 = fma(v[u][ch].x, c2[ch].x, a2);
          a2 = fma(v[u][ch].y, c2[ch].y, a2);
        }
      }
      d1[u] = a1;
      d2[u] = a2;
    }
    reduce_rows<4>(d1, lane);
    if (TWO) reduce_rows<4>(d2, lane);
    if ((lane & 7) == 0) {
      const int j = jb + (lane >> 3) * NW;
      if (j < n) {
        out1[j] = (add ? add[j] : 0.0) + sign * d1[0];
        if (raw1) raw1[j] = d1[0];
        if (TWO) out2[j] = d2[0];
      }
    }
  }
  __syncthreads();
}
#ifdef PQP_BIG
// BIG variant: any row length; the coefficient vectors are re-read per 256-column group (shared memory) instead of
// being held in registers for the whole pass.
__device__ __noinline__ void bt_dot_big(const Ctx& c, const double* __restrict__ coef1, const double* __restrict__ coef2, double* out1, const double* add, double sign, double* raw1, double* out2)
{
  constexpr int NCH = 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // box constraints: only the [A_s; C_s] columns of Bt are streamed, the box block (i_s[j] e_j) enters as one term per row
  const int n = c.n, nr = c.ne + c.ni, np = c.box ? ((nr + 1) >> 1) : (c.ldb >> 1);
  const bool two = coef2 != nullptr;
  const bool odd_tail = c.box && (nr & 1); // the last streamed pair then holds the first box column: masked
  _Pragma("unroll 1") for (int jb = warp; jb < n; jb += 4 * NW) {
    if (c.pf) {
      const int j = jb + c.pf * 4 * NW + (lane >> 3) * NW;
      if (j < n) pf_l2_span(c.Bt + (size_t)j * c.ldb, 2 * np, lane & 7, 8);
    }
    double d1[4] = { 0.0, 0.0, 0.0, 0.0 }, d2[4] = { 0.0, 0.0, 0.0, 0.0 };
    _Pragma("unroll 1") for (int g0 = 0; g0 < np; g0 += 32 * NCH) {
      bool pv[NCH];
      double2 c1[NCH], c2[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        pv[ch] = g0 + lane + 32 * ch < np;
        c1[ch] = pv[ch] ? reinterpret_cast<const double2*>(coef1)[g0 + lane + 32 * ch] : make_double2(0.0, 0.0);
        c2[ch] = (two && pv[ch]) ? reinterpret_cast<const double2*>(coef2)[g0 + lane + 32 * ch] : make_double2(0.0, 0.0);
      }
      double2 v[4][NCH];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = jb + u * NW;
        const double2* rp = reinterpret_cast<const double2*>(c.Bt + (size_t)(j < n ? j : 0) * c.ldb) + g0 + lane;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          v[u][ch] = (j < n && pv[ch]) ? rp[32 * ch] : make_double2(0.0, 0.0);
          if (odd_tail && g0 + lane + 32 * ch == np - 1) v[u][ch].y = 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          d1[u] = fma(v[u][ch].x, c1[ch].x, d1[u]);
          d1[u] = fma(v[u][ch].y, c1[ch].y, d1[u]);
          if (two) {
            d2[u] = fma(v[u][ch].x, c2[ch].x, d2[u]);
            d2[u] = fma(v[u][ch].y, c2[ch].y, d2[u]);
          }
        }
      }
    }
    reduce_rows<4>(d1, lane);
    if (two) reduce_rows<4>(d2, lane);
    if ((lane & 7) == 0) {
      const int j = jb + (lane >> 3) * NW;
      if (j < n) {
        double r1 = d1[0], r2 = d2[0];
        if (c.box) {
          r1 = fma(c.is[j], coef1[nr + j], r1);
          if (two) r2 = fma(c.is[j], coef2[nr + j], r2);
        }
        out1[j] = (add ? add[j] : 0.0) + sign * r1;
        if (raw1) raw1[j] = r1;
        if (two) out2[j] = r2;
      }
    }
  }
  __syncthreads();
}
#endif

__device__ __noinline__ void bt_dot(const Ctx& c, const double* coef1, const double* coef2, double* out1, const double* add, double sign, double* raw1, double* out2)
{
#ifdef PQP_BIG
  bt_dot_big(c, coef1, coef2, out1, add, sign, raw1, out2);
  return;
#endif
  const int np = c.ldb >> 1;
  if (coef2) {
    if (np <= 64)
      bt_dot_t<2, true>(c, coef1, coef2, out1, add, sign, raw1, out2);
    else
      bt_dot_t<4, true>(c, coef1, coef2, out1, add, sign, raw1, out2);
  } else {
    if (np <= 64)
      bt_dot_t<2, false>(c, coef1, coef2, out1, add, sign, raw1, out2);
    else if (np <= 96)
      bt_dot_t<3, false>(c, coef1, coef2, out1, add, sign, raw1, out2);
    else
      bt_dot_t<4, false>(c, coef1, coef2, out1, add, sign, raw1, out2);
  }
}

// Out[i][c] = sum_{k < K} CM[k][i] * R[k][c]   (Out = CM^T R), i < M, c < ncols,
// on the FP64 tensor cores (mma.sync m8n8k4, the one dense contraction of this
// path). All operands live in the L2 workspace, row-major with even leading
// dimensions. A warp owns a 16 x 32 output tile (2 x 4 MMA tiles); both operand
// fragments have the same access shape (row k0 + lane%4, column base + lane/4),
// so no staging or transposition is needed. `lower` skips the tiles strictly
// above the block diagonal (symmetric result, read as [max][min]).
// Used once per QP to form W = P^-1 B^T and the Gram matrix G = B W of ALL
// constraint rows, which turns every active-set insertion into a gather.
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b)
{
#ifdef PQP_CPU_EMU
  emu::dmma_8x8x4(c0, c1, a, b);
#else
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
#endif
}
__device__ __noinline__ void gemm_tn(const double* __restrict__ CM, int ldc, const double* __restrict__ R, int ldr, int K, int M, int ncols, double* __restrict__ Out, int ldo, bool lower, int col0 = 0)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int mb = (M + 15) >> 4, nbk = (ncols + 31) >> 5;
  _Pragma("unroll 1") for (int tile = warp; tile < mb * nbk; tile += NW) {
    const int i0 = (tile / nbk) << 4, c0 = (tile % nbk) << 5;
    if (lower && col0 + c0 > i0 + 15) continue; // (col0: column of the full matrix that column 0 of Out / R corresponds to)
    double acc[2][4][2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[u][j][0] = acc[u][j][1] = 0.0;
    }
    bool av[2], bv[4];
#pragma unroll
    for (int u = 0; u < 2; ++u) av[u] = i0 + 8 * u + g < M;
#pragma unroll
    for (int j = 0; j < 4; ++j) bv[j] = c0 + 8 * j + g < ncols;
    const double* ap = CM + (size_t)t * ldc + i0 + g;
    const double* bp = R + (size_t)t * ldr + c0 + g;
    _Pragma("unroll 2") for (int k0 = 0; k0 < K; k0 += 4) {
      const bool kv = k0 + t < K;
      double a[2], b[4];
#pragma unroll
      for (int u = 0; u < 2; ++u) a[u] = (av[u] && kv) ? ap[(size_t)k0 * ldc + 8 * u] : 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = (bv[j] && kv) ? bp[(size_t)k0 * ldr + 8 * j] : 0.0;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[u][j][0], acc[u][j][1], a[u], b[j]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int row = i0 + 8 * u + g;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = c0 + 8 * j + 2 * t;
        if (row < M && col < ncols) *reinterpret_cast<double2*>(Out + (size_t)row * ldo + col) = make_double2(acc[u][j][0], acc[u][j][1]);
      }
    }
  }
  __syncthreads();
}

#ifndef PQP_BIG
// ---------------------------------------------------------------------------
// Symmetric matrices in TILE storage (S^-1, and P during its inversion).
// The lower triangle is cut into 32 x 32 tiles (bi, bj), bj <= bi, each stored
// row-major with a row stride of 33 doubles, so that a warp can read a tile by
// rows (lanes = columns) AND by columns (lanes = rows) without bank conflicts.
// Diagonal tiles hold both halves. With this, y = T x needs no cross-lane
// reduction: an off-diagonal tile contributes T x_j to y_i through the column
// reading and T^T x_i to y_j through the row reading, both accumulating into
// lane-stationary registers; rank-k updates touch every stored element once.
// Entries outside the live order are kept at zero. Block row bi holds
// min(32, cap - 32 bi) rows; cap <= 128.
// ---------------------------------------------------------------------------
#define TS_LD 33
#define TS_TILE (32 * TS_LD)
__device__ __forceinline__ int ts_rows(int cap, int bi)
{
  return min(32, cap - 32 * bi);
}
__device__ __forceinline__ int ts_tile(int cap, int bi, int bj)
{
  return ((bi * (bi + 1)) >> 1) * TS_TILE + bj * ts_rows(cap, bi) * TS_LD;
}
// position of (i, j); valid for j's block <= i's block
__device__ __forceinline__ int ts_idx(int cap, int i, int j)
{
  return ts_tile(cap, i >> 5, j >> 5) + (i & 31) * TS_LD + (j & 31);
}
__device__ __forceinline__ double ts_get(const double* T, int cap, int i, int j)
{
  return (i >= j) ? T[ts_idx(cap, i, j)] : T[ts_idx(cap, j, i)];
}
// store (i, j) = (j, i) = v
__device__ __forceinline__ void ts_put(double* T, int cap, int i, int j, double v)
{
  const int hi = i >= j ? i : j, lo = i >= j ? j : i;
  T[ts_idx(cap, hi, lo)] = v;
  if ((hi >> 5) == (lo >> 5) && hi != lo) T[ts_idx(cap, lo, hi)] = v;
}
// doubles covered by the block rows holding order n
__device__ __forceinline__ int ts_extent(int cap, int n)
{
  const int nb = (n + 31) >> 5;
  return nb == 0 ? 0 : ts_tile(cap, nb - 1, 0) + nb * ts_rows(cap, nb - 1) * TS_LD;
}

// y = T x (order n). x and y must not alias. Uses c.scratch (NW x 128).
// want_dot: also returns x . y (block-uniform), reduced in the same final phase: the partial products ride on the
// combination loop, one barrier publishes the warp sums, and there is NO trailing barrier - the caller must pass one
// before c.red is written again (insert_slot: the barrier ahead of the rank-1 update).
__device__ __noinline__ double tsym_mv(const Ctx& c, const double* __restrict__ T, const double* __restrict__ x, double* __restrict__ y, int n, bool want_dot = false)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cap = c.si_cap;
  double* const scr = c.scratch;
  PQP_SM(T);
  PQP_SM(x);
  PQP_SM(y);
  PQP_SM(scr);
  const int nb = (n + 31) >> 5;
  const int r0 = RPB * warp; // this warp's rows (row reading) / columns (column reading) inside every tile
  double acc[4] = { 0.0, 0.0, 0.0, 0.0 };
  double xr[4][RPB];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
#pragma unroll
    for (int r = 0; r < RPB; ++r) {
      const int i = 32 * b + r0 + r;
      xr[b][r] = (i < n) ? x[i] : 0.0;
    }
  }
#pragma unroll
  for (int bi = 0; bi < 4; ++bi) {
    if (bi < nb) {
      const int rows = min(32, n - 32 * bi); // live rows of this block
#pragma unroll
      for (int bj = 0; bj <= bi; ++bj) {
        const double* tile = T + ts_tile(cap, bi, bj);
        if (r0 < rows) {
#pragma unroll
          for (int r = 0; r < RPB; ++r) {
            if (r0 + r < rows) acc[bj] = fma(tile[(r0 + r) * TS_LD + lane], xr[bi][r], acc[bj]);
          }
        }
        if (bj < bi && lane < rows) {
#pragma unroll
          for (int q = 0; q < RPB; ++q) acc[bi] = fma(tile[lane * TS_LD + r0 + q], xr[bj][q], acc[bi]);
        }
      }
    }
  }
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    if (b < nb) scr[warp * 128 + 32 * b + lane] = acc[b];
  }
  __syncthreads();
  double part = 0.0;
  _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += scr[w * 128 + j];
    y[j] = s;
    if (want_dot) part += x[j] * s;
  }
  if (want_dot) { // block-uniform
    double* const red = c.red;
    PQP_SM(red);
    part = warp_sum(part);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    double d = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) d += red[w];
    return d;
  }
  __syncthreads();
  return 0.0;
}

// T[i][j] += u_i v_j on every stored element with i, j < n (u, v in shared memory).
// In a diagonal tile the upper half is computed with the operands of its mirror
// element (u_j v_i for j > i), so the two stored halves stay bit-identical.
__device__ __noinline__ void tsym_rank1(const Ctx& c, double* __restrict__ T, const double* __restrict__ u, const double* __restrict__ v, int n)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cap = c.si_cap;
  PQP_SM(T);
  PQP_SM(u);
  PQP_SM(v);
  const int nb = (n + 31) >> 5;
  const int r0 = RPB * warp;
  double vl[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int j = 32 * b + lane;
    vl[b] = (j < n) ? v[j] : 0.0;
  }
#pragma unroll
  for (int bi = 0; bi < 4; ++bi) {
    if (bi < nb) {
      const int rows = min(32, n - 32 * bi);
      if (r0 < rows) {
        double ur[RPB], vr[RPB];
#pragma unroll
        for (int r = 0; r < RPB; ++r) {
          const int i = 32 * bi + r0 + r;
          ur[r] = (i < n) ? u[i] : 0.0;
          vr[r] = (i < n) ? v[i] : 0.0;
        }
        const double ul = (32 * bi + lane < n) ? u[32 * bi + lane] : 0.0;
#pragma unroll
        for (int bj = 0; bj <= bi; ++bj) {
          double* tile = T + ts_tile(cap, bi, bj);
          double a[RPB];
#pragma unroll
          for (int r = 0; r < RPB; ++r) a[r] = (r0 + r < rows) ? tile[(r0 + r) * TS_LD + lane] : 0.0;
#pragma unroll
          for (int r = 0; r < RPB; ++r) {
            if (r0 + r < rows) {
              double val = fma(ur[r], vl[bj], a[r]);
              if (bj == bi) {
                const double mir = fma(ul, vr[r], a[r]);
                val = (lane <= r0 + r) ? val : mir;
              }
              tile[(r0 + r) * TS_LD + lane] = val;
            }
          }
        }
      }
    }
  }
  __syncthreads();
}

#ifndef PQP_RANK4_DMMA
#define PQP_RANK4_DMMA 1
#endif
#ifndef PQP_DMMA_FAST_PATH
#define PQP_DMMA_FAST_PATH 0 // (measured: -1.3 % at cfg 2, profiles/r02_ab_fast_path.log)
#endif
#if PQP_RANK4_DMMA
// T[i][j] += sum_k U[i][k] V[k][j], k = 0..3, on the FP64 tensor cores: a rank-4 update of an 8 x 8 block IS one
// mma.sync m8n8k4 (A = 8 rows of U, B = 8 columns of V, C = the block; a lane holds C[g][2t], C[g][2t + 1], A[g][t] and
// B[t][g], g = lane / 4, t = lane % 4). The lower triangle of 16 x 16 blocks (P, Q), Q <= P, is dealt round-robin to the
// warps; a block is four DMMAs on two A and two B fragments, its C fragments addressed by immediates off one pointer
// (a 16 x 16 block never straddles a 32 x 32 tile). Against the scalar form (a load, four FMAs and a store per element
// and lane; 17 % of the kernel's warp instructions, profiles/r02_summary.md section 8) that is ~3 x fewer instructions.
// Inside a diagonal 32 x 32 tile both halves are stored: only the lower half (i >= j) is computed, and written to (i, j)
// and (j, i), so the halves stay bit-identical. U is stored row-interleaved (4 doubles per row i), V as four vectors of
// stride ldv.
// R1: the rank-1 update T += u v^T (U = u, V = v plain vectors) through the same blocks - the fragments carry u / v in
// their k = 0 lanes and zeros elsewhere (the three zero products add exactly, so the result is the scalar fma's).
template<bool R1>
__device__ __forceinline__ void tsym_rank_dmma_block(double* __restrict__ T, const double* __restrict__ U, const double* __restrict__ V, int ldv, int n, int cap, int P, int Q, int lane)
{
  const int g = lane >> 2, t = lane & 3;
  const int bi = P >> 1, bj = Q >> 1;
  double* const tile = T + ts_tile(cap, bi, bj);
  const int rt = 16 * (P & 1) + g, ct = 16 * (Q & 1) + 2 * t; // this lane's row / first column inside the tile (h, w add 8)
  const int i0 = 16 * P + g;
#if PQP_DMMA_FAST_PATH
  if (bi != bj && 16 * P + 15 < n) { // a block of an off-diagonal tile with all sixteen rows live: no predicates, no mirror
    double a0, a1, b0, b1;
    if (R1) {
      a0 = (t == 0) ? U[i0] : 0.0;
      a1 = (t == 0) ? U[i0 + 8] : 0.0;
      b0 = (t == 0) ? V[16 * Q + g] : 0.0;
      b1 = (t == 0) ? V[16 * Q + g + 8] : 0.0;
    } else {
      a0 = U[64 * P + lane];
      a1 = U[64 * P + 32 + lane];
      b0 = V[t * ldv + 16 * Q + g];
      b1 = V[t * ldv + 16 * Q + g + 8];
    }
    double* const p = tile + rt * TS_LD + ct;
    double c00x = p[0], c00y = p[1], c01x = p[8], c01y = p[9];
    double c10x = p[8 * TS_LD], c10y = p[8 * TS_LD + 1], c11x = p[8 * TS_LD + 8], c11y = p[8 * TS_LD + 9];
    dmma_8x8x4(c00x, c00y, a0, b0);
    dmma_8x8x4(c01x, c01y, a0, b1);
    dmma_8x8x4(c10x, c10y, a1, b0);
    dmma_8x8x4(c11x, c11y, a1, b1);
    p[0] = c00x;
    p[1] = c00y;
    p[8] = c01x;
    p[9] = c01y;
    p[8 * TS_LD] = c10x;
    p[8 * TS_LD + 1] = c10y;
    p[8 * TS_LD + 8] = c11x;
    p[8 * TS_LD + 9] = c11y;
    return;
  }
#endif
  const bool live0 = i0 < n, live1 = i0 + 8 < n;
  const int jb = 16 * Q + g;
  double a0, a1, b0, b1;
  if (R1) {
    a0 = (live0 && t == 0) ? U[i0] : 0.0;
    a1 = (live1 && t == 0) ? U[i0 + 8] : 0.0;
    b0 = (jb < n && t == 0) ? V[jb] : 0.0;
    b1 = (jb + 8 < n && t == 0) ? V[jb + 8] : 0.0;
  } else {
    a0 = live0 ? U[64 * P + lane] : 0.0;
    a1 = live1 ? U[64 * P + 32 + lane] : 0.0;
    b0 = (jb < n) ? V[t * ldv + jb] : 0.0;
    b1 = (jb + 8 < n) ? V[t * ldv + jb + 8] : 0.0;
  }
  double* const p = tile + rt * TS_LD + ct;
  double c00[2] = { 0.0, 0.0 }, c01[2] = { 0.0, 0.0 }, c10[2] = { 0.0, 0.0 }, c11[2] = { 0.0, 0.0 };
  const bool diag = P == Q;
  if (live0) {
    c00[0] = p[0];
    c00[1] = p[1];
    if (!diag) {
      c01[0] = p[8];
      c01[1] = p[9];
    }
  }
  if (live1) {
    c10[0] = p[8 * TS_LD];
    c10[1] = p[8 * TS_LD + 1];
    c11[0] = p[8 * TS_LD + 8];
    c11[1] = p[8 * TS_LD + 9];
  }
  dmma_8x8x4(c00[0], c00[1], a0, b0);
  if (!diag) dmma_8x8x4(c01[0], c01[1], a0, b1); // (block-uniform: P, Q are)
  dmma_8x8x4(c10[0], c10[1], a1, b0);
  dmma_8x8x4(c11[0], c11[1], a1, b1);
  // In a diagonal 8 x 8 sub-block every lane has read its C pair above, upper-half lanes included, and the mirror stores
  // below write those upper-half elements from other lanes. The reads feed the (warp-converged) mma.sync, so they are
  // complete before any store issues; the barrier states that order for the memory model and for racecheck.
  __syncwarp();
  if (bi != bj) {
    if (live0) {
      p[0] = c00[0];
      p[1] = c00[1];
      p[8] = c01[0];
      p[9] = c01[1];
    }
    if (live1) {
      p[8 * TS_LD] = c10[0];
      p[8 * TS_LD + 1] = c10[1];
      p[8 * TS_LD + 8] = c11[0];
      p[8 * TS_LD + 9] = c11[1];
    }
  } else {
    double* const q = tile + ct * TS_LD + rt; // mirror of this lane's first element: (row, col) -> (col, row)
    const bool lo0 = !diag || g >= 2 * t, lo1 = !diag || g >= 2 * t + 1; // lower half of an 8 x 8 block ON the diagonal
    if (live0) {
      if (lo0) {
        p[0] = c00[0];
        q[0] = c00[0];
      }
      if (lo1) {
        p[1] = c00[1];
        q[TS_LD] = c00[1];
      }
      if (!diag) {
        p[8] = c01[0];
        q[8 * TS_LD] = c01[0];
        p[9] = c01[1];
        q[9 * TS_LD] = c01[1];
      }
    }
    if (live1) {
      p[8 * TS_LD] = c10[0];
      q[8] = c10[0];
      p[8 * TS_LD + 1] = c10[1];
      q[TS_LD + 8] = c10[1];
      if (lo0) {
        p[8 * TS_LD + 8] = c11[0];
        q[8 * TS_LD + 8] = c11[0];
      }
      if (lo1) {
        p[8 * TS_LD + 9] = c11[1];
        q[9 * TS_LD + 8] = c11[1];
      }
    }
  }
}
template<bool R1>
__device__ __forceinline__ void tsym_rank_dmma(const Ctx& c, double* __restrict__ T, const double* __restrict__ U, const double* __restrict__ V, int ldv, int n)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cap = c.si_cap;
  PQP_SM(T);
  PQP_SM(U);
  PQP_SM(V);
  const int nP = (n + 15) >> 4;
  const int total = (nP * (nP + 1)) >> 1;
  int P = 0, Q = warp; // block of linear index `warp` in the row-major enumeration of the lower block triangle
  while (Q > P) {
    Q -= P + 1;
    ++P;
  }
  _Pragma("unroll 1") for (int idx = warp; idx < total; idx += NW) {
    tsym_rank_dmma_block<R1>(T, U, V, ldv, n, cap, P, Q, lane);
    Q += NW;
    while (Q > P) {
      Q -= P + 1;
      ++P;
    }
  }
  __syncthreads();
}
__device__ __noinline__ void tsym_rank4(const Ctx& c, double* __restrict__ T, const double* __restrict__ U, const double* __restrict__ V, int ldv, int n)
{
  tsym_rank_dmma<false>(c, T, U, V, ldv, n);
}
#ifndef PQP_RANK1_DMMA
#define PQP_RANK1_DMMA 1
#endif
#if PQP_RANK1_DMMA
__device__ __noinline__ void tsym_rank1_dmma(const Ctx& c, double* __restrict__ T, const double* __restrict__ u, const double* __restrict__ v, int n)
{
  tsym_rank_dmma<true>(c, T, u, v, 0, n);
}
#endif
#else
// T[i][j] += sum_k U[i][k] V[k][j], k = 0..3 in order. U is stored row-interleaved
// (4 doubles per row i), V as four vectors of stride ldv. Diagonal tiles: as above.
__device__ __noinline__ void tsym_rank4(const Ctx& c, double* __restrict__ T, const double* __restrict__ U, const double* __restrict__ V, int ldv, int n)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cap = c.si_cap;
  PQP_SM(T);
  PQP_SM(U);
  PQP_SM(V);
  const int nb = (n + 31) >> 5;
  const int r0 = RPB * warp;
  double vl[4][4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int j = 32 * b + lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) vl[k][b] = (j < n) ? V[k * ldv + j] : 0.0;
  }
#pragma unroll
  for (int bi = 0; bi < 4; ++bi) {
    if (bi < nb) {
      const int rows = min(32, n - 32 * bi);
      if (r0 < rows) {
        // this lane's own U row of the block (operands of the mirrored half of the diagonal tile)
        double2 la = make_double2(0.0, 0.0), lb = la;
        if (32 * bi + lane < n) {
          la = reinterpret_cast<const double2*>(U)[2 * (32 * bi + lane)];
          lb = reinterpret_cast<const double2*>(U)[2 * (32 * bi + lane) + 1];
        }
#pragma unroll
        for (int r = 0; r < RPB; ++r) {
          if (r0 + r < rows) {
            const int i = 32 * bi + r0 + r;
            const double2 ua = reinterpret_cast<const double2*>(U)[2 * i];
            const double2 ub = reinterpret_cast<const double2*>(U)[2 * i + 1];
            double a[4];
#pragma unroll
            for (int bj = 0; bj <= bi; ++bj) a[bj] = T[ts_tile(cap, bi, bj) + (r0 + r) * TS_LD + lane];
#pragma unroll
            for (int bj = 0; bj < bi; ++bj) T[ts_tile(cap, bi, bj) + (r0 + r) * TS_LD + lane] = fma(ub.y, vl[3][bj], fma(ub.x, vl[2][bj], fma(ua.y, vl[1][bj], fma(ua.x, vl[0][bj], a[bj]))));
            const double low = fma(ub.y, vl[3][bi], fma(ub.x, vl[2][bi], fma(ua.y, vl[1][bi], fma(ua.x, vl[0][bi], a[bi]))));
            const double mir = fma(lb.y, V[3 * ldv + i], fma(lb.x, V[2 * ldv + i], fma(la.y, V[ldv + i], fma(la.x, V[i], a[bi]))));
            T[ts_tile(cap, bi, bi) + (r0 + r) * TS_LD + lane] = (lane <= r0 + r) ? low : mir;
          }
        }
      }
    }
  }
  __syncthreads();
}
#endif // PQP_RANK4_DMMA

#ifndef PQP_SWEEP8
#define PQP_SWEEP8 0
#endif
#if PQP_SWEEP8
// T[i][j] += sum_{k < 8} U[i][k] V[k][j] with V[k][j] = -U[j][k] inv[k] (the rank-8 update of EIGHT consecutive scalar
// sweeps, see tsym_sweep_invert), on the FP64 tensor cores: per 16 x 16 block of the lower block triangle two chained
// DMMA m8n8k4 per 8 x 8 sub-block (pivots 0..3, then 4..7: the FMA order of the scalar sweeps), the C fragments loaded
// and stored once. U0 / U1: pivots 0..3 / 4..7, row-interleaved (4 doubles per row); the B fragments are the A-shaped
// loads of the column rows scaled by this lane's -inv[4 kc + t] (bit-identical to the V the 4-pivot form stored).
__device__ __noinline__ void tsym_rank8(const Ctx& c, double* __restrict__ T, const double* __restrict__ U0, const double* __restrict__ U1, const double* __restrict__ inv, int n)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int cap = c.si_cap;
  PQP_SM(T);
  PQP_SM(U0);
  PQP_SM(U1);
  PQP_SM(inv);
  const double ni0 = -inv[t], ni1 = -inv[4 + t];
  const int nP = (n + 15) >> 4;
  const int total = (nP * (nP + 1)) >> 1;
  int P = 0, Q = warp; // block of linear index `warp` in the row-major enumeration of the lower block triangle
  while (Q > P) {
    Q -= P + 1;
    ++P;
  }
  _Pragma("unroll 1") for (int idx = warp; idx < total; idx += NW) {
    const int bi = P >> 1, bj = Q >> 1;
    double* const tile = T + ts_tile(cap, bi, bj);
    const int rt = 16 * (P & 1) + g, ct = 16 * (Q & 1) + 2 * t; // this lane's row / first column inside the tile (h, w add 8)
    const int i0 = 16 * P + g, jb = 16 * Q + g;
    const bool live0 = i0 < n, live1 = i0 + 8 < n;
    double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0, b00 = 0.0, b01 = 0.0, b10 = 0.0, b11 = 0.0; // [h or w][kc]
    if (live0) {
      a00 = U0[64 * P + lane];
      a01 = U1[64 * P + lane];
    }
    if (live1) {
      a10 = U0[64 * P + 32 + lane];
      a11 = U1[64 * P + 32 + lane];
    }
    if (jb < n) {
      b00 = U0[64 * Q + lane] * ni0;
      b01 = U1[64 * Q + lane] * ni1;
    }
    if (jb + 8 < n) {
      b10 = U0[64 * Q + 32 + lane] * ni0;
      b11 = U1[64 * Q + 32 + lane] * ni1;
    }
    double* const p = tile + rt * TS_LD + ct;
    double c00[2] = { 0.0, 0.0 }, c01[2] = { 0.0, 0.0 }, c10[2] = { 0.0, 0.0 }, c11[2] = { 0.0, 0.0 };
    const bool diag = P == Q;
    if (live0) {
      c00[0] = p[0];
      c00[1] = p[1];
      if (!diag) {
        c01[0] = p[8];
        c01[1] = p[9];
      }
    }
    if (live1) {
      c10[0] = p[8 * TS_LD];
      c10[1] = p[8 * TS_LD + 1];
      c11[0] = p[8 * TS_LD + 8];
      c11[1] = p[8 * TS_LD + 9];
    }
    dmma_8x8x4(c00[0], c00[1], a00, b00);
    dmma_8x8x4(c10[0], c10[1], a10, b00);
    dmma_8x8x4(c11[0], c11[1], a10, b10);
    if (!diag) dmma_8x8x4(c01[0], c01[1], a00, b10); // (block-uniform: P, Q are)
    dmma_8x8x4(c00[0], c00[1], a01, b01);
    dmma_8x8x4(c10[0], c10[1], a11, b01);
    dmma_8x8x4(c11[0], c11[1], a11, b11);
    if (!diag) dmma_8x8x4(c01[0], c01[1], a01, b11);
    __syncwarp(); // (reads of the upper-half elements of a diagonal sub-block before the mirror stores, see tsym_rank_dmma_block)
    if (bi != bj) {
      if (live0) {
        p[0] = c00[0];
        p[1] = c00[1];
        p[8] = c01[0];
        p[9] = c01[1];
      }
      if (live1) {
        p[8 * TS_LD] = c10[0];
        p[8 * TS_LD + 1] = c10[1];
        p[8 * TS_LD + 8] = c11[0];
        p[8 * TS_LD + 9] = c11[1];
      }
    } else {
      double* const q = tile + ct * TS_LD + rt; // mirror of this lane's first element: (row, col) -> (col, row)
      const bool lo0 = !diag || g >= 2 * t, lo1 = !diag || g >= 2 * t + 1; // lower half of an 8 x 8 block ON the diagonal
      if (live0) {
        if (lo0) {
          p[0] = c00[0];
          q[0] = c00[0];
        }
        if (lo1) {
          p[1] = c00[1];
          q[TS_LD] = c00[1];
        }
        if (!diag) {
          p[8] = c01[0];
          q[8 * TS_LD] = c01[0];
          p[9] = c01[1];
          q[9 * TS_LD] = c01[1];
        }
      }
      if (live1) {
        p[8 * TS_LD] = c10[0];
        q[8] = c10[0];
        p[8 * TS_LD + 1] = c10[1];
        q[TS_LD + 8] = c10[1];
        if (lo0) {
          p[8 * TS_LD + 8] = c11[0];
          q[8 * TS_LD + 8] = c11[0];
        }
        if (lo1) {
          p[8 * TS_LD + 9] = c11[1];
          q[9 * TS_LD + 8] = c11[1];
        }
      }
    }
    Q += NW;
    while (Q > P) {
      Q -= P + 1;
      ++P;
    }
  }
  __syncthreads();
}

// In-place inverse of the SPD matrix held in tile storage by BLOCKED symmetric Gauss-Jordan sweeps (Goodnight's sweep
// operator), EIGHT pivots per pass. One scalar sweep on pivot k maps
//   T_kk -> -1/d,  T_kj -> T_kj/d,  T_ij -> T_ij - T_ik T_kj / d   (d = T_kk).
// Eight consecutive sweeps touch an entry outside the pivot rows/columns K only through
//   T_ij += sum_{k in K} U_ik V_kj,  U_ik = T^(k)_ik (column k just before its own sweep), V_kj = -U_jk / d_k,
// and U depends only on row i of the n x 8 panel T[:, K] plus the 8 x 8 pivot block. Warp 0 sweeps the pivot block
// (lane r holds row r, the pivot row is broadcast with shuffles), publishes per pivot 1/d and the scaled pivot row and
// writes the finished block back; after a barrier every thread sweeps its own panel row in registers, writes it back
// and stores U (zero on K); one rank-8 pass on the tensor cores applies the rest with the FMA chain the scalar sweeps
// would have used. Three barriers per eight pivots. After all blocks the array holds -T^-1, negated at the end.
// Requires n <= NT. `uv`: 8 * ldv doubles. Replaces Ldlt::factorize for the blocks this path inverts
// (linalg/dense/ldlt.hpp:718-744, factorize.hpp:91-148).
__device__ __noinline__ void tsym_sweep_invert(const Ctx& c, double* __restrict__ T, double* __restrict__ uv, int ldv, int n)
{
  PQP_SM(T);
  PQP_SM(uv);
  const int cap = c.si_cap;
  double* const U0 = uv;            // [n][4], pivots 0..3 of the pass
  double* const U1 = uv + 4 * ldv;  // [n][4], pivots 4..7
  double* const tab = c.red;        // inv[8] | vK[8][8]
  PQP_SM(tab);
  const int i = threadIdx.x;
  for (int k0 = 0; k0 < n; k0 += 8) {
    const int kb = min(8, n - k0);
    const int ai = i - k0; // position of this row inside the pivot block when 0 <= ai < kb
    const bool inK = (ai >= 0) && (ai < kb);
    // positions of the panel entries (i, k0 .. k0+7): row form (valid when block(k0) <= block(i)),
    // column form (valid when block(i) <= block(k0)); both inside a diagonal tile
    const bool rowv = (k0 >> 5) <= (i >> 5), colv = (i >> 5) <= (k0 >> 5);
    const int rowpos = rowv ? ts_idx(cap, i, k0) : 0;
    const int colpos = colv ? ts_idx(cap, k0, i) : 0;
    double p[8] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
    if (i < n && !inK) {
#pragma unroll
      for (int l = 0; l < 8; ++l) {
        if (l < kb) p[l] = rowv ? T[rowpos + l] : T[colpos + TS_LD * l];
      }
    }
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      const int pb = ts_idx(cap, k0, k0);
      double a[8]; // row `lane` of the (identity padded) pivot block
#pragma unroll
      for (int r = 0; r < 8; ++r) a[r] = (lane < kb && r < kb) ? T[pb + TS_LD * lane + r] : ((lane == r) ? 1.0 : 0.0);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        double rk[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) rk[r] = __shfl_sync(0xffffffffu, a[r], k);
        const double inv = 1.0 / rk[k];
        double vK[8];
#pragma unroll
        for (int l = 0; l < 8; ++l) vK[l] = -rk[l] * inv;
        if (lane == k) {
          tab[k] = inv;
#pragma unroll
          for (int l = 0; l < 8; ++l) tab[8 + 8 * k + l] = vK[l];
        }
        if (lane != k) {
          const double amk = a[k];
#pragma unroll
          for (int l = 0; l < 8; ++l) a[l] = (l == k) ? amk * inv : fma(amk, vK[l], a[l]);
        } else {
#pragma unroll
          for (int l = 0; l < 8; ++l) a[l] = (l == k) ? -inv : a[l] * inv;
        }
      }
      if (lane < kb) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          if (r < kb) T[pb + TS_LD * lane + r] = a[r]; // the finished pivot block (nobody else reads or writes it in this pass)
        }
      }
    }
    __syncthreads(); // table published; every panel row has been read
    if (i < n) {
      double ui[8] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
      if (!inK) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (k < kb) {
            const double inv = tab[k];
            const double pk = p[k];
            ui[k] = pk;
#pragma unroll
            for (int l = 0; l < 8; ++l) p[l] = (l == k) ? pk * inv : fma(pk, tab[8 + 8 * k + l], p[l]);
          }
        }
#pragma unroll
        for (int l = 0; l < 8; ++l) {
          if (l < kb) {
            if (rowv) T[rowpos + l] = p[l];
            if (colv) T[colpos + TS_LD * l] = p[l];
          }
        }
      }
      reinterpret_cast<double2*>(U0)[2 * i] = make_double2(ui[0], ui[1]);
      reinterpret_cast<double2*>(U0)[2 * i + 1] = make_double2(ui[2], ui[3]);
      reinterpret_cast<double2*>(U1)[2 * i] = make_double2(ui[4], ui[5]);
      reinterpret_cast<double2*>(U1)[2 * i + 1] = make_double2(ui[6], ui[7]);
    }
    __syncthreads();
    tsym_rank8(c, T, U0, U1, tab, n);
  }
  const int tot = ts_extent(cap, n);
  _Pragma("unroll 1") for (int e = threadIdx.x; e < tot; e += NT) T[e] = -T[e];
  __syncthreads();
}
#else // PQP_SWEEP8
#ifndef PQP_SWEEP_LA
#define PQP_SWEEP_LA 1
#endif
#if PQP_SWEEP_LA && PQP_RANK4_DMMA
// Sweep of the (identity padded) 4 x 4 pivot block at (k0, k0), redundantly by every lane of ONE warp: publishes per
// pivot 1/d and the scaled pivot row, plus the finished block (tab: inv[4] | vK[4][4] | block [4][4]).
__device__ __forceinline__ void sweep_pivot_block(const double* __restrict__ T, double* __restrict__ tab, int cap, int k0, int kb, int lane)
{
  const int pb = ts_idx(cap, k0, k0);
  double a[4][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int r = 0; r < 4; ++r) a[q][r] = (q < kb && r < kb) ? T[pb + TS_LD * q + r] : ((q == r) ? 1.0 : 0.0);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double inv = 1.0 / a[k][k];
    double vK[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) vK[l] = -a[k][l] * inv;
    if (lane == 0) {
      tab[k] = inv;
#pragma unroll
      for (int l = 0; l < 4; ++l) tab[4 + 4 * k + l] = vK[l];
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
  }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int r = 0; r < 4; ++r) tab[20 + 4 * q + r] = a[q][r];
    }
  }
}

// In-place inverse of the SPD matrix held in tile storage by BLOCKED symmetric Gauss-Jordan sweeps (Goodnight's sweep
// operator, four pivots per pass), the pivot phase one pass AHEAD. One scalar sweep on pivot k maps
//   T_kk -> -1/d,  T_kj -> T_kj/d,  T_ij -> T_ij - T_ik T_kj / d   (d = T_kk).
// Four consecutive sweeps touch an entry outside the pivot rows/columns K only through
//   T_ij += sum_{k in K} U_ik V_kj,  U_ik = T^(k)_ik (column k just before its own sweep), V_kj = -U_jk / d_k,
// and U depends only on row i of the n x 4 panel T[:, K] plus the 4 x 4 pivot block. Per pass: every thread sweeps its
// own panel row in registers with the published pivot table, writes the finished K rows / columns back and stores U, V
// (zero on K); after a barrier the rank-4 pass (tensor cores, tsym_rank_dmma_block) applies the rest - and the warp
// that owns the 16 x 16 block holding the NEXT pivot block updates that block first, sweeps the 4 x 4 pivot block (the
// only sequential chain: four dependent reciprocals) and publishes the next table while the other warps are still
// updating: the chain is off the critical path and a pass has two barriers instead of three. Same arithmetic, element
// by element, as the three-barrier form (kept below). After all blocks the array holds -T^-1, negated at the end.
// Requires n <= NT. `uv`: 8 * ldv doubles. Replaces Ldlt::factorize for the blocks this path inverts
// (linalg/dense/ldlt.hpp:718-744, factorize.hpp:91-148).
__device__ __noinline__ void tsym_sweep_invert(const Ctx& c, double* __restrict__ T, double* __restrict__ uv, int ldv, int n)
{
  PQP_SM(T);
  PQP_SM(uv);
  const int cap = c.si_cap;
  double* const U = uv;            // [n][4]
  double* const V = uv + 4 * ldv;  // [4][ldv]
  double* const tab = c.red;       // inv[4] | vK[4][4] | final pivot block [4][4]
  PQP_SM(tab);
  const int i = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nP = (n + 15) >> 4;
  const int total = (nP * (nP + 1)) >> 1;
  if (warp == 0) sweep_pivot_block(T, tab, cap, 0, min(4, n), lane);
  __syncthreads();
  for (int k0 = 0; k0 < n; k0 += 4) {
    const int kb = min(4, n - k0);
    const int ai = i - k0; // position of this row inside the pivot block when 0 <= ai < kb
    const bool inK = (ai >= 0) && (ai < kb);
    // positions of the panel entries (i, k0 .. k0+3): row form (valid when block(k0) <= block(i)),
    // column form (valid when block(i) <= block(k0)); both inside a diagonal tile
    const bool rowv = (k0 >> 5) <= (i >> 5), colv = (i >> 5) <= (k0 >> 5);
    const int rowpos = rowv ? ts_idx(cap, i, k0) : 0;
    const int colpos = colv ? ts_idx(cap, k0, i) : 0;
    if (i < n) {
      double p[4] = { 0.0, 0.0, 0.0, 0.0 };
      double ui[4] = { 0.0, 0.0, 0.0, 0.0 }, vi[4] = { 0.0, 0.0, 0.0, 0.0 };
      if (!inK) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          if (l < kb) p[l] = rowv ? T[rowpos + l] : T[colpos + TS_LD * l];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < kb) {
            const double inv = tab[k];
            const double pk = p[k];
            ui[k] = pk;
            vi[k] = -pk * inv;
#pragma unroll
            for (int l = 0; l < 4; ++l) p[l] = (l == k) ? pk * inv : fma(pk, tab[4 + 4 * k + l], p[l]);
          }
        }
      } else {
#pragma unroll
        for (int l = 0; l < 4; ++l) p[l] = tab[20 + 4 * ai + l]; // finished pivot row
      }
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        if (l < kb) {
          if (rowv) T[rowpos + l] = p[l];
          if (colv && !inK) T[colpos + TS_LD * l] = p[l];
        }
        V[l * ldv + i] = vi[l];
      }
      reinterpret_cast<double2*>(U)[2 * i] = make_double2(ui[0], ui[1]);
      reinterpret_cast<double2*>(U)[2 * i + 1] = make_double2(ui[2], ui[3]);
    }
    __syncthreads(); // panel, U, V complete; every reader of the table is past it
    // rank-4 pass; the block holding the next pivot block first, by its owner, which then sweeps that pivot block
    const int kn = k0 + 4;
    const int Pn = kn >> 4, idxn = (kn < n) ? ((Pn * (Pn + 1)) >> 1) + Pn : -1;
    if (idxn >= 0 && (idxn & (NW - 1)) == warp) {
      tsym_rank_dmma_block<false>(T, U, V, ldv, n, cap, Pn, Pn, lane);
      __syncwarp();
      sweep_pivot_block(T, tab, cap, kn, min(4, n - kn), lane);
    }
    int P = 0, Q = warp;
    while (Q > P) {
      Q -= P + 1;
      ++P;
    }
    _Pragma("unroll 1") for (int idx = warp; idx < total; idx += NW) {
      if (idx != idxn) tsym_rank_dmma_block<false>(T, U, V, ldv, n, cap, P, Q, lane);
      Q += NW;
      while (Q > P) {
        Q -= P + 1;
        ++P;
      }
    }
    __syncthreads();
  }
  const int tot = ts_extent(cap, n);
  _Pragma("unroll 1") for (int e = threadIdx.x; e < tot; e += NT) T[e] = -T[e];
  __syncthreads();
}
#else // PQP_SWEEP_LA
// In-place inverse of the SPD matrix held in tile storage by BLOCKED symmetric
// Gauss-Jordan sweeps (Goodnight's sweep operator, four pivots per pass).
// One scalar sweep on pivot k maps
//   T_kk -> -1/d,  T_kj -> T_kj/d,  T_ij -> T_ij - T_ik T_kj / d   (d = T_kk).
// Four consecutive sweeps touch an entry outside the pivot rows/columns K only
// through  T_ij += sum_{k in K} U_ik V_kj,  U_ik = T^(k)_ik (column k just
// before its own sweep), V_kj = -U_jk / d_k, and U^(k) depends only on row i of
// the n x 4 panel T[:, K] plus the 4 x 4 pivot block. Every thread sweeps its
// own panel row in registers (the pivot block is swept redundantly by all),
// writes the finished K rows/columns back, stores U, V (zero on K), and one
// rank-4 pass applies the rest with the FMA chain the scalar sweeps would have
// used. After all blocks the array holds -T^-1. Requires n <= NT.
// `uv`: 8 * ldv doubles. Replaces Ldlt::factorize for the blocks this path
// inverts (linalg/dense/ldlt.hpp:718-744, factorize.hpp:91-148).
__device__ __noinline__ void tsym_sweep_invert(const Ctx& c, double* __restrict__ T, double* __restrict__ uv, int ldv, int n)
{
  PQP_SM(T);
  PQP_SM(uv);
  const int cap = c.si_cap;
  double* const U = uv;            // [n][4]
  double* const V = uv + 4 * ldv;  // [4][ldv]
  double* const tab = c.red;       // inv[4] | vK[4][4] | final pivot block [4][4]
  PQP_SM(tab);
  const int i = threadIdx.x;
  for (int k0 = 0; k0 < n; k0 += 4) {
    const int kb = min(4, n - k0);
    const int ai = i - k0; // position of this row inside the pivot block when 0 <= ai < kb
    const bool inK = (ai >= 0) && (ai < kb);
    // positions of the panel entries (i, k0 .. k0+3): row form (valid when block(k0) <= block(i)),
    // column form (valid when block(i) <= block(k0)); both inside a diagonal tile
    const bool rowv = (k0 >> 5) <= (i >> 5), colv = (i >> 5) <= (k0 >> 5);
    const int rowpos = rowv ? ts_idx(cap, i, k0) : 0;
    const int colpos = colv ? ts_idx(cap, k0, i) : 0;
    double p[4] = { 0.0, 0.0, 0.0, 0.0 };
    if (i < n && !inK) {
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        if (l < kb) p[l] = rowv ? T[rowpos + l] : T[colpos + TS_LD * l];
      }
    }
    if (threadIdx.x < 32) {
      // warp 0 sweeps the (identity padded) pivot block and publishes, per pivot, 1/d and the
      // scaled pivot row, plus the finished block
      const int pb = ts_idx(cap, k0, k0);
      double a[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int r = 0; r < 4; ++r) a[q][r] = (q < kb && r < kb) ? T[pb + TS_LD * q + r] : ((q == r) ? 1.0 : 0.0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double inv = 1.0 / a[k][k];
        double vK[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) vK[l] = -a[k][l] * inv;
        if (threadIdx.x == 0) {
          tab[k] = inv;
#pragma unroll
          for (int l = 0; l < 4; ++l) tab[4 + 4 * k + l] = vK[l];
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
      }
      if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int r = 0; r < 4; ++r) tab[20 + 4 * q + r] = a[q][r];
        }
      }
    }
    __syncthreads(); // table published; every panel row has been read
    if (i < n) {
      double ui[4] = { 0.0, 0.0, 0.0, 0.0 }, vi[4] = { 0.0, 0.0, 0.0, 0.0 };
      if (!inK) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < kb) {
            const double inv = tab[k];
            const double pk = p[k];
            ui[k] = pk;
            vi[k] = -pk * inv;
#pragma unroll
            for (int l = 0; l < 4; ++l) p[l] = (l == k) ? pk * inv : fma(pk, tab[4 + 4 * k + l], p[l]);
          }
        }
      } else {
#pragma unroll
        for (int l = 0; l < 4; ++l) p[l] = tab[20 + 4 * ai + l]; // finished pivot row
      }
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        if (l < kb) {
          if (rowv) T[rowpos + l] = p[l];
          if (colv && !inK) T[colpos + TS_LD * l] = p[l];
        }
        V[l * ldv + i] = vi[l];
      }
      reinterpret_cast<double2*>(U)[2 * i] = make_double2(ui[0], ui[1]);
      reinterpret_cast<double2*>(U)[2 * i + 1] = make_double2(ui[2], ui[3]);
    }
    __syncthreads();
    tsym_rank4(c, T, U, V, ldv, n);
  }
  const int tot = ts_extent(cap, n);
  _Pragma("unroll 1") for (int e = threadIdx.x; e < tot; e += NT) T[e] = -T[e];
  __syncthreads();
}

#endif // PQP_SWEEP_LA
#endif // PQP_SWEEP8
#else // PQP_BIG
#ifdef PQP_CPU_EMU
#define PQP_LOADS_FIRST() ((void)0)
#define PQP_IN_SMEM(p) ((void)0)
#else
#define PQP_LOADS_FIRST() asm volatile("" ::: "memory")
#define PQP_IN_SMEM(p) __builtin_assume(__isShared(p))
#endif
// ---------------------------------------------------------------------------
// BIG variant: symmetric matrices (S^-1, and P during its inversion) as a PACKED lower triangle (row i holds the
// columns 0..i) in the per-CTA global workspace (L2 / HBM), any order. Same interface as the tile storage above;
// the primitives are loop based (no compile-time block count) and keep several independent loads in flight.
// ---------------------------------------------------------------------------
// Row i (columns 0..i) starts at an EVEN offset: rows are padded to an even length, so that a lane can move column
// PAIRS with 16-byte loads / stores (twice the bytes in flight per instruction: the passes over the triangle are bound
// by what is in flight). start(i) = i (i + 2) / 2 for even i, (i + 1)^2 / 2 for odd i. The padding
// element of an even row (column i + 1) is never read as data.
__device__ __forceinline__ int ts_idx(int, int i, int j)
{
  return ((i + 1) >> 1) * (i + 2 - (i & 1)) + j; // j <= i
}
__device__ __forceinline__ double ts_get(const double* T, int, int i, int j)
{
  return (i >= j) ? T[ts_idx(0, i, j)] : T[ts_idx(0, j, i)];
}
__device__ __forceinline__ void ts_put(double* T, int, int i, int j, double v)
{
  const int hi = i >= j ? i : j, lo = i >= j ? j : i;
  T[ts_idx(0, hi, lo)] = v;
}
__device__ __forceinline__ int ts_extent(int, int n)
{
  return ((n + 1) >> 1) * (n + 2 - (n & 1)); // = ts_idx(n, 0)
}

// y = T x (order n), x and y must not alias. Uses c.scratch (NW x n doubles).
//   y_i = sum_{j <= i} T[i][j] x_j  (row part: a warp owns row i, coalesced loads, one warp reduction)
//       + sum_{i' > i} T[i'][i] x_i' (column part, AXPY form: lane-stationary accumulators for a group of 256 columns)
// Row i belongs to the same lane-0 thread in every column group (256 is a multiple of NW): fixed summation order.
// want_dot: as in the tile version (x . y comes out of the final phase, no trailing barrier).
// A lane owns four column PAIRS of the group (16-byte loads), two rows of the warp are in flight.
__device__ __noinline__ double tsym_mv(const Ctx& c, const double* __restrict__ T, const double* __restrict__ x, double* __restrict__ y, int n, bool want_dot = false)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* const scr = c.scratch;
  _Pragma("unroll 1") for (int g0 = 0; g0 < n; g0 += 256) {
    double2 acc[4], xl[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int j = g0 + 2 * (lane + 32 * cc);
      acc[cc] = make_double2(0.0, 0.0);
      xl[cc] = make_double2((j < n) ? x[j] : 0.0, (j + 1 < n) ? x[j + 1] : 0.0);
    }
    _Pragma("unroll 1") for (int i = g0 + warp; i < n; i += 2 * NW) { // two rows of the warp in flight
      if (c.pf) { // the two rows `pf` iterations ahead, sixteen lanes each
        const int ip = i + c.pf * 2 * NW + (lane >> 4) * NW;
        if (ip < n) pf_l2_span(T + ts_idx(0, ip, 0) + g0, min(256, ip + 1 - g0), lane & 15, 16);
      }
      const int i2 = i + NW;
      const bool r2 = i2 < n;
      const double2* row = reinterpret_cast<const double2*>(T + ts_idx(0, i, 0)) + (g0 >> 1);
      const double2* row2 = reinterpret_cast<const double2*>(T + ts_idx(0, r2 ? i2 : i, 0)) + (g0 >> 1);
      const double xi = x[i], xi2 = r2 ? x[i2] : 0.0;
      double2 a[4], b[4];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int j = g0 + 2 * (lane + 32 * cc);
        a[cc] = (j <= i) ? row[lane + 32 * cc] : make_double2(0.0, 0.0);
        b[cc] = (r2 && j <= i2) ? row2[lane + 32 * cc] : make_double2(0.0, 0.0);
      }
      double d = 0.0, d2 = 0.0;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int j = g0 + 2 * (lane + 32 * cc);
        if (j + 1 > i) a[cc].y = 0.0;  // padding element of the row (or beyond it)
        if (j + 1 > i2) b[cc].y = 0.0;
        d = fma(a[cc].x, xl[cc].x, d);
        d = fma(a[cc].y, xl[cc].y, d);
        d2 = fma(b[cc].x, xl[cc].x, d2);
        d2 = fma(b[cc].y, xl[cc].y, d2);
        if (j < i) acc[cc].x = fma(a[cc].x, xi, acc[cc].x);
        if (j + 1 < i) acc[cc].y = fma(a[cc].y, xi, acc[cc].y);
        if (r2 && j < i2) acc[cc].x = fma(b[cc].x, xi2, acc[cc].x);
        if (r2 && j + 1 < i2) acc[cc].y = fma(b[cc].y, xi2, acc[cc].y);
      }
      d = warp_sum(d);
      d2 = warp_sum(d2);
      if (lane == 0) {
        y[i] = (g0 == 0) ? d : y[i] + d;
        if (r2) y[i2] = (g0 == 0) ? d2 : y[i2] + d2;
      }
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int j = g0 + 2 * (lane + 32 * cc);
      if (j < n) scr[(size_t)warp * n + j] = acc[cc].x;
      if (j + 1 < n) scr[(size_t)warp * n + j + 1] = acc[cc].y;
    }
  }
  __syncthreads();
  double part = 0.0;
  _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) {
    double sacc = y[j];
#pragma unroll
    for (int w = 0; w < NW; ++w) sacc += scr[(size_t)w * n + j];
    y[j] = sacc;
    if (want_dot) part += x[j] * sacc;
  }
  if (want_dot) { // block-uniform
    part = warp_sum(part);
    if (lane == 0) c.red[warp] = part;
    __syncthreads();
    double d = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) d += c.red[w];
    return d;
  }
  __syncthreads();
  return 0.0;
}

// T[i][j] += u_i v_j on the packed lower triangle (j <= i < n). Two rows of the warp x four column pairs per lane are in
// flight (128 bytes per thread; the triangle lives in L2 / HBM: with two 8-byte loads in flight a 300 x 300 pass ran
// at one element per 300 cycles per thread). The loads are forced ahead of the stores with a compiler barrier (ptxas
// otherwise sinks every load next to its use).
__device__ __noinline__ void tsym_rank1(const Ctx& c, double* __restrict__ T, const double* __restrict__ u, const double* __restrict__ v, int n)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  _Pragma("unroll 1") for (int i = warp; i < n; i += 2 * NW) {
    if (c.pf) {
      const int ip = i + c.pf * 2 * NW + (lane >> 4) * NW;
      if (ip < n) pf_l2_span(T + ts_idx(0, ip, 0), ip + 1, lane & 15, 16);
    }
    const int i2 = i + NW;
    const bool r2 = i2 < n;
    double2* const row = reinterpret_cast<double2*>(T + ts_idx(0, i, 0));
    double2* const row2 = reinterpret_cast<double2*>(T + ts_idx(0, r2 ? i2 : i, 0));
    const double ui = u[i], ui2 = r2 ? u[i2] : 0.0;
    const int last = r2 ? i2 : i; // the longer of the two rows
    _Pragma("unroll 1") for (int p0 = lane; 2 * p0 <= last; p0 += 128) {
      double2 a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = 2 * (p0 + 32 * q);
        a[q] = (j <= i) ? row[p0 + 32 * q] : make_double2(0.0, 0.0);
        b[q] = (r2 && j <= i2) ? row2[p0 + 32 * q] : make_double2(0.0, 0.0);
      }
      PQP_LOADS_FIRST();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = 2 * (p0 + 32 * q);
        if (j <= last) {
          const double v0 = v[j], v1 = (j + 1 <= last) ? v[j + 1] : 0.0;
          if (j <= i) {
            a[q].x = fma(ui, v0, a[q].x);
            if (j + 1 <= i) a[q].y = fma(ui, v1, a[q].y);
            row[p0 + 32 * q] = a[q];
          }
          if (r2) { // (j <= i2 holds)
            b[q].x = fma(ui2, v0, b[q].x);
            if (j + 1 <= i2) b[q].y = fma(ui2, v1, b[q].y);
            row2[p0 + 32 * q] = b[q];
          }
        }
      }
    }
  }
  (void)c;
  __syncthreads();
}

// T[i][j] += sum_{k<4} U_k[i] V_k[j] (j <= i < n), k = 0..3 in order; U_k = U + k ldv, V_k = V + k ldv (shared memory;
// fallback: the global workspace)
template<bool PANEL_IN_SMEM>
__device__ __forceinline__ void tsym_rank4_body(double* __restrict__ T, const double* __restrict__ U, const double* __restrict__ V, int ldv, int n, int pf)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (PANEL_IN_SMEM) {
    PQP_IN_SMEM(U);
    PQP_IN_SMEM(V);
  }
  _Pragma("unroll 1") for (int i = warp; i < n; i += 2 * NW) {
    if (pf) {
      const int ip = i + pf * 2 * NW + (lane >> 4) * NW;
      if (ip < n) pf_l2_span(T + ts_idx(0, ip, 0), ip + 1, lane & 15, 16);
    }
    const int i2 = i + NW;
    const bool r2 = i2 < n;
    double2* const row = reinterpret_cast<double2*>(T + ts_idx(0, i, 0));
    double2* const row2 = reinterpret_cast<double2*>(T + ts_idx(0, r2 ? i2 : i, 0));
    const double u0 = U[i], u1 = U[ldv + i], u2 = U[2 * ldv + i], u3 = U[3 * ldv + i];
    const int ic = r2 ? i2 : i;
    const double w0 = r2 ? U[ic] : 0.0, w1 = r2 ? U[ldv + ic] : 0.0, w2 = r2 ? U[2 * ldv + ic] : 0.0, w3 = r2 ? U[3 * ldv + ic] : 0.0;
    const int last = ic;
    _Pragma("unroll 1") for (int p0 = lane; 2 * p0 <= last; p0 += 128) {
      double2 a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = 2 * (p0 + 32 * q);
        a[q] = (j <= i) ? row[p0 + 32 * q] : make_double2(0.0, 0.0);
        b[q] = (r2 && j <= i2) ? row2[p0 + 32 * q] : make_double2(0.0, 0.0);
      }
      PQP_LOADS_FIRST();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = 2 * (p0 + 32 * q);
        if (j <= last) {
          const bool two = j + 1 <= last;
          const double v00 = V[j], v10 = V[ldv + j], v20 = V[2 * ldv + j], v30 = V[3 * ldv + j];
          const double v01 = two ? V[j + 1] : 0.0, v11 = two ? V[ldv + j + 1] : 0.0, v21 = two ? V[2 * ldv + j + 1] : 0.0, v31 = two ? V[3 * ldv + j + 1] : 0.0;
          if (j <= i) {
            a[q].x = fma(u3, v30, fma(u2, v20, fma(u1, v10, fma(u0, v00, a[q].x))));
            if (j + 1 <= i) a[q].y = fma(u3, v31, fma(u2, v21, fma(u1, v11, fma(u0, v01, a[q].y))));
            row[p0 + 32 * q] = a[q];
          }
          if (r2) {
            b[q].x = fma(w3, v30, fma(w2, v20, fma(w1, v10, fma(w0, v00, b[q].x))));
            if (j + 1 <= i2) b[q].y = fma(w3, v31, fma(w2, v21, fma(w1, v11, fma(w0, v01, b[q].y))));
            row2[p0 + 32 * q] = b[q];
          }
        }
      }
    }
  }
}
#ifndef PQP_BIG_RANK4_DMMA
#define PQP_BIG_RANK4_DMMA 1
#endif
#ifndef PQP_BIG_DMMA_MAX_N
#define PQP_BIG_DMMA_MAX_N 288
#endif
// The same update on the FP64 tensor cores (packed storage): the lower triangle in blocks of 16 rows x 32 columns, dealt
// round-robin to the warps; an 8 x 8 sub-block is one mma.sync m8n8k4 (A = 8 rows of U^T, B = 8 columns of V; a lane
// holds C[g][2t], C[g][2t + 1] = one 16-byte pair of the packed row, g = lane / 4, t = lane % 4). Eight pairs per lane
// are in flight per block (128 bytes, as in the scalar form); sub-blocks above the diagonal are skipped, elements above
// it masked (a pair starting ON the diagonal of an even row carries the row's padding element, never read as data).
// Roughly half the instructions of the scalar form (four FMAs per element and lane).
template<bool PANEL_IN_SMEM>
__device__ __forceinline__ void tsym_rank4_dmma_body(double* __restrict__ T, const double* __restrict__ U, const double* __restrict__ V, int ldv, int n, int pf)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  if (PANEL_IN_SMEM) {
    PQP_IN_SMEM(U);
    PQP_IN_SMEM(V);
  }
  const int nR = (n + 15) >> 4;
  int R = 0, Cb = warp; // row block R (16 rows) has the column blocks 0 .. R / 2 (32 columns each)
  while (R < nR && Cb > (R >> 1)) {
    Cb -= (R >> 1) + 1;
    ++R;
  }
  int Rp = R, Cp = Cb + pf * NW; // cursor of the software prefetch, `pf` turns ahead
  while (Rp < nR && Cp > (Rp >> 1)) {
    Cp -= (Rp >> 1) + 1;
    ++Rp;
  }
  _Pragma("unroll 1") while (R < nR) {
    const int i0 = 16 * R + g, jc = 32 * Cb;
    const bool live0 = i0 < n, live1 = i0 + 8 < n;
    if (pf) { // the block this warp takes `pf` turns ahead: one 128-byte line per lane (16 rows x 2 lines)
      const int ip = 16 * Rp + (lane & 15), jp = 32 * Cp + 16 * (lane >> 4);
      if (Rp < nR && ip < n && jp <= ip) pf_l2_span(T + ts_idx(0, ip, 0) + jp, 1, 0, 1);
      Cp += NW;
      while (Rp < nR && Cp > (Rp >> 1)) {
        Cp -= (Rp >> 1) + 1;
        ++Rp;
      }
    }
    double a[2], b[4];
    a[0] = live0 ? U[t * ldv + i0] : 0.0;
    a[1] = live1 ? U[t * ldv + i0 + 8] : 0.0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int j = jc + 8 * w + g;
      b[w] = (j < n) ? V[t * ldv + j] : 0.0;
    }
    double* const r0 = T + ts_idx(0, live0 ? i0 : 0, 0);
    double* const r1 = T + ts_idx(0, live1 ? i0 + 8 : 0, 0);
    double2 cc[2][4];
    bool ex[2][4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int j = jc + 8 * w + 2 * t;
      ex[0][w] = live0 && j <= i0;
      ex[1][w] = live1 && j <= i0 + 8;
      cc[0][w] = ex[0][w] ? *reinterpret_cast<const double2*>(r0 + j) : make_double2(0.0, 0.0);
      cc[1][w] = ex[1][w] ? *reinterpret_cast<const double2*>(r1 + j) : make_double2(0.0, 0.0);
    }
    PQP_LOADS_FIRST();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        if (jc + 8 * w <= 16 * R + 8 * h + 7) dmma_8x8x4(cc[h][w].x, cc[h][w].y, a[h], b[w]); // (warp-uniform: sub-block not above the diagonal)
      }
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int j = jc + 8 * w + 2 * t;
      if (ex[0][w]) *reinterpret_cast<double2*>(r0 + j) = cc[0][w];
      if (ex[1][w]) *reinterpret_cast<double2*>(r1 + j) = cc[1][w];
    }
    Cb += NW;
    while (R < nR && Cb > (R >> 1)) {
      Cb -= (R >> 1) + 1;
      ++R;
    }
  }
}
__device__ __noinline__ void tsym_rank4(const Ctx& c, double* __restrict__ T, const double* __restrict__ U, const double* __restrict__ V, int ldv, int n)
{
  // the sweep's / block updates' panel vectors live in c.scratch: shared memory in every layout of this variant; the
  // fallback keeps them in the global workspace (order n + n_slots does not fit shared memory)
  // tensor-core blocks up to order PQP_BIG_DMMA_MAX_N, whole-row streaming above (same-box A/B, profiles/r02_ab_big_rank4_dmma.log:
  // cfg 3 (orders ~100-130, L1 resident) +12.5 %, cfg 4 (orders 230-256, L2) +1.8 %, cfg 5 (orders 350-450, HBM) -2.7 %: there
  // the 256-byte row segments of a block lose to rows streamed end to end)
#if PQP_BIG_RANK4_DMMA
  const bool dmma = n <= PQP_BIG_DMMA_MAX_N; // block-uniform
#else
  const bool dmma = false;
#endif
  if (c.kkt_mode) {
    if (dmma)
      tsym_rank4_dmma_body<false>(T, U, V, ldv, n, c.pf);
    else
      tsym_rank4_body<false>(T, U, V, ldv, n, c.pf);
  } else {
    if (dmma)
      tsym_rank4_dmma_body<true>(T, U, V, ldv, n, c.pf);
    else
      tsym_rank4_body<true>(T, U, V, ldv, n, c.pf);
  }
  __syncthreads();
}

// In-place inverse of the SPD matrix in packed storage by blocked symmetric Gauss-Jordan sweeps (four pivots per
// pass over the triangle; see the tile version above for the algebra). Any n: a thread sweeps the panel rows
// i = tid, tid + NT, ... in registers (the 4 x 4 pivot block is swept redundantly by every thread). `uv`: 8 ldv
// doubles. After all blocks the array holds -T^-1, negated at the end.
__device__ __noinline__ void tsym_sweep_invert(const Ctx& c, double* __restrict__ T, double* __restrict__ uv, int ldv, int n)
{
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
        a0[a][bq] = (a < kb && bq < kb) ? T[ts_idx(0, hi, lo)] : ((a == bq) ? 1.0 : 0.0);
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
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const int col = k0 + l;
        double v = 0.0;
        if (l < kb) {
          if (inK) { // rows of the pivot block take their panel from the register copy (their entries are being overwritten)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (ai == q) v = a0[q][l];
            }
          } else {
            v = (i >= col) ? T[ts_idx(0, i, col)] : T[ts_idx(0, col, i)];
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
            T[ts_idx(0, i, col)] = p[l];
          else if (!inK)
            T[ts_idx(0, col, i)] = p[l];
        }
        U[l * ldv + i] = ui[l];
        V[l * ldv + i] = vi[l];
      }
    }
    __syncthreads();
    tsym_rank4(c, T, U, V, ldv, n);
  }
  const int tot = ts_extent(0, n);
  _Pragma("unroll 1") for (int e = threadIdx.x; e < tot; e += NT) T[e] = -T[e];
  __syncthreads();
}
#endif // PQP_BIG
#undef PQP_TSYM_RANK1
#if !defined(PQP_BIG) && PQP_RANK1_DMMA
#define PQP_TSYM_RANK1 tsym_rank1_dmma // (tile storage: the rank-1 updates of insertion / deletion on the tensor-core blocks)
#else
#define PQP_TSYM_RANK1 tsym_rank1
#endif
#ifndef PQP_BTLAM_AXPY
#define PQP_BTLAM_AXPY 0 // tile kernel: B^T lam of solve_kkt as an AXPY pass over the active rows of A_s / C_s. Measured (profiles/r02_ab_btlam_axpy.log):
                         // 25.07 -> 26.84 ms (-7 %): A_s / C_s join the L2 working set (+120 KB per CTA) and the extra call costs 90 bytes of spills
#endif
#ifndef PQP_TILE_BLOCK
#define PQP_TILE_BLOCK 0 // tile kernel: block (rank-4) insertion / deletion as in the big variant (A/B switch)
#endif

__device__ __forceinline__ int row_id(const Ctx& c, int s)
{
  return s < c.ne ? s : c.ne + c.slot_cons[s];
}

// y = P^-1 v  (Pi = P^-1 explicit, full square n x ldn in the L2 workspace)
__device__ __forceinline__ void apply_Pinv(const Ctx& c, const double* v, double* y)
{
#ifdef PQP_BIG
  if (c.hess != PQP_HESSIAN_DENSE) { // P^-1 = diag(1 / (H_jj + rho)) (Diagonal) or I / rho (Zero)
    _Pragma("unroll 1") for (int j = threadIdx.x; j < c.n; j += NT) y[j] = v[j] * c.d1inv[j];
    __syncthreads();
    return;
  }
#endif
  axpy_pass(c, c.Pi, c.ldn, c.n, v, c.n, y, nullptr, 1.0);
}

// out = Hmat x for the Hessian type of the batch (the tile kernel proper only takes dense Hessians)
__device__ __forceinline__ void apply_H(const Ctx& c, const double* Hmat, const double* x, double* out)
{
#ifdef PQP_BIG
  if (c.hess != PQP_HESSIAN_DENSE) {
    const int n = c.n;
    _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) out[j] = (c.hess == PQP_HESSIAN_DIAGONAL) ? Hmat[(size_t)j * n + j] * x[j] : 0.0;
    __syncthreads();
    return;
  }
#endif
  axpy_pass(c, Hmat, c.n, c.n, x, c.n, out, nullptr, 1.0);
}

// out[0 .. m) = [A_s; C_s; box rows] coef = Bt^T coef (AXPY form over the rows of Bt)
__device__ __forceinline__ void bt_axpy(const Ctx& c, const double* coef, double* out)
{
#ifdef PQP_BIG
  if (c.box) {
    // the box block of Bt is the scaled identity pattern i_s[k] e_k: its n x n zeros are not streamed (half of every
    // constraint pass at cfg 3 / cfg 5); the products are exact copies of what the dense pass would give
    const int nr = c.ne + c.ni;
    axpy_pass(c, c.Bt, c.ldb, c.n, coef, nr, out, nullptr, 1.0);
    _Pragma("unroll 1") for (int k = threadIdx.x; k < c.n; k += NT) out[nr + k] = coef[k] * c.is[k];
    __syncthreads();
    return;
  }
#endif
  axpy_pass(c, c.Bt, c.ldb, c.n, coef, c.m, out, nullptr, 1.0);
}

#ifdef PQP_BIG
// ---------------------------------------------------------------------------
// Last-resort fallback of the BIG variant: the explicit inverse of the WHOLE regularised KKT matrix
//   K = [ H_s + rho I   B^T ;  B   -Dlt ]   (order n + n_slots, quasi-definite: every pivot of the sweep is non-zero),
// re-formed from scratch by the blocked sweep whenever the active set or mu changed. The block elimination through
// S = Dlt + B P^-1 B^T squares the conditioning (cond S ~ 1e13-1e15 on the degenerate LP-like Maros-Meszaros problems,
// cond K ~ 1e7): there the S^-1 path needs ten refinement rounds per digit and stagnates, while K^-1 gains 2-4 digits
// per round (numpy prototype on QSCORPIO). It costs a sweep over (n + n_s)^2 / 2 elements per active-set change, so a
// QP only switches to it when iterative_solve has failed even after re-forming the dual block.
// Storage: K^-1 replaces S^-1 (the region is sized for order n + capacity); panels, partial sums, right-hand side and
// solution live in the W region of the workspace (free once G is built).
// ---------------------------------------------------------------------------
__device__ __noinline__ void kkt_factor(Ctx& c, double rho, double mu_eq, double mu_in)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = c.n, ns = c.ns, m = n + ns, ne = c.ne;
  double* const K = c.Si;
  for (int i = warp; i < n; i += NW) {
    double* row = K + ts_idx(0, i, 0);
    const double* h = c.Hs + (size_t)i * n;
    for (int j = lane; j <= i; j += 32) {
      double v = (c.hess == PQP_HESSIAN_DENSE) ? h[j] : ((c.hess == PQP_HESSIAN_DIAGONAL && j == i) ? h[j] : 0.0);
      row[j] = v + ((j == i) ? rho : 0.0);
    }
  }
  for (int sl = warp; sl < ns; sl += NW) {
    const int id = row_id(c, sl);
    double* row = K + ts_idx(0, n + sl, 0);
    for (int j = lane; j < n; j += 32) row[j] = c.Bt[(size_t)j * c.ldb + id];
    for (int t = lane; t <= sl; t += 32) row[n + t] = (t == sl) ? -(sl < ne ? mu_eq : mu_in) : 0.0;
  }
  __syncthreads();
  const int ldk = (n + c.cap + 2) & ~1;
  tsym_sweep_invert(c, K, c.kws, ldk, m); // panels: 8 x ldk doubles at the start of the fallback workspace
  if (threadIdx.x == 0) c.kkt_dirty = 0;
  __syncthreads();
}

__device__ __noinline__ void kkt_solve(Ctx& c, const double* b1, const double* b2, double* ox, double* os)
{
  PQP_VECS(c);
  const int n = c.n, ns = c.ns, m = n + ns, ne = c.ne;
  const int ldk = (n + c.cap + 2) & ~1;
  double* const r = c.kws + (size_t)8 * ldk;   // right-hand side
  double* const y = r + ldk;                   // solution
  double* const part = y + ldk;                // NW x m partial vectors of the symmetric mat-vec
  _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) r[j] = b1[j];
  _Pragma("unroll 1") for (int t = threadIdx.x; t < ns; t += NT) r[n + t] = b2[t];
  double* const saved = c.scratch;
  __syncthreads();
  if (threadIdx.x == 0) c.scratch = part;
  __syncthreads();
  tsym_mv(c, c.Si, r, y, m);
  if (threadIdx.x == 0) c.scratch = saved;
  __syncthreads();
  _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) ox[j] = y[j];
  _Pragma("unroll 1") for (int t = threadIdx.x; t < ns; t += NT) os[t] = y[n + t];
  __syncthreads();
  // B^T lam for kkt_residual (v_ctdz), as the dual-block path leaves it
  _Pragma("unroll 1") for (int id = threadIdx.x; id < c.ldb; id += NT) {
    double lam = 0.0;
    if (id < ne) {
      lam = os[id];
    } else if (id < c.m) {
      const int sl = c.cons_slot[id - ne];
      if (sl >= 0) lam = os[sl];
    }
    c.kt[id] = lam;
  }
  __syncthreads();
  bt_dot(c, c.kt, nullptr, v_t2, nullptr, 1.0, v_ctdz, nullptr);
}
#endif

// Solve K [ox; os] = [b1; b2],  K = [P B^T; B -Dlt], with the explicit block
// inverses:  t = P^-1 b1;  lam = S^-1 (B t - b2);  x = P^-1 (b1 - B^T lam).
// In place allowed (ox == b1, os == b2). Replaces Ldlt::solve_in_place
// (ldlt.hpp:767-782).
__device__ __noinline__ void solve_kkt(Ctx& c, const double* b1, const double* b2, double* ox, double* os, double rho, double mu_eq, double mu_in)
{
  PQP_VECS(c);
  const int ns = c.ns, n = c.n, ne = c.ne;
#ifdef PQP_BIG
  if (c.kkt_mode) { // block-uniform: written by thread 0 behind a barrier
    if (c.kkt_dirty) kkt_factor(c, rho, mu_eq, mu_in);
    kkt_solve(c, b1, b2, ox, os);
    return;
  }
#else
  (void)rho;
  (void)mu_eq;
  (void)mu_in;
#endif
  if (ns == 0) {
    apply_Pinv(c, b1, v_t1);
    _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) {
      ox[j] = v_t1[j];
      v_ctdz[j] = 0.0;
    }
    __syncthreads();
    return;
  }
  apply_Pinv(c, b1, v_t1);
  // B t for every row of [A_s; C_s] at once (Bt = B^T), then pick the slots
  bt_axpy(c, v_t1, c.kt);
  _Pragma("unroll 1") for (int s = threadIdx.x; s < ns; s += NT) v_s1[s] = c.kt[row_id(c, s)] - b2[s];
  __syncthreads();
  tsym_mv(c, c.Si, v_s1, os, ns);
#if !defined(PQP_BIG) && PQP_BTLAM_AXPY
  // t2 = b1 - B^T lam in AXPY form over the ACTIVE rows only: the n_eq rows of A_s, then row slot_cons[s] of C_s for every
  // inequality slot s, weighted by lam[s] (one pass over ~n_s rows of n doubles instead of the dot form over all n rows of
  // Bt, n_eq + n_in doubles each, with its transpose-reductions). B^T lam itself is kept (v_ctdz) for kkt_residual.
  if (!c.box) {
    axpy_pass2(c, c.As, ne, c.Cs, n, c.slot_cons, 0, ns, os, n, v_t2, b1, -1.0, v_ctdz); // (the SCALED rows: c.Am / c.Cm are the model)
    apply_Pinv(c, v_t2, ox);
    return;
  }
#endif
  // lam scattered to constraint order (zero on inactive rows), then t2 = b1 - B^T lam.
  // B^T lam itself is kept (v_ctdz): kkt_residual needs exactly this product for the x block.
  _Pragma("unroll 1") for (int id = threadIdx.x; id < c.ldb; id += NT) {
    double lam = 0.0;
    if (id < ne) {
      lam = os[id];
    } else if (id < c.m) {
      const int s = c.cons_slot[id - ne];
      if (s >= 0) lam = os[s];
    }
    c.kt[id] = lam;
  }
  __syncthreads();
  bt_dot(c, c.kt, nullptr, v_t2, b1, -1.0, v_ctdz, nullptr);
  apply_Pinv(c, v_t2, ox);
}

// Append dual slot s == c.ns (already registered in slot_cons) with proximal
// parameter mu: bordering of the explicit inverse
//   w = S^-1 g, delta = (b.P^-1 b + mu) - g.w,
//   S^-1 <- [S^-1 + w w^T/delta, -w/delta; -w^T/delta, 1/delta].
// Replaces Ldlt::insert_block_at (ldlt.hpp:431-475, modify.hpp:131-264).
__device__ __noinline__ void rebuild_Si_from_G(Ctx& c, double mu_eq, double mu_in);
// Inequality row `cons` enters as dual slot c.ns. The caller guarantees that every thread is past the previous
// barrier-terminated primitive (c.ns stable, nobody still reads the slot maps); the maps are written here, by thread 0,
// in the phase that gathers the Gram row (which takes the new slot's row id from `cons`, not from the map).
__device__ __noinline__ void insert_slot(Ctx& c, double mu, double mu_eq, int cons)
{
  PQP_VECS(c);
  const int s = c.ns;
  const int cap = c.si_cap;
  if (threadIdx.x == 0 && s + 1 <= c.cap) {
    c.slot_cons[s] = cons;
    c.cons_slot[cons] = s;
  }
#ifdef PQP_BIG
  if (c.kkt_mode) { // fallback: only the slot count moves; K^-1 is re-formed before the next solve
    __syncthreads();
    if (threadIdx.x == 0) {
      c.ns = s + 1;
      c.kkt_dirty = 1;
    }
    __syncthreads();
    return;
  }
#endif
  if (s + 1 > cap) { // does not fit the shared-memory S^-1: hand the QP to the generic kernel
    if (threadIdx.x == 0) c.overflow = 1;
    __syncthreads();
    return;
  }
  {
    // Gram row of the new slot against slots 0..s: a gather from G
    const int ids = c.ne + cons;
    _Pragma("unroll 1") for (int j = threadIdx.x; j <= s; j += NT) {
      const int idj = (j == s) ? ids : row_id(c, j);
      v_s3[j] = c.G[(size_t)max(ids, idj) * c.ldb + min(ids, idj)];
    }
    __syncthreads();
  }
  // five barriers per insertion (eight before: the kernel pays ~1.8 k cycles of dependent latency per barrier interval
  // here, 350 of a cfg-2 QP's 1600 intervals were insertions): g.w comes out of the mat-vec's final phase, the
  // border, the diagonal and the slot count are written in the phase that forms w / delta
  double delta = v_s3[s] + mu;
  if (s > 0) {
    delta -= tsym_mv(c, c.Si, v_s3, v_s1, s, true); // w = S^-1 g in v_s1, g.w returned (no trailing barrier: see below)
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
    _Pragma("unroll 1") for (int j = threadIdx.x; j < s; j += NT) { // w_j was written by this very thread
      const double wj = v_s1[j] * dinv;
      v_s2[j] = wj;
      ts_put(c.Si, cap, s, j, -wj);
    }
  }
  if (threadIdx.x == 0) {
    c.Si[ts_idx(cap, s, s)] = 1.0 / delta; // row s is outside the order-s update below
    c.ns = s + 1;
  }
  __syncthreads(); // w, w / delta complete; also the barrier c.red needed after the fused reduction
  if (s > 0) PQP_TSYM_RANK1(c, c.Si, v_s1, v_s2, s);
}

// Remove dual slot k (k >= ne): Schur complement of the explicit inverse,
//   S'^-1 = T - q q^T / T_kk   (T = S^-1 without row/column k, q = column k),
// then the last slot takes the place of k (slot order carries no meaning).
// Replaces Ldlt::delete_at (ldlt.hpp:340-387).
// Append kcnt (2..4) dual slots at once: slots s0 .. s0 + kcnt - 1 (s0 == c.ns) are already registered. Block
// bordering of the explicit inverse with Gn = Gram columns of the new rows against the old slots, D = their own Gram
// block + mu I:   W = S^-1 Gn,   Dl = D - Gn^T W,
//   S^-1 <- [ S^-1 + W Dl^-1 W^T,  -W Dl^-1 ;  -Dl^-1 W^T,  Dl^-1 ]
// kcnt read-only mat-vecs and ONE rank-4 pass over S^-1 instead of kcnt mat-vecs and kcnt rank-1 passes (a
// read-modify-write pass costs twice a mat-vec on the stored triangle). The k x k Schur block is inverted
// redundantly by every thread; a pivot that has cancelled sends the whole group to the re-formation from G.
__device__ __noinline__ void insert_block(Ctx& c, int kcnt, double mu, double mu_eq)
{
  PQP_VECS(c);
  const int s0 = c.ns, cap = c.si_cap, ldv = c.uv_ld;
#ifdef PQP_BIG
  if (c.kkt_mode) {
    __syncthreads();
    if (threadIdx.x == 0) {
      c.ns = s0 + kcnt;
      c.kkt_dirty = 1;
    }
    __syncthreads();
    return;
  }
#endif
  if (s0 + kcnt > cap) { // does not fit the shared-memory S^-1: hand the QP to the generic kernel
    if (threadIdx.x == 0) c.overflow = 1;
    __syncthreads();
    return;
  }
  double* T = c.Si;
  // W columns are parked in slot-sized vectors that are free during an active-set change (rhs / err / step of the solve)
  double* const Wst[4] = { v_rs, v_es, v_ds, v_s2 };
  double dl[4][4]; // D - Gn^T W, identity padded
#pragma unroll
  for (int p = 0; p < 4; ++p) {
#pragma unroll
    for (int q = 0; q < 4; ++q) dl[p][q] = (p == q) ? 1.0 : 0.0;
  }
  double dscale[4] = { 1.0, 1.0, 1.0, 1.0 };
  for (int a = 0; a < kcnt; ++a) {
    const int ids = row_id(c, s0 + a);
    _Pragma("unroll 1") for (int j = threadIdx.x; j <= s0 + a; j += NT) {
      const int idj = row_id(c, j);
      v_s3[j] = c.G[(size_t)max(ids, idj) * c.ldb + min(ids, idj)];
    }
    __syncthreads();
    double part[4] = { 0.0, 0.0, 0.0, 0.0 };
    double dummy[1] = { 0.0 };
    if (s0 > 0) {
      tsym_mv(c, T, v_s3, v_s1, s0);
      _Pragma("unroll 1") for (int j = threadIdx.x; j < s0; j += NT) {
        const double wj = v_s1[j];
        Wst[a][j] = wj;
      }
      __syncthreads();
      _Pragma("unroll 1") for (int j = threadIdx.x; j < s0; j += NT) {
        const double gj = v_s3[j];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (b <= a) part[b] = fma(gj, Wst[b][j], part[b]);
        }
      }
      block_reduce<4, 0>(c, part, dummy);
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (b <= a) {
        const double dab = v_s3[s0 + b] + ((b == a) ? mu : 0.0);
        const double v = dab - part[b];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if ((p == a && q == b) || (p == b && q == a)) dl[p][q] = v;
          }
        }
        if (b == a) dscale[a] = dab;
      }
    }
    __syncthreads();
  }
  // Dl^-1 by Gauss-Jordan (SPD); a cancelled pivot means the bordering has no accuracy left
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    bad = bad || !(dl[k][k] > 1e-13 * dscale[k]);
    const double inv = 1.0 / dl[k][k];
#pragma unroll
    for (int q = 0; q < 4; ++q) dl[k][q] = (q == k) ? inv : dl[k][q] * inv;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (p != k) {
        const double f = dl[p][k];
#pragma unroll
        for (int q = 0; q < 4; ++q) dl[p][q] = (q == k) ? -f * inv : fma(-f, dl[k][q], dl[p][q]);
      }
    }
  }
  if (bad) { // block-uniform: dl comes out of block reductions
    if (threadIdx.x == 0) c.ns = s0 + kcnt;
    __syncthreads();
    rebuild_Si_from_G(c, mu_eq, mu);
    return;
  }
  double* const U = v_scratch;
  double* const V = v_scratch + 4 * ldv;
  _Pragma("unroll 1") for (int i = threadIdx.x; i < s0; i += NT) {
    double w[4], u[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) w[a] = (a < kcnt) ? Wst[a][i] : 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) u[a] = (a < kcnt) ? (w[0] * dl[0][a] + w[1] * dl[1][a] + w[2] * dl[2][a] + w[3] * dl[3][a]) : 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      V[a * ldv + i] = w[a];
      if (a < kcnt) ts_put(T, cap, s0 + a, i, -u[a]);
    }
#ifdef PQP_BIG
#pragma unroll
    for (int a = 0; a < 4; ++a) U[a * ldv + i] = u[a];
#else
    reinterpret_cast<double2*>(U)[2 * i] = make_double2(u[0], u[1]);
    reinterpret_cast<double2*>(U)[2 * i + 1] = make_double2(u[2], u[3]);
#endif
  }
  if (threadIdx.x == 0) {
    for (int a = 0; a < kcnt; ++a) {
      for (int b = 0; b <= a; ++b) ts_put(T, cap, s0 + a, s0 + b, dl[a][b]);
    }
  }
  __syncthreads();
  if (s0 > 0) tsym_rank4(c, T, U, V, ldv, s0);
  if (threadIdx.x == 0) c.ns = s0 + kcnt;
  __syncthreads();
}

// Storage part of a deletion: the last slot takes the place of slot k (slot order carries no meaning), the freed last
// row / column is cleared, the slot maps are updated. Rows / columns of k must already be decoupled from the rest.
__device__ __noinline__ void drop_slot(Ctx& c, int k)
{
  PQP_VECS(c);
  const int ns = c.ns, cap = c.si_cap, L = ns - 1;
  double* T = c.Si;
  __syncthreads();
#ifdef PQP_BIG
  const bool storage = !c.kkt_mode; // fallback: K^-1 is re-formed from scratch, only the maps move
#else
  const bool storage = true;
#endif
  if (storage) {
    _Pragma("unroll 1") for (int i = threadIdx.x; i < ns; i += NT) v_s1[i] = ts_get(T, cap, i, L);
    __syncthreads();
    _Pragma("unroll 1") for (int i = threadIdx.x; i < ns; i += NT) {
      if (k != L) {
        if (i == k)
          T[ts_idx(cap, k, k)] = v_s1[L];
        else if (i != L)
          ts_put(T, cap, i, k, v_s1[i]);
      }
      ts_put(T, cap, L, i, 0.0);
    }
  }
  if (threadIdx.x == 0) {
#ifdef PQP_BIG
    if (c.kkt_mode) c.kkt_dirty = 1;
#endif
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

__device__ __noinline__ void delete_slot(Ctx& c, int k)
{
  PQP_VECS(c);
  const int ns = c.ns, cap = c.si_cap, L = ns - 1;
  double* T = c.Si;
#ifdef PQP_BIG
  if (c.kkt_mode) {
    drop_slot(c, k);
    return;
  }
#endif
  // Three barriers per deletion (seven before). The last slot L moves into the freed position k BEFORE the rank-1
  // pass: its updated column p'_i = T[i][L] + q_i q_L sinv is formed explicitly (same fma as the pass would apply to the
  // stored element (L, i)), rows / columns k of the update vectors are zero, and the pass runs on order ns - 1.
  _Pragma("unroll 1") for (int i = threadIdx.x; i < ns; i += NT) {
    v_s1[i] = ts_get(T, cap, i, k); // q: column k
    v_s2[i] = ts_get(T, cap, i, L); // p: column of the last slot
  }
  __syncthreads();
  const double sinv = -1.0 / v_s1[k];
  const double qL = (k == L) ? 0.0 : v_s1[L];
  _Pragma("unroll 1") for (int i = threadIdx.x; i < ns; i += NT) {
    const double q = (i == k) ? 0.0 : v_s1[i]; // row / column k are dropped
    const double v = q * sinv;
    const double pp = fma(qL, v, v_s2[i]);     // element (L, i) after the update
    v_s2[i] = q;
    v_s3[i] = v;
    if (k != L) {
      if (i == L)
        T[ts_idx(cap, k, k)] = pp;
      else if (i != k)
        ts_put(T, cap, i, k, pp);
    }
    ts_put(T, cap, L, i, 0.0);
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
  PQP_TSYM_RANK1(c, T, v_s2, v_s3, L);
}

// Remove up to four dual slots at once (ks[0 .. kcnt), all >= ne, distinct): block form of the Schur complement,
//   S'^-1 = T_RR - Q (T_KK)^-1 Q^T,   Q = T[:, K],
// as ONE rank-4 pass over S^-1 instead of kcnt rank-1 passes (deletions come in groups: most of the cost of a
// deletion is streaming the stored triangle). The k x k block is inverted redundantly by every thread (it is a
// principal block of an SPD matrix: Gauss-Jordan without pivoting), rows / columns K stay untouched (zero rows of U,
// zero columns of V) and are dropped afterwards, last slot into freed position, highest position first.
__device__ __noinline__ void delete_block(Ctx& c, int kcnt, int k0, int k1, int k2, int k3)
{
  PQP_VECS(c);
  const int ns = c.ns, cap = c.si_cap, ldv = c.uv_ld;
  double* T = c.Si;
  double* const U = v_scratch;
  double* const V = v_scratch + 4 * ldv;
  const int ks[4] = { k0, k1, k2, k3 };
#ifdef PQP_BIG
  if (c.kkt_mode) { // maps only, highest position first
    int ord[4] = { k0, k1, k2, k3 };
    for (int a = 0; a < 4; ++a) {
      if (a >= kcnt) ord[a] = -1;
    }
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3 - a; ++b) {
        if (ord[b] < ord[b + 1]) {
          const int t = ord[b];
          ord[b] = ord[b + 1];
          ord[b + 1] = t;
        }
      }
    }
    for (int a = 0; a < kcnt; ++a) drop_slot(c, ord[a]);
    return;
  }
#endif
  _Pragma("unroll 1") for (int i = threadIdx.x; i < ns; i += NT) {
#pragma unroll
    for (int a = 0; a < 4; ++a) V[a * ldv + i] = (a < kcnt) ? ts_get(T, cap, i, ks[a]) : 0.0;
  }
  __syncthreads();
  double w[4][4]; // becomes -(T_KK)^-1, identity padded
#pragma unroll
  for (int p = 0; p < 4; ++p) {
#pragma unroll
    for (int q = 0; q < 4; ++q) w[p][q] = (p < kcnt && q < kcnt) ? V[q * ldv + ks[p]] : ((p == q) ? 1.0 : 0.0);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { // in-place Gauss-Jordan inverse
    const double inv = 1.0 / w[k][k];
#pragma unroll
    for (int q = 0; q < 4; ++q) w[k][q] = (q == k) ? inv : w[k][q] * inv;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (p != k) {
        const double f = w[p][k];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[p][q] = (q == k) ? -f * inv : fma(-f, w[k][q], w[p][q]);
      }
    }
  }
  __syncthreads(); // every thread has read the block before rows K of V are cleared
  _Pragma("unroll 1") for (int i = threadIdx.x; i < ns; i += NT) {
    bool inK = false;
#pragma unroll
    for (int a = 0; a < 4; ++a) inK = inK || (a < kcnt && i == ks[a]);
    double q[4], u[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) q[a] = V[a * ldv + i];
#pragma unroll
    for (int a = 0; a < 4; ++a) u[a] = inK ? 0.0 : -(q[0] * w[0][a] + q[1] * w[1][a] + q[2] * w[2][a] + q[3] * w[3][a]);
    if (inK) {
#pragma unroll
      for (int a = 0; a < 4; ++a) V[a * ldv + i] = 0.0;
    }
#ifdef PQP_BIG
#pragma unroll
    for (int a = 0; a < 4; ++a) U[a * ldv + i] = u[a];
#else
    reinterpret_cast<double2*>(U)[2 * i] = make_double2(u[0], u[1]);
    reinterpret_cast<double2*>(U)[2 * i + 1] = make_double2(u[2], u[3]);
#endif
  }
  __syncthreads();
  tsym_rank4(c, T, U, V, ldv, ns);
  // drop the slots, highest position first (a lower position is never the "last slot" of a later drop)
  int order[4] = { k0, k1, k2, k3 };
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (a >= kcnt) order[a] = -1;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int b = 0; b < 3 - a; ++b) {
      if (order[b] < order[b + 1]) {
        const int t = order[b];
        order[b] = order[b + 1];
        order[b + 1] = t;
      }
    }
  }
  for (int a = 0; a < kcnt; ++a) drop_slot(c, order[a]);
}

// S^-1 from the cached Gram matrix with the given proximal parameters
// (S = Dlt + G): gather + sweep inversion. Replaces
// Ldlt::diagonal_update_clobber_indices (ldlt.hpp:516-570) used by mu_update
// (solver.hpp:130-169).
__device__ __noinline__ void rebuild_Si_from_G(Ctx& c, double mu_eq, double mu_in)
{
  PQP_VECS(c);
  const int ns = c.ns, cap = c.si_cap;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifdef PQP_BIG
  (void)cap;
  if (c.kkt_mode) { // fallback: nothing to maintain here, K^-1 is re-formed before the next solve
    __syncthreads();
    if (threadIdx.x == 0) {
      c.si_valid = 1;
      c.kkt_dirty = 1;
    }
    __syncthreads();
    return;
  }
  for (int s = warp; s < ns; s += NW) {
    const int ids = row_id(c, s);
    double* row = c.Si + ts_idx(0, s, 0);
    for (int t = lane; t <= s; t += 32) {
      const int idt = row_id(c, t);
      row[t] = c.G[(size_t)max(ids, idt) * c.ldb + min(ids, idt)] + ((t == s) ? (s < c.ne ? mu_eq : mu_in) : 0.0);
    }
  }
  __syncthreads();
  tsym_sweep_invert(c, c.Si, v_scratch, c.uv_ld, ns);
  if (threadIdx.x == 0) c.si_valid = 1;
  __syncthreads();
  return;
#else
  const int nb = (ns + 31) >> 5;
  for (int bi = 0; bi < nb; ++bi) {
    const int rst = ts_rows(cap, bi);
    for (int r = warp; r < rst; r += NW) {
      const int s = 32 * bi + r;
      const int ids = (s < ns) ? row_id(c, s) : 0;
      for (int bj = 0; bj <= bi; ++bj) {
        const int t = 32 * bj + lane;
        double v = 0.0;
        if (s < ns && t < ns) {
          const int idt = row_id(c, t);
          v = c.G[(size_t)max(ids, idt) * c.ldb + min(ids, idt)] + ((t == s) ? (s < c.ne ? mu_eq : mu_in) : 0.0);
        }
        c.Si[ts_tile(cap, bi, bj) + r * TS_LD + lane] = v;
      }
    }
  }
  __syncthreads();
  tsym_sweep_invert(c, c.Si, v_scratch, c.uv_ld, ns);
  if (threadIdx.x == 0) c.si_valid = 1;
  __syncthreads();
#endif
}

// P^-1 = (Hs + rho I)^-1, explicit: swept in the (free) S^-1 tile storage, then
// written to the L2 workspace as a full square for the AXPY passes. Replaces
// the x-block part of Ldlt::factorize (ldlt.hpp:718-744).
__device__ __noinline__ void build_Pi(Ctx& c, double rho)
{
  PQP_VECS(c);
  const int n = c.n, cap = c.si_cap;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifdef PQP_BIG
  (void)cap;
  if (c.hess != PQP_HESSIAN_DENSE) {
    _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) c.d1inv[j] = 1.0 / (((c.hess == PQP_HESSIAN_DIAGONAL) ? c.Hs[(size_t)j * n + j] : 0.0) + rho);
    __syncthreads();
    return;
  }
  {
    double* const work = c.Si; // packed region sized for max(n, capacity)
    for (int i = warp; i < n; i += NW) {
      double* row = work + ts_idx(0, i, 0);
      const double* h = c.Hs + (size_t)i * n;
      for (int j = lane; j <= i; j += 32) row[j] = h[j] + ((j == i) ? rho : 0.0);
    }
    __syncthreads();
    tsym_sweep_invert(c, work, v_scratch, c.uv_ld, n);
    for (int i = warp; i < n; i += NW) {
      for (int j = lane; j < c.ldn; j += 32) c.Pi[(size_t)i * c.ldn + j] = (j < n) ? ts_get(work, 0, i, j) : 0.0;
    }
    __syncthreads();
    return;
  }
#else
  const int nb = (n + 31) >> 5;
  double* const work = c.Si;
  for (int bi = 0; bi < nb; ++bi) {
    const int rst = ts_rows(cap, bi);
    for (int r = warp; r < rst; r += NW) {
      const int i = 32 * bi + r;
      for (int bj = 0; bj <= bi; ++bj) {
        const int j = 32 * bj + lane;
        double v = 0.0;
        if (i < n && j < n) v = c.Hs[(size_t)max(i, j) * n + min(i, j)] + ((j == i) ? rho : 0.0); // lower triangle of H_s, mirrored
        work[ts_tile(cap, bi, bj) + r * TS_LD + lane] = v;
      }
    }
  }
  __syncthreads();
  tsym_sweep_invert(c, work, v_scratch, c.uv_ld, n);
  for (int i = warp; i < n; i += NW) {
    for (int j = lane; j < c.ldn; j += 32) c.Pi[(size_t)i * c.ldn + j] = (j < n) ? ts_get(work, cap, i, j) : 0.0;
  }
  // leave the whole region clean for S^-1 (entries outside the live order must be zero)
  const int tot = ts_extent(cap, cap);
  __syncthreads();
  _Pragma("unroll 1") for (int e = threadIdx.x; e < tot; e += NT) work[e] = 0.0;
  __syncthreads();
#endif
}

// Bt = [A_s; C_s]^T (n x ldb) in the L2 workspace
__device__ __noinline__ void build_Bt(Ctx& c)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = c.n, ne = c.ne, nr = c.ne + c.ni;
  for (int r = warp; r < nr; r += NW) {
    const double* row = (r < ne) ? c.As + (size_t)r * n : c.Cs + (size_t)(r - ne) * n;
    for (int j = lane; j < n; j += 32) c.Bt[(size_t)j * c.ldb + r] = row[j];
  }
  if (c.box) {
    // box rows are the scaled unit vectors i_s[k] e_k (solver.hpp:74-81): materialised, so that
    // every product below treats the n_in + n inequality rows alike
    PQP_VECS(c);
    for (int j = warp; j < n; j += NW) {
      for (int k = lane; k < n; k += 32) c.Bt[(size_t)j * c.ldb + nr + k] = (k == j) ? v_is[j] : 0.0;
    }
  }
  if (c.ldb > c.m) {
    _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) c.Bt[(size_t)j * c.ldb + c.m] = 0.0;
  }
  __syncthreads();
}

// W = P^-1 B^T and G = B W for all rows of B = [A_s; C_s] (G is symmetric; it is
// read as G[max][min]). Depends on H_s, rho, A_s, C_s only: once per QP.
__device__ __noinline__ void build_G(Ctx& c)
{
  const int n = c.n, m = c.m;
#ifdef PQP_BIG
  if (c.hess != PQP_HESSIAN_DENSE) { // W = P^-1 B^T with a diagonal P^-1: a row scaling of Bt
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = warp; k < n; k += NW) {
      const double d = c.d1inv[k];
      for (int j = lane; j < c.ldb; j += 32) c.W[(size_t)k * c.ldb + j] = d * c.Bt[(size_t)k * c.ldb + j];
    }
    __syncthreads();
  } else
#endif
#ifndef PQP_BIG
  {
    // W never touches the L2 workspace: it is formed in two column halves inside the S^-1 region of shared memory
    // (free until the dual block is built) and consumed at once. 120 KB less per CTA at cfg 2: 296 workspaces
    // then fit the L2 together with the streaming model data.
    const int half = (((m + 1) >> 1) + 7) & ~7;
    if ((long long)n * half <= (long long)ts_extent(c.si_cap, c.si_cap)) {
      double* const Wsm = c.Si;
      for (int c0 = 0; c0 < m; c0 += half) {
        const int nc_ = min(half, m - c0);
        gemm_tn(c.Pi, c.ldn, c.Bt + c0, c.ldb, n, n, nc_, Wsm, half, false);     // W_h = Pi^T Bt[:, c0 : c0 + nc_]
        gemm_tn(c.Bt, c.ldb, Wsm, half, n, m, nc_, c.G + c0, c.ldb, true, c0);   // G[:, c0 : c0 + nc_] = Bt^T W_h (lower block triangle)
      }
      const int tot = ts_extent(c.si_cap, c.si_cap); // the region must be clean for S^-1 (entries outside the live order are zero)
      _Pragma("unroll 1") for (int e = threadIdx.x; e < tot; e += NT) Wsm[e] = 0.0;
      __syncthreads();
      return;
    }
  }
#endif
  gemm_tn(c.Pi, c.ldn, c.Bt, c.ldb, n, n, m, c.W, c.ldb, false); // W = Pi^T Bt (Pi symmetric)
  gemm_tn(c.Bt, c.ldb, c.W, c.ldb, n, m, m, c.G, c.ldb, true);   // G = Bt^T W, lower block triangle
}

// (Re)build the dual block for the slots 0..ns_target-1 currently registered:
// gather from G, then one sweep inversion. Used for the first factorisation
// (equality rows only, helpers.hpp:241-285) and by refactorize (solver.hpp:40-87).
__device__ __noinline__ void build_dual_block(Ctx& c, int ns_target, double mu_eq, double mu_in)
{
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
__device__ __noinline__ double kkt_residual(const Ctx& c, const Scal& sc, bool first)
{
  PQP_VECS(c);
  const int n = c.n, ne = c.ne, ns = c.ns;
  apply_H(c, c.Hs, v_dx, v_hdx);                                             // H dx (H symmetric)
  bt_axpy(c, v_dx, v_adx);                                                 // [A dx; C dx] (adx, cdx contiguous)
  // A^T dy + C_J^T dz_J is B^T lam of the solve that produced (dx, ds) (first call) or of the
  // refinement step just added to it: no third pass over the constraint rows
  double m = 0;
  _Pragma("unroll 1") for (int j = threadIdx.x; j < n; j += NT) {
    const double at = first ? v_ctdz[j] : v_atdy[j] + v_ctdz[j];
    v_atdy[j] = at;
    double e = v_rx[j] - (v_hdx[j] + sc.rho * v_dx[j] + at); // atdy = A^T dy + C_J^T dz_J
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
  // P^-1, W and G depend on (H_s, rho, A_s, C_s) only and carry no accumulated
  // update error: only the incrementally updated S^-1 is rebuilt
  build_dual_block(c, ns_target, sc.mu_eq, sc.mu_in);
  sc.factor_fresh = true;
}

// solver.hpp:408-541
__device__ __noinline__ void iterative_solve(Ctx& c, Scal& sc, const pqp_settings& S, double eps)
{
  PQP_VECS(c);
  for (int pass = 0; pass < 3; ++pass) {
    int it = 0, it_stab = 0;
    long long tp = PROF_T0();
    solve_kkt(c, v_rx, v_rs, v_dx, v_ds, sc.rho, sc.mu_eq, sc.mu_in);
    PROF_ADD(PH_SOLVE, tp);
    tp = PROF_T0();
    double err = kkt_residual(c, sc, true);
    PROF_ADD(PH_RESID, tp);
    ++it;
    double prev = err;
#ifdef PQP_BIG
    // Diagonal / zero Hessians (see pqp_solver_body.inl, iterative_solve): P^-1 reaches 1 / rho where H_jj = 0, the dual
    // block is then badly conditioned and its explicit inverse less accurate than the reference's LDL^T; the refinement
    // is driven down to 1e-10 instead of stopping at the inner tolerance so that the Newton steps stay as good.
    const double eps_ref = (c.hess != PQP_HESSIAN_DENSE) ? fmin(eps, 1e-10) : eps;
#else
    const double eps_ref = eps;
#endif
    while (err >= eps_ref) {
      if (it >= S.nb_iterative_refinement) break;
      ++it;
      tp = PROF_T0();
      solve_kkt(c, v_ex, v_es, v_ex, v_es, sc.rho, sc.mu_eq, sc.mu_in);
      _Pragma("unroll 1") for (int j = threadIdx.x; j < c.n; j += NT) v_dx[j] += v_ex[j];
      _Pragma("unroll 1") for (int s = threadIdx.x; s < c.ns; s += NT) v_ds[s] += v_es[s];
      __syncthreads();
      PROF_ADD(PH_SOLVE, tp);
      tp = PROF_T0();
      err = kkt_residual(c, sc, false);
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
#ifdef PQP_BIG
    // still not solved with a freshly re-formed dual block: this QP is beyond the S^-1 path (see kkt_factor); the rest
    // of it is solved through the inverse of the whole KKT matrix (err is block-uniform: it comes out of a reduction)
    if (pass <= 1 && err >= fmax(eps, S.eps_refact) && !c.kkt_mode) {
      __syncthreads();
      if (threadIdx.x == 0) {
        c.kkt_mode = 1;
        c.kkt_dirty = 1;
      }
      __syncthreads();
      continue;
    }
#endif
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
  for (int k = ndel - 1; k >= 0;) { // from the last slot to the first, four at a time
#if defined(PQP_BIG) || PQP_TILE_BLOCK
    const int cnt = min(4, k + 1);
#else
    const int cnt = 1; // shared-memory S^-1 (tile kernel): a pass is cheap, the block form only adds code (measured: -6 % at cfg 2)
#endif
    if (cnt == 1) {
      delete_slot(c, c.ne + c.list1[k]);
    } else {
      delete_block(c, cnt, c.ne + c.list1[k], c.ne + c.list1[k - 1], cnt > 2 ? c.ne + c.list1[k - 2] : -1, cnt > 3 ? c.ne + c.list1[k - 3] : -1);
    }
    k -= cnt;
  }
  PROF_ADD(PH_DELETE, tp);
  tp = PROF_T0();
  int nadd = block_compact(c, c.nc, c.list1, [&](int i) { return (c.act_up[i] || c.act_low[i]) && c.cons_slot[i] < 0; });
  // A bordering step costs two passes over S^-1; a rebuild from G costs (ns+nadd)/4 block
  // sweeps of the same size. Many simultaneous additions (the first Newton steps) are
  // therefore registered at once and S^-1 is re-formed from the Gram matrix.
  // c.ns / c.si_valid decide a block-uniform branch: every thread reads them ONCE, and thread 0 may only change
  // them after a barrier (a warp that is late here - e.g. on an instruction-cache miss - would otherwise see the
  // new slot count, take the other branch and desynchronise the CTA's barriers)
  const int base = c.ns;
  const bool rebuild = !c.si_valid || (nadd >= 8 && 2 * nadd >= (base + nadd + 3) / 4);
  if (rebuild) {
    if (base + nadd > c.si_cap) {
      if (threadIdx.x == 0) c.overflow = 1;
      __syncthreads();
    } else {
      _Pragma("unroll 1") for (int k = threadIdx.x; k < nadd; k += NT) {
        const int cons = c.list1[k];
        c.slot_cons[base + k] = cons;
        c.cons_slot[cons] = base + k;
      }
      __syncthreads();
      if (threadIdx.x == 0) c.ns = base + nadd;
      __syncthreads();
      rebuild_Si_from_G(c, sc.mu_eq, sc.mu_in);
    }
  } else {
    for (int k = 0; k < nadd;) { // four at a time where S^-1 is streamed from L2 / HBM (big variant)
#if defined(PQP_BIG) || PQP_TILE_BLOCK
      const int cnt = min(4, nadd - k);
#else
      const int cnt = 1;
#endif
      if (cnt == 1) { // (registers the slot itself: no barrier of its own for the maps)
        insert_slot(c, sc.mu_in, sc.mu_eq, c.list1[k]);
      } else {
        const int s0 = c.ns;
        __syncthreads();
        if (threadIdx.x < cnt) {
          const int cons = c.list1[k + threadIdx.x];
          c.slot_cons[s0 + threadIdx.x] = cons;
          c.cons_slot[cons] = s0 + threadIdx.x;
        }
        __syncthreads();
        insert_block(c, cnt, sc.mu_in, sc.mu_eq);
      }
      if (c.overflow) break;
      k += cnt;
    }
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
    apply_H(c, c.Hs, v_x, v_t1);
    // A^T y and C^T z from one pass over Bt: coefficient vectors [y; 0] and [0; z]
    _Pragma("unroll 1") for (int id = threadIdx.x; id < c.ldb; id += NT) {
      c.kt[id] = (id < ne) ? v_y[id] : 0.0;
      c.kt2[id] = (id >= ne && id < ne + ni) ? v_z[id - ne] : 0.0;
    }
    __syncthreads();
    bt_dot(c, c.kt, c.kt2, v_t2, nullptr, 1.0, nullptr, v_t3);
  }
  if (primal) bt_axpy(c, v_x, v_se); // [A x; C x] (se, rup contiguous)
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
    axpy_pass(c, c.Am, n, ne, v_se, n, v_ex, nullptr, 1.0);
    axpy_pass(c, c.Cm, n, ni, v_si, n, v_ex, v_ex, 1.0);
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
#ifndef PQP_SOLVE_ONE_ATTR
#define PQP_SOLVE_ONE_ATTR
#endif
__device__ PQP_SOLVE_ONE_ATTR void solve_one(Ctx& c, const PqpSolveArgs& A, int q)
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
      c.As = const_cast<double*>(Asg);
    }
    __syncthreads();
    _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_gs[j] = P.gs[(size_t)q * n + j];
#ifdef PQP_BIG
    // big layout without box constraints: the unscaled b, u, l (one use per outer iteration) are read where they lie
    // (c.b / c.u / c.l were pointed at the model arrays of this QP by the kernel body): no staging, no arena space
    const bool ext_bounds = !c.box;
#else
    const bool ext_bounds = false;
#endif
    _Pragma("unroll 1") for (int j = tid; j < ne; j += NT) {
      v_bs[j] = P.bs[(size_t)q * ne + j];
      if (!ext_bounds) v_b[j] = P.b[(size_t)q * ne + j];
    }
    _Pragma("unroll 1") for (int j = tid; j < nc; j += NT) {
      v_us[j] = P.us[(size_t)q * nc + j];
      v_ls[j] = P.ls[(size_t)q * nc + j];
      if (ext_bounds) {
      } else if (j < ni) {
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
      c.kkt_mode = A.force_kkt ? 1 : 0;
      c.kkt_dirty = c.kkt_mode;
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
  build_Bt(c);
  build_Pi(c, sc.rho);
  PROF_ADD(PH_M1, tph);
  tph = PROF_T0();
  build_G(c);
  PROF_ADD(PH_EQ, tph);
  tph = PROF_T0();
  if (prm.start_mode == PQP_START_EQ_GUESS) {
    build_dual_block(c, ne, sc.mu_eq, sc.mu_in);
  } else {
    // the equality block alone is never used: the first active-set change forms S^-1 for
    // equalities + active inequalities in one inversion
    if (tid == 0) {
      c.ns = ne;
      c.si_valid = 0;
    }
    __syncthreads();
  }
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
        int changed = 0;
        _Pragma("unroll 1") for (int i = tid; i < nc; i += NT) {
          const bool up = v_rup[i] >= 0.0, low = v_si[i] <= 0.0;
          c.act_up[i] = up;
          c.act_low[i] = low;
          changed |= ((up || low) != (c.cons_slot[i] >= 0)) ? 1 : 0;
        }
        // (the barrier doubles as the vote: most late Newton steps keep the active set, and then the two ordered
        // compactions of active_set_change - four barriers - are skipped; c.si_valid is block-uniform here)
        changed = __syncthreads_or(changed);
        if (changed || !c.si_valid) active_set_change(c, sc);
        if (c.overflow) { // S^-1 capacity exceeded: the QP is re-solved by the generic kernel
          expired = true;
          break;
        }
        // q = sum over inactive constraints with z_i != 0 of z_i c_i
        int nq = 0;
        _Pragma("unroll 1") for (int id = tid; id < c.ldb; id += NT) {
          double zq = 0.0;
          if (id >= ne && id < c.m && c.cons_slot[id - ne] < 0) zq = v_z[id - ne];
          c.kt[id] = zq;
          nq |= (zq != 0.0) ? 1 : 0;
        }
        nq = __syncthreads_or(nq);
        if (nq > 0) {
          bt_dot(c, c.kt, nullptr, v_q, nullptr, 1.0, nullptr, nullptr);
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
        _Pragma("unroll 1") for (int j = tid; j < n; j += NT) v_atdy[j] -= v_q[j]; // atdy now holds A^T dy + C^T dz
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
          double dr = v_dual[j] + alpha * (sc.rho * dxj + v_hdx[j] + v_atdy[j]);
          v_dual[j] = dr;
          mx[0] = nanmax(mx[0], fabs(dr));
          const double dxc = dlx[j] * cs;
          mx[3] = nanmax(mx[3], fabs(v_atdy[j] / dxc));
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
      _Pragma("unroll 1") for (int j = tid; j < ne + ni; j += NT) c.kt[j] = 1.0;
      __syncthreads();
      axpy_pass(c, c.Am, n, ne, c.kt, n, v_t1, nullptr, 1.0);
      axpy_pass(c, c.Cm, n, ni, c.kt, n, v_t1, v_t1, 1.0);
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
    apply_H(c, c.Hm, v_t1, v_t2);
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

#ifdef PQP_CPU_EMU
static double* const smem_dyn = emu::dyn_smem;
#else
extern __shared__ __align__(16) double smem_dyn[];
#endif

// FUSED: the feed gate + set-up of the end-to-end path are compiled in (a separate instantiation keeps the
// register allocation of the plain solve kernel untouched)
template<int FUSED>
__device__ __forceinline__ void solve_kernel_body(const PqpSolveArgs& A)
{
  __shared__ Ctx c;
  __shared__ int cur_q;
  __shared__ setupk::FeedArgs feed_args;
  if (FUSED && threadIdx.x == 0) {
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
    c.vec_smem = 1;
    c.pi_smem = 0;
    c.si_cap = L.si_cap;
    c.uv_ld = ((A.d.n > L.si_cap ? A.d.n : L.si_cap) + 2) & ~1;
    c.m = A.d.ne + A.d.nc;
    c.ldb = (c.m + 1) & ~1;
    c.ldn = (A.d.n + 1) & ~1;
    c.overflow = 0;
    double* ws = A.ws + (size_t)blockIdx.x * (size_t)L.ws_doubles;
    auto place = [&](int id) -> double* { return (L.in_smem[id] ? smem_dyn : ws) + L.off[id]; };
    c.n = A.d.n;
    c.ne = A.d.ne;
    c.ni = A.d.ni;
    c.nc = A.d.nc;
    c.box = A.d.box;
    c.hess = A.d.hess;
    c.cap = L.si_cap; // slot-indexed vectors are sized for the shared-memory S^-1
    c.ns = 0;
    c.Pi = place(PA_M1);
    c.As = nullptr;
    c.Bt = place(PA_AS); // the A_s slot of the workspace holds Bt in this layout
    c.Si = place(PA_MS);
    c.G = place(PA_G);
    c.Y = nullptr;
    c.W = place(PA_Y); // the Y slot of the workspace holds W in this layout
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
    c.kt = v + L.voff[V_KT];
    c.kt2 = v + L.voff[V_KT2];
#ifdef PQP_BIG
    // big layout (kind 2): the pass scratch, the reduction scratch and the two coefficient vectors always live in shared
    // memory (absolute offsets), whether or not the vector arena fits there
    c.scratch = smem_dyn + L.voff[V_SCRATCH];
    c.red = smem_dyn + L.voff[V_RED];
    c.kt = smem_dyn + L.voff[V_KT];
    c.kt2 = c.alphas; // shares the line-search breakpoint array (never live together)
    c.vec_smem = L.in_smem[PA_VEC];
    c.kws = c.W;
    c.pf = A.prefetch;
#else
    c.pf = 0;
#endif
    c.kkt_mode = 0;
    c.kkt_dirty = 0;
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
  while (true) {
    if (threadIdx.x == 0) cur_q = atomicAdd(A.counter, 1);
    __syncthreads();
    const int cq = cur_q;
    const int q = A.first + cq; // this launch owns the QPs [first, first + batch)
    __syncthreads();
    if (cq >= A.batch) break;
    if (FUSED && (A.ready || A.fused_setup)) {
      if (!setupk::feed_and_setup(&feed_args, cq, q, smem_dyn)) continue;
    }
    if (!A.p.params[q].active) continue;
#ifdef PQP_BIG
    if (!A.d.box) { // unscaled bounds straight from the model arrays (see solve_one, staging)
      if (threadIdx.x == 0) {
        c.b = A.p.b + (size_t)q * A.d.ne;
        c.u = A.p.u + (size_t)q * A.d.ni;
        c.l = A.p.l + (size_t)q * A.d.ni;
      }
      __syncthreads();
    }
#endif
    solve_one(c, A, q);
  }
  if (A.prof && threadIdx.x == 0) {
    for (int k = 0; k < PH_COUNT; ++k) atomicAdd(reinterpret_cast<unsigned long long*>(A.prof) + k, (unsigned long long)prof_sh[k]);
  }
}

// Plain launch: parameters by value, exactly the kernel the device-resident path has always run.
#ifdef PQP_PLAIN_GC
__global__ void __launch_bounds__(NT, PQP_MIN_CTAS) pqp_solve_kernel(const __grid_constant__ PqpSolveArgs A)
#else
__global__ void __launch_bounds__(NT, PQP_MIN_CTAS) pqp_solve_kernel(PqpSolveArgs A)
#endif
{
  solve_kernel_body<0>(A);
}
// Fused feed: __grid_constant__ lets the non-inlined set-up read the parameters in place.
__global__ void __launch_bounds__(NT, PQP_MIN_CTAS) pqp_solve_kernel_fused(const __grid_constant__ PqpSolveArgs A)
{
  solve_kernel_body<1>(A);
}
