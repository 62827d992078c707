// ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (plain C++17, no Eigen) of the reference's dense LDLT with
// in-place row/column insertion, deletion, diagonal update and rank-r update.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use anything under oracle/.
//
// Follows (file:line under /root/reference/include/proxsuite/linalg/dense):
//   ldlt.hpp:340-387   Ldlt::delete_at
//   ldlt.hpp:389-401   choose_insertion_position
//   ldlt.hpp:431-475   Ldlt::insert_block_at
//   ldlt.hpp:516-570   Ldlt::diagonal_update_clobber_indices
//   ldlt.hpp:580-609   Ldlt::rank_r_update
//   ldlt.hpp:718-744   Ldlt::factorize
//   ldlt.hpp:767-782   Ldlt::solve_in_place
//   factorize.hpp:19-87   compute_permutation / apply_permutation_tri_lower
//   factorize.hpp:91-148  factorize_unblocked_impl (left-looking)
//   update.hpp:221-287    rank_r_update_clobber_w_impl
//   modify.hpp:20-127     delete_rows_and_cols_triangular / ldlt_delete_rows_and_cols_impl
//   modify.hpp:131-264    ldlt_insert_rows_and_cols_impl
//   solve.hpp:17-26       solve_impl
// The reference dispatches factorisation to a recursive blocked variant above
// 32 columns (factorize.hpp:362-370); that variant performs the same
// eliminations in a different association order, so only rounding differs.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace oracle {

using isize = std::int64_t;

// Operation counters used to derive the algorithmic byte/flop model of
// SURVEY.md section 8(d) (streamed-operand model).
struct Counters
{
  double n_factor = 0;        // factorisations
  double factor_m2 = 0;       // sum m^2 over factorisations
  double factor_m3 = 0;       // sum m^3 over factorisations
  double n_solve = 0;         // triangular solve pairs
  double solve_m2 = 0;        // sum m^2 over solves
  double solve_m = 0;         // sum m over solves
  double n_resid = 0;         // KKT residual evaluations
  double resid_nc = 0;        // sum n_c over residual evaluations
  double rank_rt2 = 0;        // sum r * t^2 over rank-r updates
  double rank_chunk_t2 = 0;   // sum ceil(r/4) * t^2
  double rank_rt = 0;         // sum r * t
  double n_insert = 0;        // inserted columns
  double insert_bytes = 0;    // 8*(pos^2/2 + 2 t (t+pos)) per insert call
  double n_delete = 0;        // deleted columns
  double delete_t2 = 0;       // sum t^2 over delete calls
  double ls_evals = 0;        // line-search derivative evaluations
  double n_cdx = 0;           // C dx / C^T dz products
  double n_global_res = 0;    // global residual evaluations (primal + dual)
  double n_newton = 0;        // Newton steps
  void add(const Counters& o)
  {
    n_factor += o.n_factor; factor_m2 += o.factor_m2; factor_m3 += o.factor_m3;
    n_solve += o.n_solve; solve_m2 += o.solve_m2; solve_m += o.solve_m;
    n_resid += o.n_resid; resid_nc += o.resid_nc;
    rank_rt2 += o.rank_rt2; rank_chunk_t2 += o.rank_chunk_t2; rank_rt += o.rank_rt;
    n_insert += o.n_insert; insert_bytes += o.insert_bytes;
    n_delete += o.n_delete; delete_t2 += o.delete_t2;
    ls_evals += o.ls_evals; n_cdx += o.n_cdx; n_global_res += o.n_global_res;
    n_newton += o.n_newton;
  }
};

// Generator of "how many w columns are live at this column", update.hpp:282-286
// (ConstantR) and modify.hpp:58-79 (IndicesR).
struct RFn
{
  bool constant = true;
  isize r_const = 0;
  isize current_col = 0;
  isize current_r = 0;
  isize r = 0;
  const isize* indices = nullptr;
  isize operator()()
  {
    if (constant) {
      return r_const;
    }
    if (current_r == r) {
      return current_r;
    }
    while (current_col == indices[current_r] - current_r) {
      ++current_r;
      if (current_r == r) {
        return current_r;
      }
    }
    ++current_col;
    return current_r;
  }
};

// update.hpp:221-287. `ld` points at element (0,0) of an n x n column-major
// block with column stride `lds`; diagonal holds D, strict lower holds L.
inline void
rank_r_update_clobber_w(double* ld, isize lds, isize n, double* pw, isize w_stride, double* palpha, RFn r_fn, Counters* cnt)
{
  for (isize j = 0; j < n; ++j) {
    isize r = r_fn();
    isize r_done = 0;
    if (!(r_done < r)) {
      continue; // update.hpp:236-238 (pw is not advanced, as in the reference)
    }
    while (true) {
      isize r_chunk = std::min<isize>(4, r - r_done);
      double p_array[4];
      double mu_array[4];
      double dj = ld[j * lds + j];
      for (isize k = 0; k < r_chunk; ++k) {
        double& alpha = palpha[r_done + k];
        double p = pw[(r_done + k) * w_stride];
        double new_dj = dj + (alpha * p) * p;
        double mu = (alpha * p) / new_dj;
        alpha -= new_dj * (mu * mu);
        dj = new_dj;
        p_array[k] = p;
        mu_array[k] = mu;
      }
      ld[j * lds + j] = dj;
      isize rem = n - j - 1;
      double* __restrict l = ld + j * lds + j + 1;
      double* w0 = pw + 1 + r_done * w_stride;
      switch (r_chunk) {
        case 1: {
          double* __restrict wa = w0;
          const double p0 = p_array[0], m0 = mu_array[0];
          for (isize i = 0; i < rem; ++i) {
            double li = l[i];
            double a = wa[i] - p0 * li;
            li += m0 * a;
            wa[i] = a;
            l[i] = li;
          }
        } break;
        case 2: {
          double* __restrict wa = w0;
          double* __restrict wb = w0 + w_stride;
          const double p0 = p_array[0], m0 = mu_array[0];
          const double p1 = p_array[1], m1 = mu_array[1];
          for (isize i = 0; i < rem; ++i) {
            double li = l[i];
            double a = wa[i] - p0 * li;
            li += m0 * a;
            double b = wb[i] - p1 * li;
            li += m1 * b;
            wa[i] = a;
            wb[i] = b;
            l[i] = li;
          }
        } break;
        case 3: {
          double* __restrict wa = w0;
          double* __restrict wb = w0 + w_stride;
          double* __restrict wc = w0 + 2 * w_stride;
          const double p0 = p_array[0], m0 = mu_array[0];
          const double p1 = p_array[1], m1 = mu_array[1];
          const double p2 = p_array[2], m2 = mu_array[2];
          for (isize i = 0; i < rem; ++i) {
            double li = l[i];
            double a = wa[i] - p0 * li;
            li += m0 * a;
            double b = wb[i] - p1 * li;
            li += m1 * b;
            double c = wc[i] - p2 * li;
            li += m2 * c;
            wa[i] = a;
            wb[i] = b;
            wc[i] = c;
            l[i] = li;
          }
        } break;
        default: {
          double* __restrict wa = w0;
          double* __restrict wb = w0 + w_stride;
          double* __restrict wc = w0 + 2 * w_stride;
          double* __restrict wd = w0 + 3 * w_stride;
          const double p0 = p_array[0], m0 = mu_array[0];
          const double p1 = p_array[1], m1 = mu_array[1];
          const double p2 = p_array[2], m2 = mu_array[2];
          const double p3 = p_array[3], m3 = mu_array[3];
          for (isize i = 0; i < rem; ++i) {
            double li = l[i];
            double a = wa[i] - p0 * li;
            li += m0 * a;
            double b = wb[i] - p1 * li;
            li += m1 * b;
            double c = wc[i] - p2 * li;
            li += m2 * c;
            double d = wd[i] - p3 * li;
            li += m3 * d;
            wa[i] = a;
            wb[i] = b;
            wc[i] = c;
            wd[i] = d;
            l[i] = li;
          }
        } break;
      }
      if (cnt) {
        cnt->rank_rt2 += double(r_chunk) * double(rem);
        cnt->rank_chunk_t2 += double(rem);
        cnt->rank_rt += double(r_chunk);
      }
      r_done += r_chunk;
      if (!(r_done < r)) {
        break;
      }
    }
    ++pw;
  }
}

struct Ldlt
{
  // column-major L\D storage with a fixed column stride (capacity).
  std::vector<double> ld;
  isize stride = 0;
  isize n = 0; // current dimension
  std::vector<isize> perm;
  std::vector<isize> perm_inv;
  std::vector<double> maybe_sorted_diag;
  // scratch
  std::vector<double> work;
  std::vector<double> wbuf;
  std::vector<double> abuf;
  std::vector<isize> ibuf;
  Counters* cnt = nullptr;

  void reserve(isize cap)
  {
    if (cap <= stride) {
      return;
    }
    std::vector<double> nld(std::size_t(cap) * std::size_t(cap), 0.0);
    for (isize j = 0; j < n; ++j) {
      std::memcpy(&nld[std::size_t(j) * std::size_t(cap)], &ld[std::size_t(j) * std::size_t(stride)], sizeof(double) * std::size_t(n));
    }
    ld.swap(nld);
    stride = cap;
    perm.reserve(std::size_t(cap));
    perm_inv.reserve(std::size_t(cap));
    maybe_sorted_diag.reserve(std::size_t(cap));
    work.resize(std::size_t(cap));
  }
  isize dim() const { return n; }
  double& at(isize i, isize j) { return ld[std::size_t(j) * std::size_t(stride) + std::size_t(i)]; }
  double at(isize i, isize j) const { return ld[std::size_t(j) * std::size_t(stride) + std::size_t(i)]; }

  // in-place LDLT of the leading nn x nn block starting at (off, off).
  // Right-looking elimination, column by column (same eliminations as
  // factorize.hpp:91-148, different association order).
  void factorize_block(isize off, isize nn)
  {
    double* base = &ld[std::size_t(off) * std::size_t(stride) + std::size_t(off)];
    for (isize j = 0; j < nn; ++j) {
      double* cj = base + j * stride;
      double d = cj[j];
      double inv = 1.0 / d;
      isize rem = nn - j - 1;
      for (isize i = j + 1; i < nn; ++i) {
        work[std::size_t(i)] = cj[i]; // l_ij * d
        cj[i] *= inv;
      }
      for (isize k = j + 1; k < nn; ++k) {
        double* __restrict ck = base + k * stride;
        const double f = cj[k]; // l_kj
        const double* __restrict wv = work.data();
        for (isize i = k; i < nn; ++i) {
          ck[i] -= wv[i] * f;
        }
      }
      (void)rem;
    }
  }

  // ldlt.hpp:718-744. `mat` is m x m, `lda` its row length; only the lower
  // triangle (i >= j) of the symmetric matrix is read as mat[i*lda + j].
  void factorize(const double* mat, isize lda, isize m)
  {
    reserve(m);
    n = m;
    perm.resize(std::size_t(m));
    perm_inv.resize(std::size_t(m));
    maybe_sorted_diag.resize(std::size_t(m));
    for (isize k = 0; k < m; ++k) {
      perm[std::size_t(k)] = k;
    }
    std::sort(perm.begin(), perm.end(), [mat, lda](isize i, isize j) {
      double lhs = std::fabs(mat[i * lda + i]);
      double rhs = std::fabs(mat[j * lda + j]);
      if (lhs == rhs) {
        return i < j;
      }
      return lhs > rhs;
    });
    for (isize k = 0; k < m; ++k) {
      perm_inv[std::size_t(perm[std::size_t(k)])] = k;
    }
    for (isize j = 0; j < m; ++j) {
      isize pj = perm[std::size_t(j)];
      for (isize i = j; i < m; ++i) {
        isize pi = perm[std::size_t(i)];
        at(i, j) = pi >= pj ? mat[pi * lda + pj] : mat[pj * lda + pi];
      }
    }
    for (isize i = 0; i < m; ++i) {
      maybe_sorted_diag[std::size_t(i)] = at(i, i);
    }
    factorize_block(0, m);
    if (cnt) {
      cnt->n_factor += 1;
      cnt->factor_m2 += double(m) * double(m);
      cnt->factor_m3 += double(m) * double(m) * double(m);
    }
  }

  // solve.hpp:17-26 on the trailing-free full factor.
  void solve_core(double* x, isize m) const
  {
    // forward: unit lower, column oriented
    for (isize j = 0; j < m; ++j) {
      const double xj = x[j];
      const double* __restrict cj = &ld[std::size_t(j) * std::size_t(stride)];
      for (isize i = j + 1; i < m; ++i) {
        x[i] -= cj[i] * xj;
      }
    }
    for (isize j = 0; j < m; ++j) {
      x[j] /= at(j, j);
    }
    // backward: L^T, row of L^T = column of L
    for (isize j = m - 1; j >= 0; --j) {
      const double* __restrict cj = &ld[std::size_t(j) * std::size_t(stride)];
      double acc = x[j];
      for (isize i = j + 1; i < m; ++i) {
        acc -= cj[i] * x[i];
      }
      x[j] = acc;
    }
  }

  // ldlt.hpp:767-782
  void solve_in_place(double* rhs, isize m)
  {
    work.resize(std::size_t(std::max<isize>(stride, m)));
    for (isize i = 0; i < m; ++i) {
      work[std::size_t(i)] = rhs[perm[std::size_t(i)]];
    }
    solve_core(work.data(), m);
    for (isize i = 0; i < m; ++i) {
      rhs[i] = work[std::size_t(perm_inv[std::size_t(i)])];
    }
    if (cnt) {
      cnt->n_solve += 1;
      cnt->solve_m2 += double(m) * double(m);
      cnt->solve_m += double(m);
    }
  }

  // modify.hpp:20-46
  void delete_rows_and_cols_triangular(const isize* indices, isize r)
  {
    isize nn = n;
    for (isize chunk_j = 0; chunk_j < r + 1; ++chunk_j) {
      isize j_start = chunk_j == 0 ? 0 : indices[chunk_j - 1] + 1;
      isize j_finish = chunk_j == r ? nn : indices[chunk_j];
      for (isize j = j_start; j < j_finish; ++j) {
        for (isize chunk_i = chunk_j; chunk_i < r + 1; ++chunk_i) {
          isize i_start = chunk_i == chunk_j ? j : indices[chunk_i - 1] + 1;
          isize i_finish = chunk_i == r ? nn : indices[chunk_i];
          if (chunk_i != 0 || chunk_j != 0) {
            if (i_finish > i_start) {
              std::memmove(&at(i_start - chunk_i, j - chunk_j), &at(i_start, j), sizeof(double) * std::size_t(i_finish - i_start));
            }
          }
        }
      }
    }
  }

  // ldlt.hpp:340-387 + modify.hpp:82-127. `indices` sorted ascending
  // (unpermuted indices).
  void delete_at(const isize* indices, isize r)
  {
    if (r == 0) {
      return;
    }
    isize nn = n;
    std::vector<isize>& ia = ibuf;
    ia.resize(std::size_t(2 * r));
    isize* indices_actual = ia.data();
    for (isize k = 0; k < r; ++k) {
      indices_actual[k] = perm_inv[std::size_t(indices[k])];
    }
    // ldlt_delete_rows_and_cols_impl on a sorted copy
    {
      isize* idx = ia.data() + r;
      for (isize k = 0; k < r; ++k) {
        idx[k] = indices_actual[k];
      }
      std::sort(idx, idx + r);
      isize first = idx[0];
      isize w_stride = nn - first - r;
      if (w_stride < 1) {
        w_stride = 1;
      }
      wbuf.assign(std::size_t(r * w_stride), 0.0);
      abuf.resize(std::size_t(r));
      double* pw = wbuf.data();
      double* palpha = abuf.data();
      for (isize k = 0; k < r; ++k) {
        isize j = idx[k];
        palpha[k] = at(j, j);
        double* pwk = pw + k * w_stride;
        for (isize chunk_i = k + 1; chunk_i < r + 1; ++chunk_i) {
          isize i_start = idx[chunk_i - 1] + 1;
          isize i_finish = chunk_i == r ? nn : idx[chunk_i];
          if (i_finish > i_start) {
            std::memcpy(pwk + i_start - chunk_i - first, &at(i_start, j), sizeof(double) * std::size_t(i_finish - i_start));
          }
        }
      }
      delete_rows_and_cols_triangular(idx, r);
      RFn fn;
      fn.constant = false;
      fn.current_col = first;
      fn.current_r = 0;
      fn.r = r;
      fn.indices = idx;
      isize t = nn - first - r;
      if (cnt) {
        cnt->n_delete += double(r);
        cnt->delete_t2 += double(nn - first) * double(nn - first);
      }
      rank_r_update_clobber_w(&at(first, first), stride, t, pw, w_stride, palpha, fn, cnt);
    }
    // permutation bookkeeping, ldlt.hpp:366-386
    for (isize k = 0; k < r; ++k) {
      // the reference sorts indices_actual in place (modify.hpp:89) before
      // this loop, so sorted positions are paired with sorted indices.
      isize i_actual = ia[std::size_t(r + (r - 1 - k))];
      isize i = indices[r - 1 - k];
      perm.erase(perm.begin() + i_actual);
      perm_inv.erase(perm_inv.begin() + i);
      maybe_sorted_diag.erase(maybe_sorted_diag.begin() + i_actual);
      for (isize j = 0; j < nn - 1 - k; ++j) {
        isize& p_j = perm[std::size_t(j)];
        isize& pinv_j = perm_inv[std::size_t(j)];
        if (p_j > i) {
          --p_j;
        }
        if (pinv_j > i_actual) {
          --pinv_j;
        }
      }
    }
    n = nn - r;
  }

  // ldlt.hpp:431-475 + modify.hpp:131-264. `a` is column-major
  // (n + r) x r with column stride `lda` (unpermuted row indices).
  void insert_block_at(isize i, const double* a, isize lda, isize r)
  {
    if (r == 0) {
      return;
    }
    isize old_n = n;
    reserve(old_n + r);
    // choose_insertion_position(i, a.col(0))
    isize i_actual = 0;
    {
      double diag_elem = a[i];
      for (; i_actual < old_n; ++i_actual) {
        if (diag_elem >= maybe_sorted_diag[std::size_t(i_actual)]) {
          break;
        }
      }
    }
    for (isize j = 0; j < old_n; ++j) {
      isize& p_j = perm[std::size_t(j)];
      isize& pinv_j = perm_inv[std::size_t(j)];
      if (p_j >= i) {
        p_j += r;
      }
      if (pinv_j >= i_actual) {
        pinv_j += r;
      }
    }
    for (isize k = 0; k < r; ++k) {
      perm.insert(perm.begin() + (i_actual + k), i + k);
      perm_inv.insert(perm_inv.begin() + (i + k), i_actual + k);
      maybe_sorted_diag.insert(maybe_sorted_diag.begin() + (i_actual + k), a[k * lda + (i + k)]);
    }
    isize new_n = old_n + r;
    n = new_n;
    // permuted_a
    std::vector<double>& pa = abuf;
    pa.resize(std::size_t(new_n * r));
    for (isize k = 0; k < r; ++k) {
      for (isize j = 0; j < new_n; ++j) {
        pa[std::size_t(k * new_n + j)] = a[k * lda + perm[std::size_t(j)]];
      }
    }
    isize pos = i_actual;
    // shift trailing columns right by r and rows down by r (modify.hpp:145-179)
    {
      isize current_col = old_n;
      while (true) {
        if (current_col == pos) {
          break;
        }
        --current_col;
        double* src = &at(0, current_col);
        double* dst = &at(0, current_col + r);
        std::memmove(dst + pos + r, src + pos, sizeof(double) * std::size_t(old_n - pos));
        std::memmove(dst, src, sizeof(double) * std::size_t(pos));
      }
      while (true) {
        if (current_col == 0) {
          break;
        }
        --current_col;
        double* src = &at(0, current_col);
        std::memmove(src + pos + r, src + pos, sizeof(double) * std::size_t(old_n - pos));
      }
    }
    isize rem = new_n - pos - r;
    if (cnt) {
      cnt->n_insert += double(r);
      cnt->insert_bytes += 8.0 * (0.5 * double(pos) * double(pos) * double(r) + 2.0 * double(rem) * double(rem + pos));
    }
    // l10 = (L00^{-1} a01)^T D0^{-1} ; row block (r x pos) stored inside ld at rows pos..pos+r
    // First put a01^T into l10 then solve X L00^T = a01^T  <=> L00 X^T = a01.
    for (isize k = 0; k < r; ++k) {
      const double* a01 = &pa[std::size_t(k * new_n)];
      // forward substitution with unit lower L00 on a copy, column oriented
      double* x = work.data();
      for (isize j = 0; j < pos; ++j) {
        x[j] = a01[j];
      }
      for (isize j = 0; j < pos; ++j) {
        const double xj = x[j];
        const double* __restrict cj = &ld[std::size_t(j) * std::size_t(stride)];
        for (isize ii = j + 1; ii < pos; ++ii) {
          x[ii] -= cj[ii] * xj;
        }
      }
      for (isize j = 0; j < pos; ++j) {
        at(pos + k, j) = x[j] / at(j, j);
      }
    }
    // d0 x l10^T (pos x r), ld11, l21
    std::vector<double>& d0l = wbuf;
    d0l.resize(std::size_t(std::max<isize>(1, pos) * r + r * std::max<isize>(1, rem)));
    for (isize k = 0; k < r; ++k) {
      for (isize j = 0; j < pos; ++j) {
        d0l[std::size_t(k * pos + j)] = at(j, j) * at(pos + k, j);
      }
    }
    // ld11 lower = a11 lower - l10 * d0l
    for (isize c = 0; c < r; ++c) {
      for (isize rr = c; rr < r; ++rr) {
        double v = pa[std::size_t(c * new_n + pos + rr)];
        double acc = 0;
        for (isize j = 0; j < pos; ++j) {
          acc += at(pos + rr, j) * d0l[std::size_t(c * pos + j)];
        }
        at(pos + rr, pos + c) = v - acc;
      }
    }
    // l21 = a21 - l20 * d0l   (rem x r)
    for (isize c = 0; c < r; ++c) {
      double* __restrict l21c = &at(pos + r, pos + c);
      const double* a21 = &pa[std::size_t(c * new_n + pos + r)];
      for (isize ii = 0; ii < rem; ++ii) {
        l21c[ii] = a21[ii];
      }
      for (isize j = 0; j < pos; ++j) {
        const double f = d0l[std::size_t(c * pos + j)];
        const double* __restrict l20j = &at(pos + r, j);
        for (isize ii = 0; ii < rem; ++ii) {
          l21c[ii] -= l20j[ii] * f;
        }
      }
    }
    // factorize ld11 in place (r x r)
    factorize_block(pos, r);
    // l21 <- l21 * L11^{-T} * D1^{-1}
    for (isize c = 0; c < r; ++c) {
      double* __restrict l21c = &at(pos + r, pos + c);
      for (isize k = 0; k < c; ++k) {
        const double f = at(pos + c, pos + k);
        const double* __restrict l21k = &at(pos + r, pos + k);
        // l21k currently holds (column k of X) * d_k? handle scaling after loop
        for (isize ii = 0; ii < rem; ++ii) {
          l21c[ii] -= l21k[ii] * f;
        }
      }
    }
    // NOTE: the triangular solve above must use unscaled columns; scaling by
    // D1^{-1} is applied afterwards for all columns.
    for (isize c = 0; c < r; ++c) {
      double inv = 1.0 / at(pos + c, pos + c);
      double* __restrict l21c = &at(pos + r, pos + c);
      for (isize ii = 0; ii < rem; ++ii) {
        l21c[ii] *= inv;
      }
    }
    // trailing rank-r update with w = l21, alpha = -d1
    {
      isize w_stride = std::max<isize>(1, rem);
      double* pw = d0l.data() + std::max<isize>(1, pos) * r;
      std::vector<double> alpha_v(static_cast<std::size_t>(r));
      for (isize k = 0; k < r; ++k) {
        alpha_v[std::size_t(k)] = -at(pos + k, pos + k);
        const double* src = &at(pos + r, pos + k);
        for (isize ii = 0; ii < rem; ++ii) {
          pw[k * w_stride + ii] = src[ii];
        }
      }
      RFn fn;
      fn.constant = true;
      fn.r_const = r;
      rank_r_update_clobber_w(&at(pos + r, pos + r), stride, rem, pw, w_stride, alpha_v.data(), fn, cnt);
    }
  }

  // ldlt.hpp:516-570. indices are clobbered.
  void diagonal_update_clobber_indices(isize* indices, isize r, const double* alpha)
  {
    if (r == 0) {
      return;
    }
    std::vector<isize> positions(static_cast<std::size_t>(r));
    std::vector<isize> sorted_indices(static_cast<std::size_t>(r));
    for (isize k = 0; k < r; ++k) {
      indices[k] = perm_inv[std::size_t(indices[k])];
      positions[std::size_t(k)] = k;
    }
    std::sort(positions.begin(), positions.end(), [indices](isize i, isize j) { return indices[i] < indices[j]; });
    for (isize k = 0; k < r; ++k) {
      sorted_indices[std::size_t(k)] = indices[positions[std::size_t(k)]];
    }
    isize first = sorted_indices[0];
    isize nn = n - first;
    wbuf.assign(std::size_t(nn * r), 0.0);
    abuf.resize(std::size_t(r));
    for (isize k = 0; k < r; ++k) {
      abuf[std::size_t(k)] = alpha[positions[std::size_t(k)]];
      wbuf[std::size_t(k * nn + sorted_indices[std::size_t(k)] - first)] = 1.0;
    }
    RFn fn;
    fn.constant = false;
    fn.current_col = first;
    fn.current_r = 0;
    fn.r = r;
    fn.indices = sorted_indices.data();
    rank_r_update_clobber_w(&at(first, first), stride, nn, wbuf.data(), nn, abuf.data(), fn, cnt);
  }

  // ldlt.hpp:580-609. w is column-major n x r with column stride ldw
  // (unpermuted rows).
  void rank_r_update(const double* w, isize ldw, isize r, const double* alpha)
  {
    if (r == 0) {
      return;
    }
    isize nn = n;
    wbuf.resize(std::size_t(nn * r));
    abuf.resize(std::size_t(r));
    for (isize k = 0; k < r; ++k) {
      double alpha_tmp = alpha[k];
      abuf[std::size_t(k)] = alpha_tmp;
      for (isize i = 0; i < nn; ++i) {
        double w_tmp = w[k * ldw + perm[std::size_t(i)]];
        wbuf[std::size_t(k * nn + i)] = w_tmp;
        maybe_sorted_diag[std::size_t(i)] += alpha_tmp * (w_tmp * w_tmp);
      }
    }
    RFn fn;
    fn.constant = true;
    fn.r_const = r;
    rank_r_update_clobber_w(&at(0, 0), stride, nn, wbuf.data(), nn, abuf.data(), fn, cnt);
  }

  // debugging helper: reconstruct the (unpermuted) matrix into out (row-major n x n)
  void reconstruct(std::vector<double>& out) const
  {
    isize m = n;
    std::vector<double> A(std::size_t(m * m), 0.0);
    for (isize i = 0; i < m; ++i) {
      for (isize j = 0; j <= i; ++j) {
        double acc = 0;
        for (isize k = 0; k <= j; ++k) {
          double lik = (i == k) ? 1.0 : at(i, k);
          double ljk = (j == k) ? 1.0 : at(j, k);
          acc += lik * at(k, k) * ljk;
        }
        A[std::size_t(i * m + j)] = acc;
        A[std::size_t(j * m + i)] = acc;
      }
    }
    out.assign(std::size_t(m * m), 0.0);
    for (isize i = 0; i < m; ++i) {
      for (isize j = 0; j < m; ++j) {
        out[std::size_t(perm[std::size_t(i)] * m + perm[std::size_t(j)])] = A[std::size_t(i * m + j)];
      }
    }
  }
};

} // namespace oracle
