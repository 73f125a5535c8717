// highs_b200/csrc/host_prep.cpp -- see host_prep.hpp.
//
// The arithmetic (which quotient is taken when, in which order sums run) follows
// the reference so that the scaled data are bit-identical to what cuPDLP-C
// iterates on; the loop structure does not: the reference makes four sweeps over
// the nonzeros per Ruiz pass (column norms, row norms, row division, column
// division), here every pass is one sweep that applies pass k and gathers the
// norms of pass k+1 at the same time.
#include "host_prep.hpp"

#include <algorithm>
#include <cmath>
#include <limits>
#include <stdexcept>

namespace b200 {

// formulateLP_highs, CupdlpWrapper.cpp:280-448
void formulate(const b200pdlp_lp& lp, StdForm& f) {
  const int n0 = lp.num_col, m = lp.num_row;
  const int nnz0 = lp.a_start[n0];
  f = StdForm();
  f.n_orig = n0;
  f.m = m;
  f.sense = lp.sense;
  f.offset = lp.offset;
  f.row_class.resize(m);
  f.row_new_idx.resize(m);
  int n_eq_like = 0, n_bound = 0;
  for (int i = 0; i < m; i++) {
    const bool lo = lp.row_lower[i] > -1e20, up = lp.row_upper[i] < 1e20;  // :316-317
    RowClass c;
    if (lo && up && lp.row_lower[i] == lp.row_upper[i]) c = kEq;
    else if (lo && !up) c = kGeq;
    else if (!lo && up) c = kLeq;
    else c = kBound;  // ranged rows and (with a warning in the reference) free rows, :328-345
    f.row_class[i] = c;
    if (c == kEq || c == kBound) n_eq_like++;
    if (c == kBound) n_bound++;
  }
  f.neq = n_eq_like;
  f.n = n0 + n_bound;
  f.nnz = nnz0 + n_bound;
  // new row order: EQ/BOUND rows (original order), then LEQ/GEQ rows (:380-404)
  f.rhs.assign(m, 0.0);
  {
    int head = 0, tail = f.neq;
    for (int i = 0; i < m; i++) {
      switch (f.row_class[i]) {
        case kEq: f.rhs[head] = lp.row_lower[i]; f.row_new_idx[i] = head++; break;
        case kBound: f.rhs[head] = 0.0; f.row_new_idx[i] = head++; break;
        case kLeq: f.rhs[tail] = -lp.row_upper[i]; f.row_new_idx[i] = tail++; break;
        case kGeq: f.rhs[tail] = lp.row_lower[i]; f.row_new_idx[i] = tail++; break;
      }
    }
  }
  f.cost.assign(f.n, 0.0);
  f.lower.resize(f.n);
  f.upper.resize(f.n);
  for (int j = 0; j < n0; j++) {
    f.cost[j] = lp.col_cost[j] * lp.sense;
    f.lower[j] = lp.col_lower[j];
    f.upper[j] = lp.col_upper[j];
  }
  f.cbeg.resize(f.n + 1);
  f.cidx.resize(f.nnz);
  f.cval.resize(f.nnz);
  // structural columns: within a column EQ/BOUND entries first, then LEQ (negated) / GEQ (:410-433)
  int k = 0;
  for (int j = 0; j < n0; j++) {
    f.cbeg[j] = k;
    for (int p = lp.a_start[j]; p < lp.a_start[j + 1]; p++) {
      const int c = f.row_class[lp.a_index[p]];
      if (c == kEq || c == kBound) { f.cidx[k] = f.row_new_idx[lp.a_index[p]]; f.cval[k] = lp.a_value[p]; k++; }
    }
    for (int p = lp.a_start[j]; p < lp.a_start[j + 1]; p++) {
      const int c = f.row_class[lp.a_index[p]];
      if (c == kLeq) { f.cidx[k] = f.row_new_idx[lp.a_index[p]]; f.cval[k] = -lp.a_value[p]; k++; }
      else if (c == kGeq) { f.cidx[k] = f.row_new_idx[lp.a_index[p]]; f.cval[k] = lp.a_value[p]; k++; }
    }
  }
  // one slack column per BOUND row: A x - z = 0, lo <= z <= up (:328-331,367-373,439-445)
  int j = n0;
  for (int i = 0; i < m; i++) {
    if (f.row_class[i] != kBound) continue;
    f.cbeg[j] = k;
    f.cidx[k] = f.row_new_idx[i];
    f.cval[k] = -1.0;
    k++;
    f.lower[j] = lp.row_lower[i];
    f.upper[j] = lp.row_upper[i];
    j++;
  }
  f.cbeg[f.n] = k;
  const double inf = std::numeric_limits<double>::infinity();
  for (int q = 0; q < f.n; q++) {  // :375-378
    if (f.lower[q] < -1e20) f.lower[q] = -inf;
    if (f.upper[q] > 1e20) f.upper[q] = inf;
  }
  // Init_Scaling (cupdlp_scaling.c:395-425): norms of the unscaled data, unit scales
  f.col_scale.assign(f.n, 1.0);
  f.row_scale.assign(f.m, 1.0);
  double s = 0.0;
  for (int q = 0; q < f.n; q++) s += f.cost[q] * f.cost[q];
  f.norm_cost = std::sqrt(s);
  s = 0.0;
  for (int i = 0; i < f.m; i++) s += f.rhs[i] * f.rhs[i];
  f.norm_rhs = std::sqrt(s);
}

namespace {
// divide/multiply the vectors by one pass's factors and fold them into the running scales
// (scale_problem, cupdlp_scaling.c:17-31, and the cdot updates at :110-111)
void apply_to_vectors(StdForm& f, const std::vector<double>& cs, const std::vector<double>& rs) {
  for (int j = 0; j < f.n; j++) {
    f.cost[j] /= cs[j];
    f.lower[j] *= cs[j];
    f.upper[j] *= cs[j];
    f.col_scale[j] *= cs[j];
  }
  for (int i = 0; i < f.m; i++) {
    f.rhs[i] /= rs[i];
    f.row_scale[i] *= rs[i];
  }
}
}  // namespace

// PDHG_Scale_Data with Init_Scaling's fixed recipe: 10 inf-norm Ruiz passes, then
// Pock-Chambolle with alpha = 1 (cupdlp_scaling.c:47-120, 174-231, 395-409)
void scale(StdForm& f, bool do_scale) {
  const int n = f.n, m = f.m;
  double amax = 0.0;
  if (!do_scale) {
    for (int p = 0; p < f.nnz; p++) amax = std::max(amax, std::fabs(f.cval[p]));
    f.amax = amax;
    return;
  }
  std::vector<double> cs(n), rs(m), cs_next(n), rs_next(m);
  // norms for the first Ruiz pass
  std::fill(rs.begin(), rs.end(), 0.0);
  for (int j = 0; j < n; j++) {
    double mx = 0.0;
    for (int p = f.cbeg[j]; p < f.cbeg[j + 1]; p++) {
      const double a = std::fabs(f.cval[p]);
      if (a > mx) mx = a;
      if (rs[f.cidx[p]] < a) rs[f.cidx[p]] = a;
    }
    cs[j] = mx;
  }
  const int kRuiz = 10;
  for (int it = 0; it <= kRuiz; it++) {
    // turn the gathered norms into this pass's factors (sqrt, with 0 -> 1)
    for (int j = 0; j < n; j++) { const double v = std::sqrt(cs[j]); cs[j] = (v == 0.0) ? 1.0 : v; }
    for (int i = 0; i < m; i++) rs[i] = (rs[i] == 0.0) ? 1.0 : std::sqrt(rs[i]);
    apply_to_vectors(f, cs, rs);
    const bool next_is_pc = (it == kRuiz - 1);  // after the 10th Ruiz pass gather 1-norms
    const bool last = (it == kRuiz);            // the Pock-Chambolle pass itself
    std::fill(rs_next.begin(), rs_next.end(), 0.0);
    for (int j = 0; j < n; j++) {
      double acc = 0.0;
      const double cj = cs[j];
      for (int p = f.cbeg[j]; p < f.cbeg[j + 1]; p++) {
        const int i = f.cidx[p];
        double v = f.cval[p] / rs[i];   // row division first, then column division (:33-41)
        v /= cj;
        f.cval[p] = v;
        const double a = std::fabs(v);
        if (last) { if (a > amax) amax = a; }
        else if (next_is_pc) { acc += a; rs_next[i] += a; }
        else { if (a > acc) acc = a; if (rs_next[i] < a) rs_next[i] = a; }
      }
      cs_next[j] = acc;
    }
    cs.swap(cs_next);
    rs.swap(rs_next);
  }
  f.amax = amax;
}

std::vector<int> partition_rows(const StdForm& f, int world) {
  std::vector<long long> cnt(f.m + 1, 0);
  for (int p = 0; p < f.nnz; p++) cnt[f.cidx[p] + 1]++;
  for (int i = 0; i < f.m; i++) cnt[i + 1] += cnt[i];
  std::vector<int> b(world + 1, 0);
  b[world] = f.m;
  // balance nnz + a per-row cost (vector traffic ~ 7 doubles per row ~ 5 nonzeros)
  const long long total = cnt[f.m] + 5LL * f.m;
  for (int g = 1; g < world; g++) {
    const long long target = total * g / world;
    int lo = b[g - 1], hi = f.m;
    while (lo < hi) {
      const int mid = (lo + hi) / 2;
      if (cnt[mid] + 5LL * mid < target) lo = mid + 1; else hi = mid;
    }
    b[g] = lo;
  }
  return b;
}

void build_row_major(const StdForm& f, int r0, int r1, Csr& a) {
  a = Csr();
  a.nrows = r1 - r0;
  a.ncols = f.n;
  a.rowptr.assign(a.nrows + 1, 0);
  for (int p = 0; p < f.nnz; p++) { const int i = f.cidx[p]; if (i >= r0 && i < r1) a.rowptr[i - r0 + 1]++; }
  for (int i = 0; i < a.nrows; i++) a.rowptr[i + 1] += a.rowptr[i];
  a.nnz = a.rowptr[a.nrows];
  a.col.assign(a.nnz, 0);
  a.val.assign(a.nnz, 0.0);
  std::vector<int> w(a.rowptr.begin(), a.rowptr.end() - 1);
  for (int j = 0; j < f.n; j++)
    for (int p = f.cbeg[j]; p < f.cbeg[j + 1]; p++) {
      const int i = f.cidx[p];
      if (i < r0 || i >= r1) continue;
      const int q = w[i - r0]++;
      a.col[q] = j;
      a.val[q] = f.cval[p];
    }
}

void build_col_major(const StdForm& f, int r0, int r1, Csr& at) {
  // rows of A^T = columns of A restricted to rows [r0,r1), entries sorted by row
  // (the order in which the reference's row-scatter A^T y accumulates, cupdlp_linalg.c:73-109)
  at = Csr();
  at.nrows = f.n;
  at.ncols = r1 - r0;
  at.rowptr.assign(f.n + 1, 0);
  for (int j = 0; j < f.n; j++) {
    int c = 0;
    for (int p = f.cbeg[j]; p < f.cbeg[j + 1]; p++) c += (f.cidx[p] >= r0 && f.cidx[p] < r1);
    at.rowptr[j + 1] = at.rowptr[j] + c;
  }
  at.nnz = at.rowptr[f.n];
  at.col.assign(at.nnz, 0);
  at.val.assign(at.nnz, 0.0);
  std::vector<std::pair<int, double>> tmp;
  for (int j = 0; j < f.n; j++) {
    int q = at.rowptr[j];
    bool sorted = true;
    int prev = -1;
    for (int p = f.cbeg[j]; p < f.cbeg[j + 1]; p++) {
      const int i = f.cidx[p];
      if (i < r0 || i >= r1) continue;
      if (i < prev) sorted = false;
      prev = i;
      at.col[q] = i - r0;
      at.val[q] = f.cval[p];
      q++;
    }
    if (!sorted) {
      tmp.clear();
      for (int t = at.rowptr[j]; t < q; t++) tmp.emplace_back(at.col[t], at.val[t]);
      std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<int, double>& a, const std::pair<int, double>& b) { return a.first < b.first; });
      for (size_t t = 0; t < tmp.size(); t++) { at.col[at.rowptr[j] + t] = tmp[t].first; at.val[at.rowptr[j] + t] = tmp[t].second; }
    }
  }
}

std::vector<int> make_perm(const std::vector<int>& rowptr, int boundary, bool sort) {
  const int n = (int)rowptr.size() - 1;
  std::vector<int> perm(n);
  for (int i = 0; i < n; i++) perm[i] = i;
  if (!sort) return perm;
  auto len = [&](int r) { return rowptr[r + 1] - rowptr[r]; };
  auto sort_range = [&](int b, int e) {
    for (int w = b; w < e; w += kSortWindow) {
      const int we = std::min(w + kSortWindow, e);
      std::stable_sort(perm.begin() + w, perm.begin() + we, [&](int x, int y) { return len(x) > len(y); });
    }
  };
  boundary = std::max(0, std::min(boundary, n));
  sort_range(0, boundary);
  sort_range(boundary, n);
  return perm;
}

std::vector<int> invert_perm(const std::vector<int>& perm) {
  std::vector<int> inv(perm.size());
  for (size_t i = 0; i < perm.size(); i++) inv[perm[i]] = (int)i;
  return inv;
}

void build_sell(const Csr& a, const std::vector<int>& perm, const std::vector<int>& colmap, int long_threshold,
                SellMatrix& out) {
  out = SellMatrix();
  out.nrows = a.nrows;
  out.ncols = a.ncols;
  out.nnz = a.nnz;
  const int nslices = (a.nrows + 31) / 32;
  out.slices.resize(nslices);
  auto len = [&](int newrow) { const int r = perm[newrow]; return a.rowptr[r + 1] - a.rowptr[r]; };
  long long total = 0;
  for (int s = 0; s < nslices; s++) {
    int mx = 0;
    unsigned mask = 0;
    for (int l = 0; l < 32; l++) {
      const int nr = s * 32 + l;
      if (nr >= a.nrows) { mask |= 1u << l; continue; }
      const int ln = len(nr);
      if (ln > long_threshold) { mask |= 1u << l; continue; }
      mx = std::max(mx, ln);
    }
    if (total + 32LL * mx > 2000000000LL) throw std::runtime_error("b200pdlp: matrix too large for 32-bit slice offsets");
    out.slices[s] = {(int)total, mx, mask, 0};
    total += 32LL * mx;
  }
  out.padded = total;
  out.col.assign((size_t)total + 32, 0);
  out.val.assign((size_t)total + 32, 0.0);
  for (int s = 0; s < nslices; s++) {
    const SellMatrix::Slice& sl = out.slices[s];
    for (int l = 0; l < 32; l++) {
      if ((sl.skipmask >> l) & 1u) continue;
      const int r = perm[s * 32 + l];
      int k = 0;
      for (int p = a.rowptr[r]; p < a.rowptr[r + 1]; p++, k++) {
        out.col[(size_t)sl.ptr + 32 * (size_t)k + l] = colmap[a.col[p]];
        out.val[(size_t)sl.ptr + 32 * (size_t)k + l] = a.val[p];
      }
    }
  }
  // long rows -> segments
  for (int nr = 0; nr < a.nrows; nr++) {
    const int ln = len(nr);
    if (ln <= long_threshold) continue;
    const int r = perm[nr];
    const int nseg = (ln + kNnzPerBlock - 1) / kNnzPerBlock;
    const int base = (int)out.lcol.size();
    out.long_rows.push_back({nr, (int)out.segs.size(), nseg, out.n_partials});
    for (int p = a.rowptr[r]; p < a.rowptr[r + 1]; p++) { out.lcol.push_back(colmap[a.col[p]]); out.lval.push_back(a.val[p]); }
    for (int sg = 0; sg < nseg; sg++)
      out.segs.push_back({nr, base + sg * kNnzPerBlock, std::min(base + (sg + 1) * kNnzPerBlock, base + ln), (int)out.long_rows.size() - 1});
    out.n_partials += nseg;
    while (out.lcol.size() % 4) { out.lcol.push_back(0); out.lval.push_back(0.0); }   // segments start 16-byte aligned
  }
  out.lcol.resize(out.lcol.size() + 8, 0);
  out.lval.resize(out.lval.size() + 8, 0.0);
}

}  // namespace b200
