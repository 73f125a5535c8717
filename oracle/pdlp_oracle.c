/* oracle/pdlp_oracle.c -- TEST INFRASTRUCTURE ONLY (see pdlp_oracle.h).
 *
 * Plain-C restatement of the reference's CPU pdlp (cuPDLP-C as vendored by HiGHS
 * 1.15.1).  Written from the algorithm, not transcribed: one flat state struct,
 * explicit loops, no BLAS shims -- but every floating-point expression keeps the
 * reference's evaluation ORDER (copy, then axpy, then projection ...) so that
 * results are bit-identical to oracle/_ref on x86-64 without FMA contraction.
 * Compile with -O2 -ffp-contract=off.
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference/highs/pdlp/).
 */
#include "pdlp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define NEW(T, k) ((T*)calloc((size_t)((k) > 0 ? (k) : 1), sizeof(T)))

/* ------------------------------------------------------------------ helpers */
/* cupdlp_linalg.c:320-336 (USE_MY_BLAS dot), :111-126 (nrm2) */
/* TEST SWITCH (not the reference's behaviour): orc_set_sum_block(k > 0) makes the long sums add in blocks of k elements and
 * then add the block sums -- a different, equally valid rounding order, like the GPU engine's tree reductions.  Used to
 * measure how far a mere change of summation order moves a trajectory (tests/test_gpu_instances.py's tree-mode criterion). */
static int g_sum_block = 0;
void orc_set_sum_block(int k) { g_sum_block = k; }
static double vdot(int n, const double* a, const double* b) {
  double s = 0.0;
  if (g_sum_block > 0) {
    for (int i0 = 0; i0 < n; i0 += g_sum_block) {
      double t = 0.0;
      const int i1 = i0 + g_sum_block < n ? i0 + g_sum_block : n;
      for (int i = i0; i < i1; i++) t += a[i] * b[i];
      s += t;
    }
    return s;
  }
  for (int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}
static double vnrm2(int n, const double* a) {
  double s = 0.0;
  if (g_sum_block > 0) return sqrt(vdot(n, a, a));
  for (int i = 0; i < n; i++) s += a[i] * a[i];
  return sqrt(s);
}
/* y += w*x  (AddToVector, cupdlp_linalg.c:346-357) */
static void vaxpy(int n, double w, const double* x, double* y) {
  for (int i = 0; i < n; i++) y[i] += w * x[i];
}
static void vscale(int n, double w, double* x) {
  for (int i = 0; i < n; i++) x[i] *= w;
}
static void vcopy(int n, double* dst, const double* src) { memcpy(dst, src, sizeof(double) * (size_t)n); }
static void vmul(int n, double* x, const double* y) { for (int i = 0; i < n; i++) x[i] *= y[i]; }
static void vdiv(int n, double* x, const double* y) { for (int i = 0; i < n; i++) x[i] /= y[i]; }
/* cupdlp_linalg.c:211-241 */
static void clamp_hi_vec(int n, double* x, const double* ub) { for (int i = 0; i < n; i++) x[i] = x[i] < ub[i] ? x[i] : ub[i]; }
static void clamp_lo_vec(int n, double* x, const double* lb) { for (int i = 0; i < n; i++) x[i] = x[i] > lb[i] ? x[i] : lb[i]; }
static void clamp_lo(int n, double* x, double lb) { for (int i = 0; i < n; i++) x[i] = x[i] > lb ? x[i] : lb; }
static void clamp_hi(int n, double* x, double ub) { for (int i = 0; i < n; i++) x[i] = x[i] < ub ? x[i] : ub; }

static double now_seconds(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---------------------------------------------------------------- formulate */
/* CupdlpWrapper.cpp:280-448: classify rows with the +-1e20 thresholds, add one
 * slack column (coefficient -1, rhs 0, bounds = row bounds) per BOUND/free row,
 * order rows EQ/BOUND first then LEQ/GEQ, negate LEQ rows, c *= sense, and map
 * bounds beyond +-1e20 to +-inf. */
int orc_formulate(const orc_lp* lp, orc_form* f) {
  memset(f, 0, sizeof(*f));
  const int n0 = lp->n, m = lp->m, nnz0 = lp->start[n0];
  int n = n0, nnz = nnz0, neq = 0;
  f->row_type = NEW(int, m);
  f->row_new_idx = NEW(int, m);
  for (int i = 0; i < m; i++) {
    int has_lo = lp->row_lower[i] > -1e20, has_up = lp->row_upper[i] < 1e20;
    if (has_lo && has_up && lp->row_lower[i] == lp->row_upper[i]) {
      f->row_type[i] = ORC_EQ; neq++;
    } else if (has_lo && !has_up) {
      f->row_type[i] = ORC_GEQ;
    } else if (!has_lo && has_up) {
      f->row_type[i] = ORC_LEQ;
    } else {               /* ranged, or free treated as ranged (:328-345) */
      f->row_type[i] = ORC_BOUND; n++; nnz++; neq++;
    }
  }
  f->n = n; f->m = m; f->nnz = nnz; f->neq = neq; f->n_orig = n0;
  f->sense = lp->sense; f->offset = lp->offset;
  f->cost = NEW(double, n); f->lower = NEW(double, n); f->upper = NEW(double, n);
  f->rhs = NEW(double, m);
  f->cbeg = NEW(int, n + 1); f->cidx = NEW(int, nnz); f->cval = NEW(double, nnz);
  for (int j = 0; j < n0; j++) {
    f->cost[j] = lp->cost[j] * lp->sense;
    f->lower[j] = lp->col_lower[j];
    f->upper[j] = lp->col_upper[j];
  }
  for (int j = n0; j < n; j++) f->cost[j] = 0.0;
  for (int i = 0, j = n0; i < m; i++)
    if (f->row_type[i] == ORC_BOUND) { f->lower[j] = lp->row_lower[i]; f->upper[j] = lp->row_upper[i]; j++; }
  for (int j = 0; j < n; j++) {
    if (f->lower[j] < -1e20) f->lower[j] = -INFINITY;
    if (f->upper[j] > 1e20) f->upper[j] = INFINITY;
  }
  int k = 0;
  for (int i = 0; i < m; i++) {
    if (f->row_type[i] == ORC_EQ) { f->rhs[k] = lp->row_lower[i]; f->row_new_idx[i] = k++; }
    else if (f->row_type[i] == ORC_BOUND) { f->rhs[k] = 0.0; f->row_new_idx[i] = k++; }
  }
  for (int i = 0; i < m; i++) {
    if (f->row_type[i] == ORC_LEQ) { f->rhs[k] = -lp->row_upper[i]; f->row_new_idx[i] = k++; }
    else if (f->row_type[i] == ORC_GEQ) { f->rhs[k] = lp->row_lower[i]; f->row_new_idx[i] = k++; }
  }
  for (int j = 0; j <= n0; j++) f->cbeg[j] = lp->start[j];
  for (int j = n0 + 1; j <= n; j++) f->cbeg[j] = f->cbeg[j - 1] + 1;
  k = 0;
  for (int j = 0; j < n0; j++) {
    for (int p = lp->start[j]; p < lp->start[j + 1]; p++) {
      int t = f->row_type[lp->index[p]];
      if (t == ORC_EQ || t == ORC_BOUND) { f->cidx[k] = f->row_new_idx[lp->index[p]]; f->cval[k] = lp->value[p]; k++; }
    }
    for (int p = lp->start[j]; p < lp->start[j + 1]; p++) {
      int t = f->row_type[lp->index[p]];
      if (t == ORC_LEQ) { f->cidx[k] = f->row_new_idx[lp->index[p]]; f->cval[k] = -lp->value[p]; k++; }
      else if (t == ORC_GEQ) { f->cidx[k] = f->row_new_idx[lp->index[p]]; f->cval[k] = lp->value[p]; k++; }
    }
  }
  for (int i = 0, j = n0; i < m; i++)
    if (f->row_type[i] == ORC_BOUND) { f->cidx[f->cbeg[j]] = f->row_new_idx[i]; f->cval[f->cbeg[j]] = -1.0; j++; }
  /* Init_Scaling, cupdlp_scaling.c:395-425: norms of the UNSCALED data */
  f->col_scale = NEW(double, n); f->row_scale = NEW(double, m);
  for (int j = 0; j < n; j++) f->col_scale[j] = 1.0;
  for (int i = 0; i < m; i++) f->row_scale[i] = 1.0;
  f->norm_cost = vnrm2(n, f->cost);
  f->norm_rhs = vnrm2(m, f->rhs);
  return 0;
}

/* ------------------------------------------------------------------ scaling */
/* scale_problem, cupdlp_scaling.c:17-45 */
static void apply_scaling(orc_form* f, const double* cs, const double* rs) {
  vdiv(f->n, f->cost, cs);
  vmul(f->n, f->lower, cs);
  vmul(f->n, f->upper, cs);
  vdiv(f->m, f->rhs, rs);
  for (int p = 0; p < f->cbeg[f->n]; p++) f->cval[p] /= rs[f->cidx[p]];
  for (int j = 0; j < f->n; j++)
    for (int p = f->cbeg[j]; p < f->cbeg[j + 1]; p++) f->cval[p] /= cs[j];
  vmul(f->n, f->col_scale, cs);
  vmul(f->m, f->row_scale, rs);
}

/* PDHG_Scale_Data with Init_Scaling's hard-wired choice: 10 Ruiz (inf-norm)
 * passes then Pock-Chambolle alpha=1 (cupdlp_scaling.c:47-120,174-231,233-393). */
void orc_scale(orc_form* f, int scaling) {
  if (scaling) {
    double* cs = NEW(double, f->n);
    double* rs = NEW(double, f->m);
    for (int it = 0; it < 10; it++) {
      memset(cs, 0, sizeof(double) * (size_t)f->n);
      memset(rs, 0, sizeof(double) * (size_t)f->m);
      for (int j = 0; j < f->n; j++) {
        double mx = 0.0;
        for (int p = f->cbeg[j]; p < f->cbeg[j + 1]; p++) { double a = fabs(f->cval[p]); if (a > mx) mx = a; }
        cs[j] = (f->cbeg[j] == f->cbeg[j + 1]) ? 0.0 : sqrt(mx);
      }
      for (int j = 0; j < f->n; j++) if (cs[j] == 0.0) cs[j] = 1.0;
      for (int p = 0; p < f->cbeg[f->n]; p++) { double a = fabs(f->cval[p]); if (rs[f->cidx[p]] < a) rs[f->cidx[p]] = a; }
      for (int i = 0; i < f->m; i++) rs[i] = rs[i] == 0.0 ? 1.0 : sqrt(rs[i]);
      apply_scaling(f, cs, rs);
    }
    const double alpha = 1.0;
    memset(cs, 0, sizeof(double) * (size_t)f->n);
    memset(rs, 0, sizeof(double) * (size_t)f->m);
    if (f->m > 0) {
      for (int j = 0; j < f->n; j++) {
        for (int p = f->cbeg[j]; p < f->cbeg[j + 1]; p++) cs[j] += pow(fabs(f->cval[p]), alpha);
        cs[j] = sqrt(pow(cs[j], 1.0 / alpha));
        if (cs[j] == 0.0) cs[j] = 1.0;
      }
      for (int p = 0; p < f->cbeg[f->n]; p++) rs[f->cidx[p]] += pow(fabs(f->cval[p]), 2.0 - alpha);
      for (int i = 0; i < f->m; i++) {
        rs[i] = sqrt(pow(rs[i], 1.0 / (2.0 - alpha)));
        if (rs[i] == 0.0) rs[i] = 1.0;
      }
    }
    apply_scaling(f, cs, rs);
    free(cs); free(rs);
  }
  /* problem_alloc, CupdlpWrapper.cpp:559-560: max |a| of the scaled matrix */
  double mx = 0.0;
  for (int p = 0; p < f->nnz; p++) { double a = fabs(f->cval[p]); if (a > mx) mx = a; }
  f->amax = mx;
}

/* csc2csr via counting-sort transpose (cupdlp_cs.c:189-214) */
void orc_build_csr(orc_form* f) {
  free(f->rbeg); free(f->ridx); free(f->rval);
  f->rbeg = NEW(int, f->m + 1); f->ridx = NEW(int, f->nnz); f->rval = NEW(double, f->nnz);
  int* w = NEW(int, f->m);
  for (int p = 0; p < f->nnz; p++) w[f->cidx[p]]++;
  int acc = 0;
  for (int i = 0; i < f->m; i++) { f->rbeg[i] = acc; acc += w[i]; w[i] = f->rbeg[i]; }
  f->rbeg[f->m] = acc;
  for (int j = 0; j < f->n; j++)
    for (int p = f->cbeg[j]; p < f->cbeg[j + 1]; p++) { int q = w[f->cidx[p]]++; f->ridx[q] = j; f->rval[q] = f->cval[p]; }
  free(w);
}

void orc_form_free(orc_form* f) {
  free(f->cost); free(f->lower); free(f->upper); free(f->rhs);
  free(f->cbeg); free(f->cidx); free(f->cval); free(f->rbeg); free(f->ridx); free(f->rval);
  free(f->col_scale); free(f->row_scale); free(f->row_new_idx); free(f->row_type);
  memset(f, 0, sizeof(*f));
}

/* AxCPU: column scatter over CSC (cupdlp_linalg.c:17-24,35-71) */
void orc_ax(const orc_form* f, const double* x, double* ax) {
  memset(ax, 0, sizeof(double) * (size_t)f->m);
  for (int j = 0; j < f->n; j++)
    for (int p = f->cbeg[j]; p < f->cbeg[j + 1]; p++) ax[f->cidx[p]] += f->cval[p] * x[j];
}
/* ATyCPU: row scatter over CSR (cupdlp_linalg.c:26-33,73-109) */
void orc_aty(const orc_form* f, const double* y, double* aty) {
  memset(aty, 0, sizeof(double) * (size_t)f->n);
  for (int i = 0; i < f->m; i++)
    for (int p = f->rbeg[i]; p < f->rbeg[i + 1]; p++) aty[f->ridx[p]] += f->rval[p] * y[i];
}

/* -------------------------------------------------------------- solver state */
typedef struct {
  const orc_form* f;
  int n, m, neq;
  double *x[2], *y[2], *ax[2], *aty[2];
  double *xs, *ys, *xa, *ya, *axa, *atya, *xlr, *ylr;
  double *has_lo, *has_up, *lo_f, *up_f;
  double *sp, *sn, *spa, *sna;              /* dSlackPos/Neg (+Average) */
  double *bn, *bm, *bn2;                    /* scratch */
  double tau, sigma, beta, sum_step;        /* dSumPrimalStep == dSumDualStep always */
  int step_iter, last_restart_iter, iter;
  /* resobj */
  double pobj, dobj, pfeas, dfeas, gap, relgap;
  double pobja, dobja, pfeasa, dfeasa, gapa, relgapa;
  double pfeas_lr, dfeas_lr, gap_lr, pfeas_lc, dfeas_lc, gap_lc;
  double pinf_obj, pinf_res, dinf_obj, dinf_res, pinf_obja, pinf_resa, dinf_obja, dinf_resa;
} orc_state;

/* PDHG_Compute_Primal_Feasibility, cupdlp_solver.c:12-67 */
static void primal_feas(orc_state* s, const double* ax, const double* x, double* feas, double* obj) {
  const orc_form* f = s->f;
  *obj = vdot(s->n, x, f->cost) * f->sense + f->offset;
  double* r = s->bm;
  vcopy(s->m, r, ax);
  vaxpy(s->m, -1.0, f->rhs, r);
  clamp_hi(s->m - s->neq, r + s->neq, 0.0);
  vmul(s->m, r, f->row_scale);
  *feas = vnrm2(s->m, r);
}

/* PDHG_Compute_Dual_Feasibility, cupdlp_solver.c:69-204 (CPU branch) */
static void dual_feas(orc_state* s, const double* aty, const double* y, double* feas, double* obj,
                      double* sp, double* sn) {
  const orc_form* f = s->f;
  double d = vdot(s->m, y, f->rhs);
  double* rc = s->bn;
  vcopy(s->n, rc, aty);
  vscale(s->n, -1.0, rc);
  vaxpy(s->n, 1.0, f->cost, rc);
  vcopy(s->n, sp, rc);
  clamp_lo(s->n, sp, 0.0);
  vmul(s->n, sp, s->has_lo);
  d += vdot(s->n, sp, s->lo_f);
  vcopy(s->n, sn, rc);
  clamp_hi(s->n, sn, 0.0);
  vscale(s->n, -1.0, sn);
  vmul(s->n, sn, s->has_up);
  d -= vdot(s->n, sn, s->up_f);
  *obj = d * f->sense + f->offset;
  vaxpy(s->n, -1.0, sp, rc);
  vaxpy(s->n, 1.0, sn, rc);
  vmul(s->n, rc, f->col_scale);
  *feas = vnrm2(s->n, rc);
}

/* PDHG_Compute_Residuals, cupdlp_solver.c:473-529 */
static void compute_residuals(orc_state* s) {
  int k = s->iter % 2;
  primal_feas(s, s->ax[k], s->x[k], &s->pfeas, &s->pobj);
  dual_feas(s, s->aty[k], s->y[k], &s->dfeas, &s->dobj, s->sp, s->sn);
  primal_feas(s, s->axa, s->xa, &s->pfeasa, &s->pobja);
  dual_feas(s, s->atya, s->ya, &s->dfeasa, &s->dobja, s->spa, s->sna);
  s->gap = s->pobj - s->dobj;
  s->relgap = fabs(s->pobj - s->dobj) / (1.0 + fabs(s->pobj) + fabs(s->dobj));
  s->gapa = s->pobja - s->dobja;
  s->relgapa = fabs(s->pobja - s->dobja) / (1.0 + fabs(s->pobja) + fabs(s->dobja));
}

/* PDHG_Compute_Primal_Infeasibility, cupdlp_solver.c:206-311 (CPU branch) */
static void primal_infeas(orc_state* s, const double* y, const double* sp, const double* sn, const double* aty,
                          double dobj, double* obj, double* res) {
  const orc_form* f = s->f;
  double ny = vdot(s->m, y, y), np = vdot(s->n, sp, sp), nn = vdot(s->n, sn, sn);
  double sc = sqrt(ny + np + nn);
  if (sc < 1e-8) sc = 1.0;
  double *lbray = s->bn, *ubray = s->bn2;
  vcopy(s->n, lbray, sp); vscale(s->n, 1 / sc, lbray);
  vcopy(s->n, ubray, sn); vscale(s->n, 1 / sc, ubray);
  *obj = (dobj - f->offset) / f->sense / sc;
  double* c = NEW(double, s->n);
  vcopy(s->n, c, aty);
  vscale(s->n, 1.0 / sc, c);
  vaxpy(s->n, 1.0, lbray, c);
  vaxpy(s->n, -1.0, ubray, c);
  vmul(s->n, c, f->col_scale);
  *res = vnrm2(s->n, c);
  free(c);
}

/* PDHG_Compute_Dual_Infeasibility, cupdlp_solver.c:313-429 (CPU branch) */
static void dual_infeas(orc_state* s, const double* x, const double* ax, double pobj, double* obj, double* res) {
  const orc_form* f = s->f;
  double* ray = s->bn;
  vcopy(s->n, ray, x);
  double sc = vnrm2(s->n, ray);
  if (sc < 1e-8) sc = 1.0;
  vscale(s->n, 1.0 / sc, ray);
  *obj = (pobj - f->offset) / f->sense / sc;
  double* c = s->bm;
  vcopy(s->m, c, ax);
  vscale(s->m, 1.0 / sc, c);
  clamp_hi(s->m - s->neq, c + s->neq, 0.0);
  vmul(s->m, c, f->row_scale);
  double rc = vdot(s->m, c, c);
  double* bd = s->bn2;
  vcopy(s->n, bd, ray); clamp_hi(s->n, bd, 0.0); vmul(s->n, bd, s->has_lo); vdiv(s->n, bd, f->col_scale);
  double rl = vdot(s->n, bd, bd);
  vcopy(s->n, bd, ray); clamp_lo(s->n, bd, 0.0); vmul(s->n, bd, s->has_up); vdiv(s->n, bd, f->col_scale);
  double ru = vdot(s->n, bd, bd);
  *res = sqrt(rc + rl + ru);
}

/* PDHG_Compute_Infeas_Residuals, cupdlp_solver.c:433-471 */
static void compute_infeas(orc_state* s) {
  int k = s->iter % 2;
  primal_infeas(s, s->y[k], s->sp, s->sn, s->aty[k], s->dobj, &s->pinf_obj, &s->pinf_res);
  dual_infeas(s, s->x[k], s->ax[k], s->pobj, &s->dinf_obj, &s->dinf_res);
  primal_infeas(s, s->ya, s->spa, s->sna, s->atya, s->dobja, &s->pinf_obja, &s->pinf_resa);
  dual_infeas(s, s->xa, s->axa, s->pobja, &s->dinf_obja, &s->dinf_resa);
}

/* PDHG_Compute_Average_Iterate, cupdlp_step.c:377-420 */
static void compute_average(orc_state* s) {
  double w = s->sum_step > 0.0 ? 1.0 / s->sum_step : 1.0;
  vcopy(s->n, s->xa, s->xs); vcopy(s->m, s->ya, s->ys);
  vscale(s->n, w, s->xa); vscale(s->m, w, s->ya);
  orc_ax(s->f, s->xa, s->axa);
  orc_aty(s->f, s->ya, s->atya);
}

/* PDHG_Restart_Score_GPU, cupdlp_restart.c:113-124 */
static double score(double b, double pf, double df, double gap) { return sqrt(b * pf * pf + df * df / b + gap * gap); }

/* PDHG_Check_Restart_GPU, cupdlp_restart.c:3-99. returns 0 none, 1 average, 2 current */
static int check_restart(orc_state* s) {
  if (s->iter == s->last_restart_iter) {
    s->pfeas_lr = s->pfeas; s->dfeas_lr = s->dfeas; s->gap_lr = s->gap;
    s->pfeas_lc = s->pfeas; s->dfeas_lc = s->dfeas; s->gap_lc = s->gap;
    return 0;
  }
  double mu_cur = score(s->beta, s->pfeas, s->dfeas, s->gap);
  double mu_avg = score(s->beta, s->pfeasa, s->dfeasa, s->gapa);
  int choice = mu_cur < mu_avg ? 2 : 1;
  double mu_cand = mu_cur < mu_avg ? mu_cur : mu_avg;
  if ((s->iter - s->last_restart_iter) >= 0.36 * s->iter) {
    /* artificial restart */
  } else {
    double mu_lr = score(s->beta, s->pfeas_lr, s->dfeas_lr, s->gap_lr);
    if (mu_cand < 0.2 * mu_lr) {
      /* sufficient decay */
    } else {
      double mu_lc = score(s->beta, s->pfeas_lc, s->dfeas_lc, s->gap_lc);
      if (mu_cand < 0.8 * mu_lr && mu_cand > mu_lc) {
        /* necessary decay */
      } else {
        choice = 0;
      }
    }
  }
  if (mu_cur < mu_avg) { s->pfeas_lc = s->pfeas; s->dfeas_lc = s->dfeas; s->gap_lc = s->gap; }
  else { s->pfeas_lc = s->pfeasa; s->dfeas_lc = s->dfeasa; s->gap_lc = s->gapa; }
  return choice;
}

/* PDHG_Compute_Step_Size_Ratio, cupdlp_step.c:147-176 */
static void step_size_ratio(orc_state* s) {
  int k = s->iter % 2;
  double mean = sqrt(s->tau * s->sigma);
  double* b = s->bn;
  vcopy(s->n, b, s->x[k]); vaxpy(s->n, -1.0, s->xlr, b);
  double dx = vnrm2(s->n, b);
  b = s->bm;
  vcopy(s->m, b, s->y[k]); vaxpy(s->m, -1.0, s->ylr, b);
  double dy = vnrm2(s->m, b);
  if (fmin(dx, dy) > 1e-10) {
    double upd = dy / dx;
    double lg = 0.5 * log(upd) + 0.5 * log(sqrt(s->beta));
    s->beta = exp(lg) * exp(lg);
  }
  s->tau = mean / sqrt(s->beta);
  s->sigma = s->tau * s->beta;
}

/* PDHG_Restart_Iterate_GPU, cupdlp_proj.c:88-148 */
static int restart_iterate(orc_state* s) {
  int choice = check_restart(s);
  if (choice == 0) return 0;
  int k = s->iter % 2;
  s->sum_step = 0.0;
  memset(s->xs, 0, sizeof(double) * (size_t)s->n);
  memset(s->ys, 0, sizeof(double) * (size_t)s->m);
  if (choice == 1) {
    s->pfeas_lr = s->pfeasa; s->dfeas_lr = s->dfeasa; s->gap_lr = s->gapa;
    vcopy(s->n, s->x[k], s->xa); vcopy(s->m, s->y[k], s->ya);
    vcopy(s->m, s->ax[k], s->axa); vcopy(s->n, s->aty[k], s->atya);
  } else {
    s->pfeas_lr = s->pfeas; s->dfeas_lr = s->dfeas; s->gap_lr = s->gap;
  }
  step_size_ratio(s);
  vcopy(s->n, s->xlr, s->x[k]); vcopy(s->m, s->ylr, s->y[k]);
  s->last_restart_iter = s->iter;
  compute_residuals(s);
  return choice;
}

/* PDHG_primalGradientStep (CPU branch), cupdlp_step.c:16-40 */
static void primal_step(orc_state* s, double* xn, const double* x, const double* aty, double tau) {
  const orc_form* f = s->f;
  vcopy(s->n, xn, x);
  vaxpy(s->n, -tau, f->cost, xn);
  vaxpy(s->n, tau, aty, xn);
  clamp_hi_vec(s->n, xn, f->upper);
  clamp_lo_vec(s->n, xn, f->lower);
}
/* PDHG_dualGradientStep (CPU branch), cupdlp_step.c:43-69 */
static void dual_step(orc_state* s, double* yn, const double* y, const double* ax, const double* axn, double sigma) {
  const orc_form* f = s->f;
  vcopy(s->m, yn, y);
  vaxpy(s->m, sigma, f->rhs, yn);
  vaxpy(s->m, -2.0 * sigma, axn, yn);
  vaxpy(s->m, sigma, ax, yn);
  clamp_lo(s->m - s->neq, yn + s->neq, 0.0);
}

/* cupdlp_compute_interaction_and_movement (CPU branch), cupdlp_linalg.c:772-801 */
static void movement_interaction(orc_state* s, double* mov, double* inter, int row_side) {
  int k = s->iter % 2, k1 = (s->iter + 1) % 2;
  double rb = sqrt(s->beta);
  double* b = s->bn;
  vcopy(s->n, b, s->x[k]); vaxpy(s->n, -1.0, s->x[k1], b);
  double dx = vdot(s->n, b, b);
  double* bm = s->bm;
  vcopy(s->m, bm, s->y[k]); vaxpy(s->m, -1.0, s->y[k1], bm);
  double dy = vdot(s->m, bm, bm);
  double* b2 = s->bn2;
  vcopy(s->n, b2, s->aty[k]); vaxpy(s->n, -1.0, s->aty[k1], b2);
  *inter = vdot(s->n, b, b2);
  if (row_side) {   /* Δy'(AΔx): same number in exact arithmetic (cupdlp_step.c:259-264) */
    double acc = 0.0;
    for (int i = 0; i < s->m; i++) acc += (s->ax[k][i] - s->ax[k1][i]) * bm[i];
    *inter = acc;
  }
  *mov = dx * 0.5 * rb + dy / (2.0 * rb);
}

/* PDHG_Update_Iterate_Adaptive_Step_Size, cupdlp_step.c:215-310.
 * returns 1 if the time limit fired on a rejected step (CUPDLP_CHECK_TIMEOUT) */
static int adaptive_update(orc_state* s, double t_begin, double t_lim, int row_side) {
  int k = s->iter % 2, k1 = (s->iter + 1) % 2;
  double eta = sqrt(s->tau * s->sigma);
  int done = 0;
  while (!done) {
    ++s->step_iter;
    double tau = eta / sqrt(s->beta), sigma = eta * sqrt(s->beta);
    primal_step(s, s->x[k1], s->x[k], s->aty[k], tau);
    orc_ax(s->f, s->x[k1], s->ax[k1]);
    dual_step(s, s->y[k1], s->y[k], s->ax[k], s->ax[k1], sigma);
    orc_aty(s->f, s->y[k1], s->aty[k1]);
    double mov, inter;
    movement_interaction(s, &mov, &inter, row_side);
    double lim = inter != 0.0 ? mov / fabs(inter) : INFINITY;
    if (eta <= lim) done = 1;
    else if (t_lim > 0 && now_seconds() - t_begin > t_lim) return 1;
    double first = (1.0 - pow(s->step_iter + 1.0, -0.3)) * lim;
    double second = (1.0 + pow(s->step_iter + 1.0, -0.6)) * eta;
    eta = fmin(first, second);
  }
  s->tau = eta / sqrt(s->beta);
  s->sigma = eta * sqrt(s->beta);
  return 0;
}

/* PDHG_Update_Iterate_Constant_Step_Size, cupdlp_step.c:178-206 */
static void constant_update(orc_state* s) {
  int k = s->iter % 2, k1 = (s->iter + 1) % 2;
  orc_ax(s->f, s->x[k], s->ax[k]);
  orc_aty(s->f, s->y[k], s->aty[k]);
  primal_step(s, s->x[k1], s->x[k], s->aty[k], s->tau);
  orc_ax(s->f, s->x[k1], s->ax[k1]);
  dual_step(s, s->y[k1], s->y[k], s->ax[k], s->ax[k1], s->sigma);
  orc_aty(s->f, s->y[k1], s->aty[k1]);
}

/* PDHG_Update_Average, cupdlp_step.c:422-442 (weight = sqrt of the NEW tau*sigma) */
static void update_average(orc_state* s) {
  int k1 = (s->iter + 1) % 2;
  double w = sqrt(s->tau * s->sigma);
  vaxpy(s->n, w, s->x[k1], s->xs);
  vaxpy(s->m, w, s->y[k1], s->ys);
  s->sum_step += w;
}

/* PDHG_Power_Method, cupdlp_step.c:71-145 */
static double power_method(orc_state* s) {
  double* q = s->bm;
  double lambda = 0.0;
  for (int i = 0; i < s->m; i++) q[i] = 1.0;
  for (int it = 0; it < 20; it++) {
    orc_aty(s->f, q, s->aty[0]);
    orc_ax(s->f, s->aty[0], s->ax[0]);
    vcopy(s->m, q, s->ax[0]);
    double qn = vnrm2(s->m, q);
    vscale(s->m, 1.0 / qn, q);
    orc_aty(s->f, q, s->aty[0]);
    lambda = vdot(s->n, s->aty[0], s->aty[0]);
  }
  return lambda;
}

static void trace_row(orc_result* r, const orc_state* s, int restart) {
  if (!r->trace || r->trace_len >= r->trace_cap) return;
  double* t = r->trace + (size_t)r->trace_len * ORC_TRACE_COLS;
  t[0] = s->iter; t[1] = s->pobj; t[2] = s->dobj; t[3] = s->pfeas; t[4] = s->dfeas;
  t[5] = s->pobja; t[6] = s->dobja; t[7] = s->pfeasa; t[8] = s->dfeasa;
  t[9] = s->tau; t[10] = s->sigma; t[11] = s->beta; t[12] = restart; t[13] = s->step_iter;
  t[14] = s->sum_step; t[15] = 0;
  r->trace_len++;
}

/* LP_SolvePDHG = PreSolve + PDHG_Solve + PostSolve, cupdlp_solver.c:899-1498 */
int orc_solve(const orc_lp* lp, const orc_params* p, orc_result* r) {
  orc_form F;
  orc_formulate(lp, &F);
  orc_scale(&F, p->scaling);
  orc_build_csr(&F);
  orc_state S; memset(&S, 0, sizeof(S));
  orc_state* s = &S;
  s->f = &F; s->n = F.n; s->m = F.m; s->neq = F.neq;
  const int n = F.n, m = F.m;
  for (int k = 0; k < 2; k++) { s->x[k] = NEW(double, n); s->y[k] = NEW(double, m); s->ax[k] = NEW(double, m); s->aty[k] = NEW(double, n); }
  s->xs = NEW(double, n); s->ys = NEW(double, m); s->xa = NEW(double, n); s->ya = NEW(double, m);
  s->axa = NEW(double, m); s->atya = NEW(double, n); s->xlr = NEW(double, n); s->ylr = NEW(double, m);
  s->has_lo = NEW(double, n); s->has_up = NEW(double, n); s->lo_f = NEW(double, n); s->up_f = NEW(double, n);
  s->sp = NEW(double, n); s->sn = NEW(double, n); s->spa = NEW(double, n); s->sna = NEW(double, n);
  s->bn = NEW(double, n); s->bn2 = NEW(double, n); s->bm = NEW(double, m);
  /* problem_alloc / resobj_Alloc: CupdlpWrapper.cpp:574-575, cupdlp_utils.c:885-886 */
  for (int j = 0; j < n; j++) {
    s->has_lo[j] = F.lower[j] > -INFINITY ? 1.0 : 0.0;
    s->has_up[j] = F.upper[j] < INFINITY ? 1.0 : 0.0;
    s->lo_f[j] = F.lower[j] > -INFINITY ? F.lower[j] : 0.0;
    s->up_f[j] = F.upper[j] < INFINITY ? F.upper[j] : 0.0;
  }
  s->pinf_res = s->dinf_res = s->pinf_resa = s->dinf_resa = 1.0;
  r->term_code = ORC_TIMELIMIT_OR_ITERLIMIT; r->term_iterate = 0; r->trace_len = 0;

  /* PDHG_PreSolve (hot start), cupdlp_solver.c:1217-1279 */
  int has_vars = (r->value_valid + r->dual_valid) != 0;
  if (r->value_valid && r->dual_valid) {
    int jc = 0;
    for (; jc < F.n_orig; jc++) s->x[0][jc] = r->col_value[jc];
    for (int i = 0; i < m; i++) {
      double mu = F.row_type[i] == ORC_LEQ ? -1 : 1;
      s->y[0][F.row_new_idx[i]] = F.sense * mu * r->row_dual[i];
      if (F.row_type[i] == ORC_BOUND) s->x[0][jc++] = r->row_value[i];
    }
    vmul(n, s->x[0], F.col_scale);
    vmul(m, s->y[0], F.row_scale);
  }

  double t_begin = now_seconds();
  double t_lim = (p->time_limit > 0 && isfinite(p->time_limit)) ? p->time_limit : 0.0;

  /* PDHG_Init_Step_Sizes, cupdlp_step.c:312-375 */
  {
    double a = vdot(n, F.cost, F.cost), b = vdot(m, F.rhs, F.rhs);
    if (!p->adaptive_step) {
      double lambda = power_method(s);
      s->beta = fmin(a, b) > 1e-6 ? a / b : 1.0;
      s->tau = 0.8 / sqrt(lambda);
      s->sigma = s->tau;
      s->tau /= sqrt(s->beta);
      s->sigma *= sqrt(s->beta);
    } else {
      s->beta = fmin(a, b) > 1e-6 ? a / b : 1.0;
      s->tau = (1.0 / F.amax) / sqrt(s->beta);
      s->sigma = s->tau * s->beta;
    }
    s->last_restart_iter = 0; s->sum_step = 0.0;
  }
  /* PDHG_Init_Variables, cupdlp_solver.c:531-591 */
  s->iter = 0;
  if (!has_vars) memset(s->x[0], 0, sizeof(double) * (size_t)n);
  clamp_hi_vec(n, s->x[0], F.upper); clamp_lo_vec(n, s->x[0], F.lower);
  if (!has_vars) memset(s->y[0], 0, sizeof(double) * (size_t)m);
  orc_ax(&F, s->x[0], s->ax[0]);
  orc_aty(&F, s->y[0], s->aty[0]);
  memset(s->xs, 0, sizeof(double) * (size_t)n); memset(s->ys, 0, sizeof(double) * (size_t)m);
  memset(s->xa, 0, sizeof(double) * (size_t)n); memset(s->ya, 0, sizeof(double) * (size_t)m);
  clamp_hi_vec(n, s->xs, F.upper); clamp_lo_vec(n, s->xs, F.lower);
  clamp_hi_vec(n, s->xa, F.upper); clamp_lo_vec(n, s->xa, F.lower);
  memset(s->xlr, 0, sizeof(double) * (size_t)n); memset(s->ylr, 0, sizeof(double) * (size_t)m);

  /* main loop, cupdlp_solver.c:939-1106 */
  const double tol_p = p->tol_primal * (1.0 + F.norm_rhs), tol_d = p->tol_dual * (1.0 + F.norm_cost);
  for (s->iter = 0; s->iter < p->iter_limit; ++s->iter) {
    double elapsed = now_seconds() - t_begin;
    int timed_out = t_lim > 0 && elapsed > t_lim;
    int checking = (s->iter < 10) || (s->iter == p->iter_limit - 1) || timed_out || !(s->iter % 40);
    if (checking) {
      compute_average(s);
      compute_residuals(s);
      compute_infeas(s);
      int k = s->iter % 2;
      if (s->pfeas < tol_p && s->dfeas < tol_d && s->relgap < p->tol_gap) {
        r->term_iterate = 0; r->term_code = ORC_OPTIMAL; trace_row(r, s, 0); break;
      }
      if (s->pfeasa < tol_p && s->dfeasa < tol_d && s->relgapa < p->tol_gap) {
        vcopy(n, s->x[k], s->xa); vcopy(m, s->y[k], s->ya);
        vcopy(m, s->ax[k], s->axa); vcopy(n, s->aty[k], s->atya);
        vcopy(n, s->sp, s->spa); vcopy(n, s->sn, s->sna);
        r->term_iterate = 1; r->term_code = ORC_OPTIMAL; trace_row(r, s, 0); break;
      }
      /* PDHG_Check_Infeasibility, cupdlp_solver.c:740-795, dFeasTol = 1e-8 */
      {
        const double ft = 1e-8;
        int inf = 0;
        if (s->pinf_obj > 0.0 && s->pinf_res < ft * s->pinf_obj) inf = 1;
        if (s->dinf_obj < 0.0 && s->dinf_res < -ft * s->dinf_obj) inf = 1;
        if (s->pinf_obja > 0.0 && s->pinf_resa < ft * s->pinf_obja) inf = 1;
        if (s->dinf_obja < 0.0 && s->dinf_resa < -ft * s->dinf_obja) inf = 1;
        if (inf) { r->term_code = ORC_INFEASIBLE_OR_UNBOUNDED; trace_row(r, s, 0); break; }
      }
      if (timed_out) { r->term_code = ORC_TIMELIMIT_OR_ITERLIMIT; trace_row(r, s, 0); break; }
      if (s->iter >= p->iter_limit - 1) { r->term_code = ORC_TIMELIMIT_OR_ITERLIMIT; trace_row(r, s, 0); break; }
      int rs = p->restart ? restart_iterate(s) : 0;
      trace_row(r, s, rs);
    }
    if (p->adaptive_step) {
      if (adaptive_update(s, t_begin, t_lim, p->interaction_row_side)) { r->term_code = ORC_TIMELIMIT_OR_ITERLIMIT; break; }
    } else {
      constant_update(s);
    }
    update_average(s);
  }
  r->iters = s->iter;
  if (r->term_code == ORC_OPTIMAL && r->term_iterate == 1) {
    r->pobj = s->pobja; r->dobj = s->dobja; r->pfeas = s->pfeasa; r->dfeas = s->dfeasa; r->gap = s->gapa; r->relgap = s->relgapa;
  } else {
    r->pobj = s->pobj; r->dobj = s->dobj; r->pfeas = s->pfeas; r->dfeas = s->dfeas; r->gap = s->gap; r->relgap = s->relgap;
  }

  /* PDHG_PostSolve, cupdlp_solver.c:1281-1435 */
  {
    int k = s->iter % 2;
    double *x = s->x[k], *y = s->y[k], *ax = s->ax[k], *aty = s->aty[k];
    vdiv(n, x, F.col_scale); vdiv(m, y, F.row_scale);
    vmul(n, s->sp, F.col_scale); vmul(n, s->sn, F.col_scale);
    vmul(m, ax, F.row_scale); vmul(n, aty, F.col_scale);
    for (int j = 0; j < F.n_orig; j++) r->col_value[j] = x[j];
    for (int i = 0; i < m; i++) r->row_value[i] = ax[F.row_new_idx[i]];
    for (int i = 0, j = 0; i < m; i++) {
      if (F.row_type[i] == ORC_LEQ) r->row_value[i] = -r->row_value[i];
      else if (F.row_type[i] == ORC_BOUND) { r->row_value[i] = r->row_value[i] + x[F.n_orig + j]; j++; }
    }
    for (int j = 0; j < F.n_orig; j++) r->col_dual[j] = s->sp[j] - s->sn[j];
    vscale(F.n_orig, F.sense, r->col_dual);
    for (int i = 0; i < m; i++) r->row_dual[i] = y[F.row_new_idx[i]];
    vscale(m, F.sense, r->row_dual);
    for (int i = 0; i < m; i++) if (F.row_type[i] == ORC_LEQ) r->row_dual[i] = -r->row_dual[i];
    r->value_valid = 1; r->dual_valid = 1;
  }
  for (int k = 0; k < 2; k++) { free(s->x[k]); free(s->y[k]); free(s->ax[k]); free(s->aty[k]); }
  free(s->xs); free(s->ys); free(s->xa); free(s->ya); free(s->axa); free(s->atya); free(s->xlr); free(s->ylr);
  free(s->has_lo); free(s->has_up); free(s->lo_f); free(s->up_f);
  free(s->sp); free(s->sn); free(s->spa); free(s->sna); free(s->bn); free(s->bn2); free(s->bm);
  orc_form_free(&F);
  return 0;
}
