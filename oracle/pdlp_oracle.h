/* oracle/pdlp_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, single thread) of the reference's PDLP path
 *   Highs::run -> solveLp -> solveLpCupdlp -> LP_SolvePDHG
 * (/root/reference/highs/pdlp/CupdlpWrapper.cpp, highs/pdlp/cupdlp/ sources).
 * It reproduces the reference's floating-point operation ORDER so that it can
 * be pinned bit-for-bit against oracle/_ref (the unmodified reference built by
 * oracle/build_ref.py); see tests/test_oracle_vs_ref.py and tests/golden/.
 * Parity status: PINNED (iteration counts and objectives of the reference's own
 * goldens, check/TestPdlp.cpp:11-64 and check/CMakeLists.txt:321-335).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
 * this library; the product (highs_b200/csrc) never links or calls it.
 */
#ifndef PDLP_ORACLE_H_
#define PDLP_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* row classes, numbering of cupdlp_defs.h (EQ=0, LEQ=1, GEQ=2, BOUND=3) */
enum { ORC_EQ = 0, ORC_LEQ = 1, ORC_GEQ = 2, ORC_BOUND = 3 };

/* termination codes, numbering of cupdlp_defs.h:61-68 */
enum {
  ORC_OPTIMAL = 0, ORC_INFEASIBLE = 1, ORC_UNBOUNDED = 2,
  ORC_INFEASIBLE_OR_UNBOUNDED = 3, ORC_TIMELIMIT_OR_ITERLIMIT = 4, ORC_FEASIBLE = 5
};

typedef struct {
  int n, m;                 /* columns, rows of the HighsLp */
  const int* start;         /* column-wise A: start[n+1], index[nnz], value[nnz] */
  const int* index;
  const double* value;
  const double *cost, *col_lower, *col_upper, *row_lower, *row_upper;
  double sense;             /* +1 minimise, -1 maximise */
  double offset;
} orc_lp;

typedef struct {
  int iter_limit;           /* pdlp_iteration_limit */
  double tol_primal, tol_dual, tol_gap;
  double time_limit;        /* seconds; <=0 or inf: none */
  int scaling;              /* 1 = Ruiz(10)+Pock-Chambolle(1) */
  int adaptive_step;        /* 1 = adaptive line search, 0 = fixed (power method) */
  int restart;              /* 1 = cuPDLP "GPU" restart scheme */
  int interaction_row_side; /* 0 = reference: (x-x').(A'y-A'y'), cupdlp_linalg.c:797; 1 = (Ax-Ax').(y-y'), the
                               algebraically identical form cupdlp_step.c:259-264 calls dInteractiony -- what the
                               multi-GPU engine uses; not bit-identical to the reference */
} orc_params;

/* standard form produced by formulate+scale (host arrays owned by the struct) */
typedef struct {
  int n, m, nnz, neq, n_orig;
  double *cost, *lower, *upper, *rhs;
  int* cbeg; int* cidx; double* cval;   /* CSC of the (scaled) matrix */
  int* rbeg; int* ridx; double* rval;   /* CSR copy (counting-sort transpose) */
  double *col_scale, *row_scale;
  int *row_new_idx, *row_type;          /* indexed by ORIGINAL row */
  double sense, offset;
  double norm_cost, norm_rhs;           /* of the UNSCALED formulated data */
  double amax;                          /* max |a_ij| after scaling */
} orc_form;

#define ORC_TRACE_COLS 16
/* one trace row per check iteration:
 * 0 iter, 1 pobj, 2 dobj, 3 pfeas, 4 dfeas, 5 pobjAvg, 6 dobjAvg, 7 pfeasAvg,
 * 8 dfeasAvg, 9 tau, 10 sigma, 11 beta, 12 restart(0 none,1 avg,2 cur),
 * 13 nStepSizeIter, 14 sumStep, 15 reserved */

typedef struct {
  double *col_value, *col_dual, *row_value, *row_dual;  /* caller-allocated n,n,m,m */
  int value_valid, dual_valid;   /* in: hot start flags; out: validity */
  int term_code;                 /* ORC_* */
  int term_iterate;              /* 0 last, 1 average */
  int iters;
  double pobj, dobj, pfeas, dfeas, gap, relgap;  /* of the returned iterate */
  double* trace;                 /* optional [trace_cap][ORC_TRACE_COLS] */
  int trace_cap, trace_len;
} orc_result;

int orc_formulate(const orc_lp* lp, orc_form* f);            /* CupdlpWrapper.cpp:280-448 */
void orc_scale(orc_form* f, int scaling);                    /* cupdlp_scaling.c:233-425 */
void orc_build_csr(orc_form* f);                             /* cupdlp_cs.c:189-214 */
void orc_form_free(orc_form* f);
void orc_ax(const orc_form* f, const double* x, double* ax);   /* cupdlp_linalg.c:35-71 */
void orc_aty(const orc_form* f, const double* y, double* aty); /* cupdlp_linalg.c:73-109 */
int orc_solve(const orc_lp* lp, const orc_params* p, orc_result* r);

#ifdef __cplusplus
}
#endif
#endif
