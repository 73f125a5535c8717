/* include/b200pdlp.h -- C ABI of the Blackwell-native PDLP engine (libb200pdlp.so).
 *
 * This is the drop-in boundary underneath HiGHS's
 *     HighsStatus solveLpCupdlp(HighsLpSolverObject&)
 * (/root/reference/highs/pdlp/CupdlpWrapper.cpp:23-278, called from
 * /root/reference/highs/lp_data/HighsSolve.cpp:97-104).  Plain pointers and
 * sizes only -- no HiGHS, torch or C++ types -- so the same library is bound
 * from the C++ shim in highs_b200/csrc/highs_shim.cpp and from
 * Python (ctypes, highs_b200/engine.py).
 *
 * All arithmetic is IEEE fp64; indices are 32-bit (cupdlp_int,
 * highs/pdlp/cupdlp/glbopts.h:249-253; HighsInt without HIGHSINT64).
 * Every entry point returns 0 on success and a negative b200pdlp_error
 * otherwise; b200pdlp_last_error() gives the message.  There is NO CPU
 * fallback: without a CUDA device the compute entry points fail with
 * B200PDLP_ERR_CUDA.
 */
#ifndef B200PDLP_H_
#define B200PDLP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200PDLP_VERSION 100

typedef enum {
  B200PDLP_OK = 0,
  B200PDLP_ERR_ARG = -1,     /* bad argument */
  B200PDLP_ERR_CUDA = -2,    /* CUDA runtime / no device */
  B200PDLP_ERR_NCCL = -3,
  B200PDLP_ERR_ALLOC = -4,
  B200PDLP_ERR_STATE = -5
} b200pdlp_error;

/* termination codes: numbering of cupdlp_defs.h:61-68 */
typedef enum {
  B200PDLP_OPTIMAL = 0,
  B200PDLP_INFEASIBLE = 1,
  B200PDLP_UNBOUNDED = 2,
  B200PDLP_INFEASIBLE_OR_UNBOUNDED = 3,
  B200PDLP_TIMELIMIT_OR_ITERLIMIT = 4,
  B200PDLP_FEASIBLE = 5
} b200pdlp_term;

/* HighsLp members read by solveLpCupdlp (HighsLp.h:23-35, CupdlpWrapper.cpp:280-300).
 * The matrix is column-wise (asserted by the caller, Highs.cpp:4130). */
typedef struct {
  int32_t num_col, num_row;
  const int32_t* a_start;    /* [num_col+1] */
  const int32_t* a_index;    /* [nnz] row indices */
  const double* a_value;     /* [nnz] */
  const double* col_cost;    /* [num_col] */
  const double* col_lower;   /* [num_col]; <= -1e20 means -inf (CupdlpWrapper.cpp:375-378) */
  const double* col_upper;
  const double* row_lower;   /* [num_row]; thresholds +-1e20 (CupdlpWrapper.cpp:316-317) */
  const double* row_upper;
  double sense;              /* +1 minimise, -1 maximise (ObjSense) */
  double offset;
} b200pdlp_lp;

/* What getUserParamsFromOptions (CupdlpWrapper.cpp:642-717) hands to cuPDLP-C. */
typedef struct {
  int32_t iter_limit;        /* pdlp_iteration_limit (N_ITER_LIM) */
  double tol_primal;         /* D_PRIMAL_TOL */
  double tol_dual;           /* D_DUAL_TOL */
  double tol_gap;            /* D_GAP_TOL */
  double time_limit;         /* D_TIME_LIM, seconds; < 0 or inf = none; 0 = already over (the first check ends the run) */
  int32_t scaling;           /* IF_SCALING: 1 = Ruiz x10 + Pock-Chambolle(1) */
  int32_t adaptive_step;     /* E_LINE_SEARCH_METHOD: 1 adaptive, 0 fixed (power method) */
  int32_t restart;           /* E_RESTART_METHOD: 1 on, 0 off */
  int32_t log_level;         /* N_LOG_LEVEL 0/1/2 */
  int32_t check_interval;    /* CUPDLP_RELEASE_INTERVAL; 0 -> 40 */
  int32_t device;            /* CUDA device ordinal; -1 = current */
  int32_t graph_passes;      /* PDHG passes captured per CUDA graph; 0 -> check_interval */
  int32_t ordered_max;       /* problems with max(cols,rows) <= this reduce in the reference's
                                sequential order (bit-identical trajectories); 0 -> 4096, <0 -> off */
  int32_t device_scaling;    /* where the prologue (formulate, scaling, transposition, layouts) runs: 0 = on the device for one
                                GPU in tree mode (device_prep.cu; host threads otherwise), -1 = always host threads,
                                1 / 2 = round-1 staged variants (host formulate + device scaling [+ device ELL fill]) */
  int32_t reserved[2];
} b200pdlp_params;

/* Hot start = incoming HighsSolution when value_valid && dual_valid
 * (PDHG_PreSolve, cupdlp_solver.c:1217-1279).  NULL pointer = cold start. */
typedef struct {
  const double* col_value;   /* [num_col] */
  const double* row_value;   /* [num_row] (slack of BOUND rows) */
  const double* row_dual;    /* [num_row] */
} b200pdlp_warm;

typedef struct {
  /* caller-allocated HighsSolution storage, original LP space, HiGHS signs
   * (PDHG_PostSolve, cupdlp_solver.c:1281-1435) */
  double* col_value;         /* [num_col] */
  double* col_dual;          /* [num_col] */
  double* row_value;         /* [num_row] */
  double* row_dual;          /* [num_row] */
  int32_t value_valid, dual_valid;
  int32_t term_code;         /* b200pdlp_term */
  int32_t term_iterate;      /* 0 last iterate, 1 average iterate */
  int32_t iters;             /* -> HighsInfo::pdlp_iteration_count */
  int32_t passes;            /* PDHG passes executed (iters + rejected line-search steps) */
  int32_t restarts;
  int32_t kernel_launches;   /* kernels this library launched inside the solve */
  /* cuPDLP-C's own residuals of the returned iterate (unscaled space) */
  double primal_obj, dual_obj, primal_feas, dual_feas, gap, rel_gap;
  /* timings, seconds */
  double setup_seconds;      /* formulate + scale + layout + H2D */
  double solve_seconds;      /* PDHG loop, host wall clock */
  double iter_device_ms;     /* CUDA-event time of the PDHG passes only */
  double loop_device_ms;     /* CUDA-event time of the whole loop: passes + check iterations + restarts */
  /* form dimensions actually solved */
  int32_t form_cols, form_rows, form_nnz, form_neq;
  /* optional trajectory trace, one row of B200PDLP_TRACE_COLS doubles per check iteration:
   * 0 iter, 1 pobj, 2 dobj, 3 pfeas, 4 dfeas (current iterate), 5-8 same for the average,
   * 9 tau, 10 sigma, 11 beta (after a restart's update), 12 restart (0 none, 1 average, 2 current),
   * 13 nStepSizeIter, 14 sum of step weights, 15 reserved */
  double* trace;             /* caller-allocated [trace_cap][16] or NULL */
  int32_t trace_cap, trace_len;
} b200pdlp_result;
#define B200PDLP_TRACE_COLS 16

/* ---- whole solve: the function the HiGHS shim calls ------------------------ */
void b200pdlp_default_params(b200pdlp_params* p);
int b200pdlp_solve(const b200pdlp_lp* lp, const b200pdlp_params* params,
                   const b200pdlp_warm* warm, b200pdlp_result* out);

/* The same solve on `ngpus` devices of THIS process (devices[g] or, if NULL, ordinals 0..ngpus-1): one problem per device
 * wired with b200pdlp_p2p_link_local, one host thread per rank, no NCCL.  ngpus <= 1 is b200pdlp_solve.
 * (what the HiGHS shim uses when B200PDLP_GPUS > 1; the one-process-per-GPU entry points are further below) */
int b200pdlp_solve_multi(const b200pdlp_lp* lp, const b200pdlp_params* params, const b200pdlp_warm* warm,
                         b200pdlp_result* out, int32_t ngpus, const int32_t* devices);

/* ---- persistent problem handle (bench + kernel-level parity tests) --------- */
typedef struct b200pdlp_problem b200pdlp_problem;

/* formulate + scale on the host, build the blocked row-major layouts of A and
 * A^T, upload to HBM.  rank/world select a contiguous, nnz-balanced block of
 * the (permuted) rows for multi-GPU; pass 0/1 for a single GPU. */
int b200pdlp_problem_create(const b200pdlp_lp* lp, const b200pdlp_params* params,
                            int32_t rank, int32_t world, b200pdlp_problem** out);
void b200pdlp_problem_destroy(b200pdlp_problem* p);
/* dims[0..7] = form cols, form rows (global), nnz (global), neq, local rows,
 * local row offset, local nnz, original cols */
int b200pdlp_problem_dims(const b200pdlp_problem* p, int32_t dims[8]);
/* host copies of the scaled standard form (for parity tests against the oracle):
 * which = 0 cost, 1 lower, 2 upper, 3 rhs, 4 col_scale, 5 row_scale (doubles);
 * returns number of doubles written (<= cap) */
int b200pdlp_problem_get_vector(const b200pdlp_problem* p, int32_t which, double* dst, int32_t cap);
/* CSR of the scaled matrix (global row numbering of this rank's rows) */
int b200pdlp_problem_get_csr(const b200pdlp_problem* p, int32_t* rowptr, int32_t* col, double* val);

/* single kernels on host vectors (H2D, kernel, D2H) -- parity tests only:
 * ax[m_local] = A_local x[n];  aty[n] = A_local^T y[m_local] */
int b200pdlp_spmv_ax(b200pdlp_problem* p, const double* x, double* ax);
int b200pdlp_spmv_aty(b200pdlp_problem* p, const double* y, double* aty);
/* single FUSED kernels of the hot path on host vectors in standard-form order (parity tests; single GPU).  Reference
 * analogues: the extern "C" launchers of highs/pdlp/cupdlp/cuda/cupdlp_cudalinalg.cu:13-386.
 *  primal step (K1): x_new = proj_[l,u](x - tau (c - aty))   (PDHG_primalGradientStep, cupdlp_step.c:16-40), *dx2 = |x - x_new|^2
 *  dual step (K2):   ax_new = A x_new;  y_new = y + sigma (b - 2 ax_new + ax), max(.,0) on inequality rows
 *                    (PDHG_dualGradientStep, cupdlp_step.c:43-69), *dy2 = |y - y_new|^2
 *  residuals:        out[10] = pobj, dobj, primal feasibility, dual feasibility, gap, relative gap, primal-ray objective
 *                    and residual, dual-ray objective and residual of (x, y) (cupdlp_solver.c:12-204, :206-471)
 * all on the SCALED standard form the problem holds (b200pdlp_problem_get_vector / get_csr return it). */
int b200pdlp_primal_step(b200pdlp_problem* p, const double* x, const double* aty, double tau, double* x_new, double* dx2);
int b200pdlp_dual_step(b200pdlp_problem* p, const double* x_new, const double* y, const double* ax, double sigma,
                       double* y_new, double* ax_new, double* dy2);
int b200pdlp_residuals(b200pdlp_problem* p, const double* x, const double* y, double out[10]);

/* device-pointer variants on the problem's stream (bench roofline loop):
 * time `reps` back-to-back launches with CUDA events; returns ms in *ms_total */
int b200pdlp_bench_spmv(b200pdlp_problem* p, int32_t which /*0 Ax, 1 ATy*/, int32_t reps, float* ms_total);

/* per-kernel CUDA-event timing of `reps` PDHG passes in their real sequence (K1 primal step,
 * K2 A x + dual step, K3 A'y + interaction [+ all-reduce and K3b when world > 1]); call after a
 * solve (it continues from that state).  ms[0..3] = summed ms of K1, K2, K3, rest */
int b200pdlp_bench_pass(b200pdlp_problem* p, int32_t reps, float ms[4]);

/* run the PDHG loop on an uploaded problem (bench "value": inputs resident in HBM) */
int b200pdlp_problem_solve(b200pdlp_problem* p, const b200pdlp_params* params,
                           const b200pdlp_warm* warm, b200pdlp_result* out);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink ---------------------- */
/* rank 0 obtains the id and ships it to the other ranks (torch.distributed /
 * any byte transport); every rank then calls comm_init.  id is 128 bytes. */
int b200pdlp_nccl_unique_id(uint8_t id[128]);
int b200pdlp_comm_init(b200pdlp_problem* p, const uint8_t id[128]);

/* fused NVLink path (optional, after comm_init): every rank exports CUDA-IPC handles of its exchange
 * buffers (B200PDLP_IPC_BYTES bytes), the host application all-gathers them (rank order) and every rank
 * imports the world x B200PDLP_IPC_BYTES blob.  From then on the per-iteration reduce-scatter and
 * all-gather run inside the engine's own kernels over peer memory; NCCL is used at check iterations only. */
#define B200PDLP_IPC_BYTES 256
int b200pdlp_p2p_export(b200pdlp_problem* p, uint8_t handles[B200PDLP_IPC_BYTES]);
int b200pdlp_p2p_import(b200pdlp_problem* p, const uint8_t* all_handles);
/* unmap the peers' buffers (call on every rank, then synchronise the ranks, BEFORE any rank destroys its problem) */
int b200pdlp_p2p_release(b200pdlp_problem* p);
/* G logical shards inside ONE process ("the same code path must run with G logical shards on 1
 * device"): probs[] holds one problem per rank of the same world, all created by this process (normally on the
 * same device); their buffers are wired to each other directly instead of through CUDA IPC, and no NCCL
 * communicator is needed.  Every rank's b200pdlp_problem_solve must then run on its own host thread. */
int b200pdlp_p2p_link_local(b200pdlp_problem** probs, int32_t count);
/* device-side timeline of the fused path since the last call, per-pass averages in us: [0] primal-shard phase,
 * [1] barrier 0 total, [2] of which waiting, [3] A x + A'y phase, [4] barrier 1 total, [5] of which waiting, [6] passes */
int b200pdlp_p2p_timeline(b200pdlp_problem* p, double out_us[8]);

/* ---- host-only view of the standard form (no GPU needed; parity tests) -------
 * formulate (CupdlpWrapper.cpp:280-448) + scale (cupdlp_scaling.c:233-425) only. */
typedef struct b200pdlp_form b200pdlp_form;
int b200pdlp_form_create(const b200pdlp_lp* lp, int32_t scaling, b200pdlp_form** out);
void b200pdlp_form_destroy(b200pdlp_form* f);
/* dims[0..4] = cols, rows, nnz, neq, original cols; scalars[0..2] = |c|_2, |b|_2 (unscaled), max|a_ij| (scaled) */
int b200pdlp_form_dims(const b200pdlp_form* f, int32_t dims[5], double scalars[3]);
/* which = 0 cost, 1 lower, 2 upper, 3 rhs, 4 col_scale, 5 row_scale; returns count written */
int b200pdlp_form_get_vector(const b200pdlp_form* f, int32_t which, double* dst, int32_t cap);
int b200pdlp_form_get_csc(const b200pdlp_form* f, int32_t* start, int32_t* index, double* value);
/* row-major copy (the reference's csc2csr, cupdlp_utils.c:1222-1254): rowptr[rows+1], col[nnz], val[nnz] */
int b200pdlp_form_get_csr(b200pdlp_form* f, int32_t* rowptr, int32_t* col, double* val);
/* row_new_idx[m], row_class[m] by ORIGINAL row (EQ=0, LEQ=1, GEQ=2, BOUND=3) */
int b200pdlp_form_get_row_map(const b200pdlp_form* f, int32_t* row_new_idx, int32_t* row_class);

/* HiPDLP mode (solver=hipdlp), host prologue only so far: PDLPSolver::preprocessLp
 * (hipdlp/pdhg.cc:152-358) + Scaling::scaleProblem (hipdlp/scaling.cc; scaling_mode bits 1 Ruiz, 4 Pock-Chambolle, 2 L2 =
 * HighsOptions::pdlp_scaling_mode, ruiz_iterations = pdlp_ruiz_iterations).  The form answers the b200pdlp_form_* getters;
 * rows carry an upper bound too: b200pdlp_form_get_vector(which = 6).  row_class uses 4 for FREE rows. */
int b200pdlp_hipdlp_form_create(const b200pdlp_lp* lp, int32_t scaling_mode, int32_t ruiz_iterations, b200pdlp_form** out);
/* The HiPDLP solve on one GPU: what a shim replacing pdlp/HiPdlpWrapper.cpp (solveLpHiPdlp, :26-141) forwards to.
 * Parameters are the HighsOptions the reference's setup() reads (hipdlp/pdhg.cc:1783-1874).  term_code is B200PDLP_OPTIMAL or
 * B200PDLP_TIMELIMIT_OR_ITERLIMIT (term_iterate = 2 when it was the time limit); like the reference, a run that does not
 * converge returns x = y = 0 (the iterate is copied out only on convergence, pdhg.cc:784-899).
 * STATUS (round 2): runs on hardware; tests/test_gpu_hipdlp.py (bit-exact against the oracle), tests/test_gpu_dropin.py (through
 * the shim), bench.py --solver hipdlp. */
typedef struct {
  double tolerance;           /* pdlp_optimality_tolerance, or kkt_tolerance when set */
  int32_t iter_limit;         /* pdlp_iteration_limit */
  int32_t scaling_mode;       /* pdlp_scaling_mode: bits 1 Ruiz, 2 L2, 4 Pock-Chambolle (default 5); 0 when scaling is off */
  int32_t ruiz_iterations;    /* pdlp_ruiz_iterations (default 10) */
  int32_t step_size_strategy; /* pdlp_step_size_strategy: 0 fixed primal weight, anything else PID */
  double time_limit;          /* seconds; < 0 or inf = none */
  int32_t ordered_max;        /* as in b200pdlp_params: checks add in the reference's order up to this size */
  int32_t device;             /* CUDA device ordinal; -1 = current */
  int32_t log_level;
  int32_t reserved[3];
} b200pdlp_hipdlp_params;
void b200pdlp_hipdlp_default_params(b200pdlp_hipdlp_params* p);
int b200pdlp_solve_hipdlp(const b200pdlp_lp* lp, const b200pdlp_hipdlp_params* params, b200pdlp_result* out);

/* host-only (no GPU; parity tests): the HOST control of the HiPDLP loop -- step sizes, fixed-point error, convergence test,
 * restart criteria, PID primal weight (HipController, host_prep_hipdlp.cpp; hipdlp/pdhg.cc:494-707,901-927,1944-2050) --
 * replayed over `nblocks` blocks of 40 steps from recorded sums: sums[b][9] = the nine sums of block b's closing check,
 * restart_sums[b][3] = the three fixed-point sums of block b's first step (used when block b follows a restart);
 * out[b][8] = restart decided, primal weight, primal step, dual step, fixed-point error, converged, iterations, halpern_iteration */
int b200pdlp_hipdlp_controller_replay(double norm_cost, double norm_rhs, double op_norm_sq, double tolerance, int32_t strategy,
                                       int32_t nblocks, const double* sums, const double* restart_sums, double* out);

/* PDLPSolver::powerMethod (hipdlp/pdhg.cc:1529-1671): 20 iterations on A A' from the ones vector -> estimate of |A|_2^2 */
int b200pdlp_hipdlp_power_method(const b200pdlp_form* f, double* lambda);

/* host-only (no GPU): build the device layout of rank `rank` of `world` exactly as b200pdlp_problem_create does
 * (row block, length-sorted device orderings, sliced ELL of A_g and A_g', long-row segments, segmented column
 * positions, the A_g' output map) and evaluate it on the host in the kernels' traversal order.
 * x[n], y[m] in standard-form order; ax[m]: rows [r0,r1) receive A_g x (others untouched); aty[n] = A_g' y_g.
 * stats: [0] r0 [1] r1 [2] ordered mode [3] A padded slots [4] A long rows [5] A segments [6] A' padded slots
 *        [7] A' long rows [8] A' segments [9] shard_len [10] seg_len [11] layout build time, ms */
int b200pdlp_form_layout_eval(b200pdlp_form* f, int32_t rank, int32_t world, int32_t ordered_max, const double* x,
                              const double* y, double* ax, double* aty, double stats[12]);

/* host-only helper (no GPU): nnz-balanced contiguous row partition of the
 * formulated LP; bounds[world+1] receives the row offsets */
int b200pdlp_partition_rows(const b200pdlp_lp* lp, int32_t world, int32_t* bounds);

/* Test entry point (needs a GPU): run the device-resident prologue (formulate + scale + row index + orderings +
 * sliced-ELL layouts as kernels, highs_b200/csrc/device_prep.cu) and its host twin (host_prep.cpp) on the same LP and
 * compare them array by array, bit for bit.  report[k] = number of mismatching entries:
 *  0-3 n, m, nnz, neq   4 cbeg 5 cidx 6 cval 7 cost 8 lower 9 upper 10 col_scale 11 rhs 12 row_scale (standard-form order)
 *  13 rptr 14 rpos (row-major index)  15 row_new_idx 16 row_class  17 row ordering 18 column ordering
 *  19-21 A: slice descriptors (-1: sizes differ), col, val   22-24 the same for A'   25, 26 long-row arrays of A, A'
 *  27 max|a_ij|  28, 29 relative error of |c|_2, |b|_2 (tree sums vs sequential)  30 vectors in device order
 *  31 (info) 1 if every column is stored with ascending rows */
int b200pdlp_debug_prep_compare(const b200pdlp_lp* lp, int32_t scaling, double report[32]);

/* The engine keeps the device blocks of finished solves in a process-wide cache (a solve allocates ~60 buffers; cudaMalloc /
 * cudaFree would cost more than the prologue).  This returns them to the driver.  B200PDLP_CACHE_MB caps the cache. */
void b200pdlp_release_cache(void);

/* ---- post-solve KKT assessment -------------------------------------------------
 * What Highs::run() does with the HighsSolution after solveLpCupdlp returns: lpKktCheck -> getKktFailures
 * (/root/reference/highs/lp_data/HighsSolution.cpp:1043-1327, :73-495): A x and A'y with double-double accumulation,
 * bound / sign violations, absolute and relative measures (relative to 1 + |active bounds|_inf, 1 + |costs with ~0 dual|_inf),
 * residual errors |A x - row_value|, |A'y - c + col_dual|, complementarity, P-D objective error, and the status rules for
 * a solution without a basis.  b200pdlp_kkt_check runs on the GPU (no CPU fallback); b200pdlp_kkt_check_host is its host
 * twin over the same per-variable arithmetic (highs_b200/csrc/kkt_logic.hpp), kept so that the logic is checked against
 * the reference's own lpKktCheck without a GPU.  Tolerances: the five HighsOptions after the kkt_tolerance override. */
typedef struct {
  double primal_feasibility_tolerance, dual_feasibility_tolerance, primal_residual_tolerance, dual_residual_tolerance,
         optimality_tolerance;
} b200pdlp_kkt_tolerances;
typedef struct {   /* HighsInfo fields of the same names (HighsInfo.h:91-131) */
  double objective_function_value, dual_objective_value, primal_dual_objective_error;
  double max_primal_infeasibility, sum_primal_infeasibilities, max_dual_infeasibility, sum_dual_infeasibilities;
  double max_relative_primal_infeasibility, max_relative_dual_infeasibility;
  double max_primal_residual_error, max_dual_residual_error, max_relative_primal_residual_error, max_relative_dual_residual_error;
  double max_complementarity_violation;
  double norm_bounds, norm_costs;   /* getKktFailures' highs_norm_bounds / highs_norm_costs */
  int32_t num_primal_infeasibilities, num_dual_infeasibilities, num_relative_primal_infeasibilities, num_relative_dual_infeasibilities;
  int32_t num_primal_residual_errors, num_dual_residual_errors, num_relative_primal_residual_errors, num_relative_dual_residual_errors;
  int32_t num_complementarity_violations;
  int32_t primal_solution_status, dual_solution_status;   /* kSolutionStatusInfeasible 1 / Feasible 2 */
  int32_t model_status;   /* in: HighsModelStatus code returned by the solver; out: after lpKktCheck's adjustment */
} b200pdlp_kkt_info;
void b200pdlp_kkt_default_tolerances(b200pdlp_kkt_tolerances* t, double kkt_tolerance /* <= 0: the default 1e-7 */);
int b200pdlp_kkt_check(const b200pdlp_lp* lp, const double* col_value, const double* col_dual, const double* row_value,
                       const double* row_dual, const b200pdlp_kkt_tolerances* tol, b200pdlp_kkt_info* inout);
int b200pdlp_kkt_check_host(const b200pdlp_lp* lp, const double* col_value, const double* col_dual, const double* row_value,
                            const double* row_dual, const b200pdlp_kkt_tolerances* tol, b200pdlp_kkt_info* inout);

/* Page-lock / unlock caller memory (cudaHostRegister) so that the copies of b200pdlp_solve run at PCIe speed; optional
 * (pageable buffers work, staged by the driver).  A host application that keeps its HighsLp / HighsSolution vectors for
 * many solves registers them once. */
int b200pdlp_host_register(void* ptr, size_t bytes);
int b200pdlp_host_unregister(void* ptr);

const char* b200pdlp_last_error(void);
int b200pdlp_version(void);
int b200pdlp_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200PDLP_H_ */
