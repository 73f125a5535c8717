// highs_b200/csrc/highs_shim.cpp -- the translation unit that replaces
// /root/reference/highs/pdlp/CupdlpWrapper.cpp when HiGHS is linked against the B200 engine.
//
// It defines the two symbols HiGHS calls on the solver=pdlp path,
//     HighsStatus solveLpCupdlp(HighsLpSolverObject&)                       (CupdlpWrapper.h:92)
//     HighsStatus solveLpCupdlp(const HighsOptions&, HighsTimer&, const HighsLp&, HighsBasis&,
//                               HighsSolution&, HighsModelStatus&, HighsInfo&, HighsCallback&)  (:94-98)
// (sole call site: highs/lp_data/HighsSolve.cpp:97-104) with the behaviour of
// CupdlpWrapper.cpp:30-278 -- reset status/info, map options to solver parameters (:642-717),
// hot start from a valid incoming HighsSolution, resize and fill the four solution vectors, set
// pdlp_iteration_count, invalidate the basis, map the termination code (:220-251) -- and forwards
// the numerical work to libb200pdlp.so through the C ABI of include/b200pdlp.h.
// It also keeps getCupdlpLogLevel (CupdlpWrapper.h:106), which HiPDLP's wrapper includes.
//
// This file needs the HiGHS headers, so it is compiled only where a HiGHS source tree is available
// (oracle/build_ref.py --shim builds oracle/_ref/libhighs_b200.so = reference objects minus
// CupdlpWrapper.o and the cuPDLP-C objects, plus this shim).  Unlike the
// reference GPU build it does NOT call cudaDeviceReset() (CupdlpWrapper.cpp:271): the engine frees
// what it allocated and leaves sibling CUDA contexts alone.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "lp_data/HighsLpSolverObject.h"
#include "lp_data/HighsSolution.h"

#include "b200pdlp.h"

HighsInt getCupdlpLogLevel(const HighsOptions& options) {
  if (options.output_flag) return options.log_dev_level ? 2 : 1;
  return 0;
}

HighsStatus solveLpCupdlp(const HighsOptions& options, HighsTimer& timer, const HighsLp& lp,
                          HighsBasis& highs_basis, HighsSolution& highs_solution,
                          HighsModelStatus& model_status, HighsInfo& highs_info, HighsCallback& callback) {
  (void)callback;  // unused by the reference too
  (void)timer;     // the reference computes the remaining time but passes the full limit (:701-707)
  resetModelStatusAndHighsInfo(model_status, highs_info);

  static_assert(sizeof(HighsInt) == 4, "the B200 engine uses 32-bit indices (build HiGHS without HIGHSINT64)");
  b200pdlp_lp clp;
  clp.num_col = lp.num_col_;
  clp.num_row = lp.num_row_;
  clp.a_start = lp.a_matrix_.start_.data();
  clp.a_index = lp.a_matrix_.index_.data();
  clp.a_value = lp.a_matrix_.value_.data();
  clp.col_cost = lp.col_cost_.data();
  clp.col_lower = lp.col_lower_.data();
  clp.col_upper = lp.col_upper_.data();
  clp.row_lower = lp.row_lower_.data();
  clp.row_upper = lp.row_upper_.data();
  clp.sense = lp.sense_ == ObjSense::kMaximize ? -1.0 : 1.0;
  clp.offset = lp.offset_;

  // getUserParamsFromOptions, CupdlpWrapper.cpp:642-717
  b200pdlp_params prm;
  b200pdlp_default_params(&prm);
  prm.iter_limit = (int32_t)std::min<int64_t>((int64_t)options.pdlp_iteration_limit, (int64_t)kHighsIInf32);
  prm.log_level = (int32_t)getCupdlpLogLevel(options);
  prm.scaling = (options.pdlp_features_off & kPdlpScalingOff) == 0 ? 1 : 0;
  prm.adaptive_step = (options.pdlp_features_off & kPdlpAdaptiveStepSizeOff) == 0 ? 1 : 0;
  prm.tol_primal = options.primal_feasibility_tolerance;
  prm.tol_dual = options.dual_feasibility_tolerance;
  prm.tol_gap = options.pdlp_optimality_tolerance;
  if (options.kkt_tolerance != kDefaultKktTolerance)
    prm.tol_primal = prm.tol_dual = prm.tol_gap = options.kkt_tolerance;
  prm.time_limit = options.time_limit < kHighsInf ? options.time_limit : -1.0;   // < 0 = none; 0 ends the run at the first check
  int restart_on = (options.pdlp_features_off & kPdlpRestartOff) == 0 ? 1 : 0;
  if (options.pdlp_cupdlpc_restart_method == 0) restart_on = 0;
  prm.restart = restart_on;

  // hot start iff the incoming solution is valid (PDHG_PreSolve, cupdlp_solver.c:1224-1227)
  b200pdlp_warm warm{nullptr, nullptr, nullptr};
  const bool hot = highs_solution.value_valid && highs_solution.dual_valid &&
                   (HighsInt)highs_solution.col_value.size() == lp.num_col_ &&
                   (HighsInt)highs_solution.row_value.size() == lp.num_row_ &&
                   (HighsInt)highs_solution.row_dual.size() == lp.num_row_;
  std::vector<double> w_col, w_row, w_dual;
  if (hot) {
    w_col = highs_solution.col_value;
    w_row = highs_solution.row_value;
    w_dual = highs_solution.row_dual;
    warm.col_value = w_col.data();
    warm.row_value = w_row.data();
    warm.row_dual = w_dual.data();
  }

  // the engine writes straight into the HighsSolution storage, like cuPDLP-C (:190-193)
  highs_solution.col_value.resize(lp.num_col_);
  highs_solution.row_value.resize(lp.num_row_);
  highs_solution.col_dual.resize(lp.num_col_);
  highs_solution.row_dual.resize(lp.num_row_);
  b200pdlp_result res{};
  res.col_value = highs_solution.col_value.data();
  res.col_dual = highs_solution.col_dual.data();
  res.row_value = highs_solution.row_value.data();
  res.row_dual = highs_solution.row_dual.data();

  // B200PDLP_GPUS=N: N devices of this process (row blocks + column shards, peer memory; b200pdlp_solve_multi)
  int ngpus = 1;
  if (const char* e = getenv("B200PDLP_GPUS")) ngpus = std::max(1, std::min(atoi(e), b200pdlp_device_count()));
  const int rc = ngpus > 1 ? b200pdlp_solve_multi(&clp, &prm, hot ? &warm : nullptr, &res, ngpus, nullptr)
                           : b200pdlp_solve(&clp, &prm, hot ? &warm : nullptr, &res);
  model_status = HighsModelStatus::kUnknown;
  highs_basis.valid = false;
  if (rc != B200PDLP_OK) {
    highsLogUser(options.log_options, HighsLogType::kError, "B200 PDLP engine failed: %s\n", b200pdlp_last_error());
    highs_solution.value_valid = false;
    highs_solution.dual_valid = false;
    model_status = HighsModelStatus::kSolveError;
    return HighsStatus::kError;
  }
  highs_info.pdlp_iteration_count = res.iters;
  highs_solution.value_valid = res.value_valid != 0;
  highs_solution.dual_valid = res.dual_valid != 0;
  switch (res.term_code) {   // CupdlpWrapper.cpp:225-245
    case B200PDLP_OPTIMAL: model_status = HighsModelStatus::kOptimal; break;
    case B200PDLP_INFEASIBLE: model_status = HighsModelStatus::kInfeasible; break;
    case B200PDLP_UNBOUNDED: model_status = HighsModelStatus::kUnbounded; break;
    case B200PDLP_INFEASIBLE_OR_UNBOUNDED: model_status = HighsModelStatus::kUnboundedOrInfeasible; break;
    case B200PDLP_TIMELIMIT_OR_ITERLIMIT:
      model_status = res.iters >= prm.iter_limit - 1 ? HighsModelStatus::kIterationLimit : HighsModelStatus::kTimeLimit;
      break;
    default: model_status = HighsModelStatus::kUnknown; break;
  }
  return HighsStatus::kOk;
}

HighsStatus solveLpCupdlp(HighsLpSolverObject& solver_object) {
  return solveLpCupdlp(solver_object.options_, solver_object.timer_, solver_object.lp_, solver_object.basis_,
                       solver_object.solution_, solver_object.model_status_, solver_object.highs_info_,
                       solver_object.callback_);
}
