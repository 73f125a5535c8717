// highs_b200/csrc/highs_pdlp_cleanup.hpp -- presolve / postsolve interplay of the PDLP drop-in.
//
// With the default presolve=choose, Highs::run() hands the REDUCED LP to solveLpCupdlp and postsolves its solution
// (/root/reference/highs/lp_data/Highs.cpp:1554-1694, :1802-1931).  Postsolve can magnify the residuals of a first-order
// solution, so that lpKktCheck downgrades kOptimal to kUnknown (HighsSolution.cpp:1303-1310).  The reference carries a
// remedy -- solve the ORIGINAL LP with PDLP from the postsolved solution -- but has it compiled out
// (Highs.cpp:1940 `consider_pdlp_cleanup = false`; body :1947-1982; decision rule Highs::tryPdlpCleanup,
// HighsInterface.cpp:4210-4272).  On the CPU the clean-up costs about as much as the first solve; with the B200 engine
// behind solver=pdlp it costs milliseconds, so this header provides it, driven through HiGHS's PUBLIC API (the reference
// library stays unmodified): same condition, same decision rule, same iteration limit, same bookkeeping.
//
// Header-only, needs only Highs.h; works with any HiGHS build (the CPU reference in the CPU tests, the drop-in library
// linked against libb200pdlp.so on the GPU box).
#pragma once
#include <algorithm>
#include <string>

#include "Highs.h"

struct B200PdlpCleanupReport {
  bool considered = false;           // the first run ended kUnknown without a basis (the reference's condition, :1941-1944)
  bool attempted = false;            // ... and the KKT errors were within the margin (tryPdlpCleanup returned true)
  double max_relative_violation = 0; // of the five KKT measures tryPdlpCleanup looks at
  HighsInt iteration_limit = 0;      // max(10000, first run's PDLP iterations / 10), or 1000 after IPX
  HighsInt first_pdlp_iterations = 0;
  HighsModelStatus first_status = HighsModelStatus::kNotset;
  HighsModelStatus final_status = HighsModelStatus::kNotset;
};

// Highs::tryPdlpCleanup (HighsInterface.cpp:4210-4272) on the public info / options of a finished run
// (tolerance_margin: the reference's constant 1e2.  lpKktCheck itself downgrades kOptimal -> kUnknown only beyond 100x
//  the tolerance, so with kkt_tolerance set the reference's margin rejects exactly the cases that need the clean-up --
//  presumably why it is compiled out; a host application may pass a wider margin.)
inline bool b200TryPdlpCleanup(const Highs& highs, HighsInt& pdlp_cleanup_iteration_limit, double& max_relative_violation,
                               const double tolerance_margin = 1e2) {
  const HighsInfo& info = highs.getInfo();
  const HighsOptions& options = highs.getOptions();
  max_relative_violation = 0;
  auto measure = [&](const double kkt_error, const double kkt_tolerance) {
    const double use_kkt_tolerance = options.kkt_tolerance != kDefaultKktTolerance ? options.kkt_tolerance : kkt_tolerance;
    max_relative_violation = std::max(kkt_error / use_kkt_tolerance, max_relative_violation);
  };
  measure(info.max_relative_primal_infeasibility, options.primal_feasibility_tolerance);
  measure(info.max_relative_dual_infeasibility, options.dual_feasibility_tolerance);
  measure(info.max_relative_primal_residual_error, options.primal_residual_tolerance);
  measure(info.max_relative_dual_residual_error, options.dual_residual_tolerance);
  measure(info.primal_dual_objective_error, options.optimality_tolerance);
  if (max_relative_violation > tolerance_margin) return false;   // too far off: not worth it
  if (info.pdlp_iteration_count > 0)
    pdlp_cleanup_iteration_limit = std::max(HighsInt(10000), HighsInt(info.pdlp_iteration_count / 10));
  else
    pdlp_cleanup_iteration_limit = 1000;   // IPX without crossover was used
  return true;
}

// Highs::run() followed, when the reference's condition holds, by the clean-up solve of Highs.cpp:1947-1982: PDLP on the
// original LP (presolve off) from the incumbent solution, with the clean-up iteration limit; options are restored and the
// first run's PDLP iterations are added to the count of the clean-up (:1978-1981 -- here reported, HighsInfo is const).
// tighten < 1: the clean-up solve runs with kkt_tolerance * tighten (the reference keeps the tolerance, with which a
// hot-started cuPDLP-C often stops at once: its own 2-norm criteria are already met by the postsolved point while
// lpKktCheck's infinity-norm measures are not)
inline HighsStatus b200RunWithPdlpCleanup(Highs& highs, B200PdlpCleanupReport* report = nullptr,
                                          const double tolerance_margin = 1e2, const double tighten = 1.0) {
  B200PdlpCleanupReport local;
  B200PdlpCleanupReport& r = report ? *report : local;
  r = B200PdlpCleanupReport();
  HighsStatus status = highs.run();
  r.first_status = r.final_status = highs.getModelStatus();
  r.first_pdlp_iterations = highs.getInfo().pdlp_iteration_count > 0 ? highs.getInfo().pdlp_iteration_count : 0;
  if (status == HighsStatus::kError) return status;
  const HighsOptions& options = highs.getOptions();
  // !basis_.valid && model_status_ == kUnknown && allow_pdlp_cleanup && !run_centring  (:1941-1944)
  r.considered = !highs.getBasis().valid && highs.getModelStatus() == HighsModelStatus::kUnknown && !options.run_centring &&
                 highs.getSolution().value_valid && highs.getSolution().dual_valid && !highs.getLp().isMip();
  if (!r.considered) return status;
  if (!b200TryPdlpCleanup(highs, r.iteration_limit, r.max_relative_violation, tolerance_margin)) return status;
  r.attempted = true;
  const std::string solver = options.solver, presolve = options.presolve;
  const HighsInt pdlp_iteration_limit = options.pdlp_iteration_limit;
  highs.setOptionValue("solver", kPdlpString);
  highs.setOptionValue("presolve", kHighsOffString);          // the ORIGINAL LP, hot-started from the incumbent solution
  highs.setOptionValue("pdlp_iteration_limit", r.iteration_limit);
  const double kkt_tolerance = options.kkt_tolerance;
  if (tighten != 1.0) {
    const double base = kkt_tolerance != kDefaultKktTolerance ? kkt_tolerance
                                                               : std::min(options.primal_feasibility_tolerance, options.dual_feasibility_tolerance);
    highs.setOptionValue("kkt_tolerance", base * tighten);
  }
  status = highs.run();
  if (tighten != 1.0) highs.setOptionValue("kkt_tolerance", kkt_tolerance);
  highs.setOptionValue("solver", solver);
  highs.setOptionValue("presolve", presolve);
  highs.setOptionValue("pdlp_iteration_limit", pdlp_iteration_limit);
  r.final_status = highs.getModelStatus();
  return status;
}
