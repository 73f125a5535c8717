// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never on the product path).
//
// Our own small driver around the UNMODIFIED reference library built by
// oracle/build_ref.py (oracle/_ref/libhighs_ref.so).  It loads an LP (from our
// .b2lp interchange file, or from an .mps through the reference's own reader),
// runs Highs::run() with whatever options are given (normally solver=pdlp
// presolve=off, i.e. the path Highs::run -> solveLp -> solveLpCupdlp,
// /root/reference/highs/lp_data/HighsSolve.cpp:94-117), and prints one JSON
// line with model status, pdlp_iteration_count, objective and the HighsInfo KKT
// fields that lpKktCheck fills (highs/lp_data/HighsSolution.cpp:1043-1320).
// Optionally dumps the LP (--dump-lp) and the HighsSolution (--sol) so that
// golden fixtures can be generated (tests/golden/make_golden.py).
//
// .b2lp layout (little endian): int64 magic 'B2LP', int64 n, m, nnz,
//   double sense (+1 min / -1 max), double offset,
//   double c[n], lo[n], up[n], rl[m], ru[m], int32 start[n+1], int32 index[nnz],
//   double value[nnz]   (column-wise, as HighsLp.a_matrix_)
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "Highs.h"
#include "lp_data/HighsSolution.h"

#include "../highs_b200/csrc/highs_pdlp_cleanup.hpp"   // --pdlp-cleanup: the product's clean-up flow under test

static const int64_t kMagic = 0x504C3242;  // "B2LP"

static bool readB2lp(const std::string& path, HighsLp& lp) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  int64_t hdr[4];
  double sd[2];
  if (fread(hdr, 8, 4, f) != 4 || hdr[0] != kMagic) return false;
  if (fread(sd, 8, 2, f) != 2) return false;
  int64_t n = hdr[1], m = hdr[2], nnz = hdr[3];
  lp.num_col_ = (HighsInt)n;
  lp.num_row_ = (HighsInt)m;
  lp.sense_ = sd[0] < 0 ? ObjSense::kMaximize : ObjSense::kMinimize;
  lp.offset_ = sd[1];
  lp.col_cost_.resize(n); lp.col_lower_.resize(n); lp.col_upper_.resize(n);
  lp.row_lower_.resize(m); lp.row_upper_.resize(m);
  lp.a_matrix_.format_ = MatrixFormat::kColwise;
  lp.a_matrix_.num_col_ = (HighsInt)n;
  lp.a_matrix_.num_row_ = (HighsInt)m;
  lp.a_matrix_.start_.resize(n + 1);
  lp.a_matrix_.index_.resize(nnz);
  lp.a_matrix_.value_.resize(nnz);
  bool ok = true;
  ok &= fread(lp.col_cost_.data(), 8, n, f) == (size_t)n;
  ok &= fread(lp.col_lower_.data(), 8, n, f) == (size_t)n;
  ok &= fread(lp.col_upper_.data(), 8, n, f) == (size_t)n;
  ok &= fread(lp.row_lower_.data(), 8, m, f) == (size_t)m;
  ok &= fread(lp.row_upper_.data(), 8, m, f) == (size_t)m;
  static_assert(sizeof(HighsInt) == 4, "oracle build uses 32-bit HighsInt");
  ok &= fread(lp.a_matrix_.start_.data(), 4, n + 1, f) == (size_t)(n + 1);
  ok &= fread(lp.a_matrix_.index_.data(), 4, nnz, f) == (size_t)nnz;
  ok &= fread(lp.a_matrix_.value_.data(), 8, nnz, f) == (size_t)nnz;
  fclose(f);
  return ok;
}

static bool writeB2lp(const std::string& path, const HighsLp& lp) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  HighsSparseMatrix a = lp.a_matrix_;
  a.ensureColwise();
  int64_t n = lp.num_col_, m = lp.num_row_, nnz = a.start_[n];
  int64_t hdr[4] = {kMagic, n, m, nnz};
  double sd[2] = {lp.sense_ == ObjSense::kMaximize ? -1.0 : 1.0, lp.offset_};
  fwrite(hdr, 8, 4, f); fwrite(sd, 8, 2, f);
  fwrite(lp.col_cost_.data(), 8, n, f);
  fwrite(lp.col_lower_.data(), 8, n, f);
  fwrite(lp.col_upper_.data(), 8, n, f);
  fwrite(lp.row_lower_.data(), 8, m, f);
  fwrite(lp.row_upper_.data(), 8, m, f);
  fwrite(a.start_.data(), 4, n + 1, f);
  fwrite(a.index_.data(), 4, nnz, f);
  fwrite(a.value_.data(), 8, nnz, f);
  fclose(f);
  return true;
}

// solution file: int64 n, m, int64 value_valid, dual_valid, then col_value[n],
// col_dual[n], row_value[m], row_dual[m]
static void writeSol(const std::string& path, const HighsSolution& s) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return;
  int64_t hdr[4] = {(int64_t)s.col_value.size(), (int64_t)s.row_value.size(),
                    s.value_valid ? 1 : 0, s.dual_valid ? 1 : 0};
  fwrite(hdr, 8, 4, f);
  fwrite(s.col_value.data(), 8, s.col_value.size(), f);
  fwrite(s.col_dual.data(), 8, s.col_dual.size(), f);
  fwrite(s.row_value.data(), 8, s.row_value.size(), f);
  fwrite(s.row_dual.data(), 8, s.row_dual.size(), f);
  fclose(f);
}

static bool readSol(const std::string& path, HighsSolution& s) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  int64_t hdr[4];
  if (fread(hdr, 8, 4, f) != 4) return false;
  s.col_value.resize(hdr[0]); s.col_dual.resize(hdr[0]);
  s.row_value.resize(hdr[1]); s.row_dual.resize(hdr[1]);
  bool ok = fread(s.col_value.data(), 8, hdr[0], f) == (size_t)hdr[0];
  ok &= fread(s.col_dual.data(), 8, hdr[0], f) == (size_t)hdr[0];
  ok &= fread(s.row_value.data(), 8, hdr[1], f) == (size_t)hdr[1];
  ok &= fread(s.row_dual.data(), 8, hdr[1], f) == (size_t)hdr[1];
  s.value_valid = hdr[2] != 0;
  s.dual_valid = hdr[3] != 0;
  fclose(f);
  return ok;
}

int main(int argc, char** argv) {
  std::string lp_path, mps_path, dump_lp, sol_path, warm_path, kkt_path;
  int kkt_status = (int)HighsModelStatus::kOptimal;
  std::vector<std::pair<std::string, std::string>> opts;
  bool quiet = true, pdlp_cleanup = false;
  double cleanup_margin = 1e2, cleanup_tighten = 1.0;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() { return std::string(i + 1 < argc ? argv[++i] : ""); };
    if (a == "--lp") lp_path = next();
    else if (a == "--mps") mps_path = next();
    else if (a == "--dump-lp") dump_lp = next();
    else if (a == "--sol") sol_path = next();
    else if (a == "--warm") warm_path = next();
    else if (a == "--kkt-of") kkt_path = next();
    else if (a == "--kkt-status") kkt_status = atoi(next().c_str());
    else if (a == "--verbose") quiet = false;
    else if (a == "--pdlp-cleanup") pdlp_cleanup = true;
    else if (a == "--cleanup-margin") cleanup_margin = atof(next().c_str());
    else if (a == "--cleanup-tighten") cleanup_tighten = atof(next().c_str());
    else if (a == "--opt") {
      std::string kv = next();
      size_t e = kv.find('=');
      if (e == std::string::npos) { fprintf(stderr, "bad --opt %s\n", kv.c_str()); return 2; }
      opts.push_back({kv.substr(0, e), kv.substr(e + 1)});
    } else { fprintf(stderr, "unknown arg %s\n", a.c_str()); return 2; }
  }
  Highs highs;
  if (quiet) highs.setOptionValue("output_flag", false);
  if (!mps_path.empty()) {
    if (highs.readModel(mps_path) != HighsStatus::kOk) { fprintf(stderr, "readModel failed\n"); return 3; }
  } else if (!lp_path.empty()) {
    HighsLp lp;
    if (!readB2lp(lp_path, lp)) { fprintf(stderr, "cannot read %s\n", lp_path.c_str()); return 3; }
    if (highs.passModel(lp) != HighsStatus::kOk) { fprintf(stderr, "passModel failed\n"); return 3; }
  } else { fprintf(stderr, "need --lp or --mps\n"); return 2; }
  if (!dump_lp.empty()) writeB2lp(dump_lp, highs.getLp());
  for (auto& kv : opts) {
    // try typed setters in turn: bool / int / double / string
    HighsOptionType t;
    if (highs.getOptionType(kv.first, t) != HighsStatus::kOk) { fprintf(stderr, "unknown option %s\n", kv.first.c_str()); return 2; }
    HighsStatus st = HighsStatus::kOk;
    if (t == HighsOptionType::kBool) st = highs.setOptionValue(kv.first, kv.second == "true" || kv.second == "1" || kv.second == "on");
    else if (t == HighsOptionType::kInt) st = highs.setOptionValue(kv.first, (HighsInt)atol(kv.second.c_str()));
    else if (t == HighsOptionType::kDouble) st = highs.setOptionValue(kv.first, atof(kv.second.c_str()));
    else st = highs.setOptionValue(kv.first, kv.second);
    if (st == HighsStatus::kError) { fprintf(stderr, "bad value for %s\n", kv.first.c_str()); return 2; }
  }
  if (!warm_path.empty()) {
    HighsSolution ws;
    if (!readSol(warm_path, ws)) { fprintf(stderr, "cannot read warm start\n"); return 3; }
    highs.setSolution(ws);
  }
  // --kkt-of: do not solve; evaluate a GIVEN HighsSolution with the reference's own
  // lpKktCheck (highs/lp_data/HighsSolution.cpp:1043-1320), i.e. exactly what Highs::run()
  // does after solveLpCupdlp returns (Highs.cpp:1990).  Used to measure the product's KKT
  // residuals with the reference's definitions.
  HighsInfo kkt_info;
  HighsModelStatus kkt_model_status = (HighsModelStatus)kkt_status;
  double secs = 0.0;
  HighsStatus rs = HighsStatus::kOk;
  B200PdlpCleanupReport cleanup;
  if (!kkt_path.empty()) {
    HighsSolution sol;
    if (!readSol(kkt_path, sol)) { fprintf(stderr, "cannot read solution\n"); return 3; }
    HighsBasis basis;
    kkt_info.invalidate();
    kkt_info.pdlp_iteration_count = -1;
    lpKktCheck(kkt_model_status, kkt_info, highs.getLp(), sol, basis, highs.getOptions(), "ref_driver --kkt-of");
  } else {
    auto t0 = std::chrono::steady_clock::now();
    rs = pdlp_cleanup ? b200RunWithPdlpCleanup(highs, &cleanup, cleanup_margin, cleanup_tighten) : highs.run();
    secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  const HighsInfo& info = kkt_path.empty() ? highs.getInfo() : kkt_info;
  const HighsModelStatus final_status = kkt_path.empty() ? highs.getModelStatus() : kkt_model_status;
  const HighsLp& lp = highs.getLp();
  if (!sol_path.empty()) writeSol(sol_path, highs.getSolution());
  printf("{\"run_status\": %d, \"model_status\": \"%s\", \"model_status_code\": %d, "
         "\"pdlp_iteration_count\": %d, \"objective_function_value\": %.17g, "
         "\"primal_dual_objective_error\": %.17g, "
         "\"num_primal_infeasibilities\": %d, \"max_primal_infeasibility\": %.17g, \"sum_primal_infeasibilities\": %.17g, "
         "\"num_dual_infeasibilities\": %d, \"max_dual_infeasibility\": %.17g, \"sum_dual_infeasibilities\": %.17g, "
         "\"max_relative_primal_infeasibility\": %.17g, \"max_relative_dual_infeasibility\": %.17g, "
         "\"max_primal_residual_error\": %.17g, \"max_dual_residual_error\": %.17g, "
         "\"max_relative_primal_residual_error\": %.17g, \"max_relative_dual_residual_error\": %.17g, "
         "\"max_complementarity_violation\": %.17g, \"num_complementarity_violations\": %d, "
         "\"primal_solution_status\": %d, \"dual_solution_status\": %d, "
         "\"cleanup_considered\": %d, \"cleanup_attempted\": %d, \"cleanup_iteration_limit\": %d, \"cleanup_first_status_code\": %d, "
         "\"cleanup_first_pdlp_iterations\": %d, \"cleanup_max_relative_violation\": %.17g, "
         "\"num_col\": %d, \"num_row\": %d, \"num_nz\": %d, \"run_seconds\": %.6f}\n",
         (int)rs, highs.modelStatusToString(final_status).c_str(), (int)final_status,
         (int)info.pdlp_iteration_count, info.objective_function_value,
         info.primal_dual_objective_error,
         (int)info.num_primal_infeasibilities, info.max_primal_infeasibility, info.sum_primal_infeasibilities,
         (int)info.num_dual_infeasibilities, info.max_dual_infeasibility, info.sum_dual_infeasibilities,
         info.max_relative_primal_infeasibility, info.max_relative_dual_infeasibility,
         info.max_primal_residual_error, info.max_dual_residual_error,
         info.max_relative_primal_residual_error, info.max_relative_dual_residual_error,
         info.max_complementarity_violation, (int)info.num_complementarity_violations,
         (int)info.primal_solution_status, (int)info.dual_solution_status,
         (int)cleanup.considered, (int)cleanup.attempted, (int)cleanup.iteration_limit, (int)cleanup.first_status,
         (int)cleanup.first_pdlp_iterations, cleanup.max_relative_violation,
         (int)lp.num_col_, (int)lp.num_row_, (int)lp.a_matrix_.numNz(), secs);
  return 0;
}
