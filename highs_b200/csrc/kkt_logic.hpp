// highs_b200/csrc/kkt_logic.hpp -- the per-variable arithmetic of HiGHS's post-solve KKT assessment, usable from host
// and device code.
//
// Restates, term for term,
//   getVariableKktFailures        /root/reference/highs/lp_data/HighsSolution.cpp:567-660
//   infeasibility()               highs/util/HighsUtils.h:219-272
//   the two passes of getKktFailures                      HighsSolution.cpp:73-495
//   getComplementarityViolations  :996-1032,  computeDualObjectiveValue :1345-1386
//   lpKktCheck's status rules for a solution without a basis (:1043-1327)
// kkt_check.cu runs them as kernels over the solution resident in HBM; the host twin (same header, sequential loops)
// exists so that this logic is checked on the CPU against the reference's own lpKktCheck (tests/test_kkt_host.py).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace b200 {

// double-double accumulator: what HighsCDouble does in productQuad / productTransposeQuad (HighsSparseMatrix.cpp:1220-1290):
// every product is rounded to double, the SUM is carried in two doubles (Knuth two-sum)
struct DD {
  double hi = 0.0, lo = 0.0;
  B200_HD void add(double a) {
    const double s = hi + a;
    const double bb = s - hi;
    const double err = (hi - (s - bb)) + (a - bb);
    hi = s;
    lo += err;
  }
  B200_HD void add(const DD& o) { add(o.hi); lo += o.lo; }
  B200_HD double value() const { return hi + lo; }
};

struct KktTolerances {   // lpKktCheck :1049-1061: kkt_tolerance, when set, overrides all five
  double primal_feasibility, dual_feasibility, primal_residual, dual_residual, optimality;
};

enum : uint8_t { kAtNo = 0, kAtLo = 1, kAtUp = 2 };

struct VarKkt {
  double primal_infeasibility, dual_infeasibility;
  uint8_t at_status, mid_status;
};

// getVariableKktFailures for a continuous variable (`dual` already multiplied by the sense)
B200_HD VarKkt variable_kkt(const KktTolerances& t, double lower, double upper, double value, double dual) {
  VarKkt r;
  // infeasibility(): the residual with zero tolerance, capped by the tolerance when nothing exceeds it (#2653)
  double infeas = 0.0, residual = 0.0;
  const double tol = t.primal_feasibility;
  if (value < lower - tol) infeas = lower - value;
  if (value > upper + tol) infeas = value - upper;
  if (tol > 0) {
    if (value < lower) residual = lower - value;
    if (value > upper) residual = value - upper;
  } else {
    residual = infeas;
  }
  if (infeas == 0) residual = residual < tol ? residual : tol;
  r.primal_infeasibility = residual;
  r.at_status = kAtNo;
  double bound_residual = std::fabs(lower - value);
  if (bound_residual * bound_residual <= tol) {
    r.at_status = kAtLo;
  } else {
    bound_residual = std::fabs(value - upper);
    if (bound_residual * bound_residual <= tol) r.at_status = kAtUp;
  }
  r.mid_status = kAtNo;
  r.dual_infeasibility = 0.0;
  if (lower < upper) {
    const double length = upper - lower;
    if (lower <= -INFINITY && upper >= INFINITY) {
      r.dual_infeasibility = std::fabs(dual);
    } else if (length * length > tol) {
      const double middle = (lower + upper) * 0.5;
      if (value < middle) { r.mid_status = kAtLo; r.dual_infeasibility = -dual > 0. ? -dual : 0.; }
      else { r.mid_status = kAtUp; r.dual_infeasibility = dual > 0. ? dual : 0.; }
    }
  }
  return r;
}

// what pass 1 of getKktFailures accumulates (one partial per thread / block / the whole problem)
struct KktSums {
  int num_primal_infeasibility = 0, num_dual_infeasibility = 0, num_relative_primal_infeasibility = 0,
      num_relative_dual_infeasibility = 0, num_primal_residual_error = 0, num_dual_residual_error = 0,
      num_relative_primal_residual_error = 0, num_relative_dual_residual_error = 0, num_complementarity_violation = 0;
  double max_primal_infeasibility = 0, max_dual_infeasibility = 0, max_relative_primal_infeasibility = 0,
         max_relative_dual_infeasibility = 0, max_primal_residual_error = 0, max_dual_residual_error = 0,
         max_relative_primal_residual_error = 0, max_relative_dual_residual_error = 0, max_complementarity_violation = 0;
  double sum_primal_infeasibility = 0, sum_dual_infeasibility = 0;
  double dual_objective = 0;   // sum of bound * dual (computeDualObjectiveValue, without the offset)
  double objective = 0;        // sum of cost * value (HighsLp::objectiveValue, without the offset)
};

B200_HD double kmax(double a, double b) { return a < b ? b : a; }

// pass 0 for one variable: contribution to highs_norm_bounds (|active bound|) and, for columns, highs_norm_costs
B200_HD void kkt_pass0(const KktTolerances& t, bool is_col, double cost, double lower, double upper, double value, double dual_in,
                       double sense, double& norm_bounds, double& norm_costs) {
  if (is_col && dual_in * dual_in < t.dual_feasibility) norm_costs = kmax(std::fabs(cost), norm_costs);   // :259-262 (unflipped dual)
  const VarKkt v = variable_kkt(t, lower, upper, value, dual_in * sense);
  if (v.at_status == kAtLo) norm_bounds = kmax(std::fabs(lower), norm_bounds);
  else if (v.at_status == kAtUp) norm_bounds = kmax(std::fabs(upper), norm_bounds);
}

// pass 1 for one variable.  activity_residual: |A x - row_value| for a row, |A'y - c + col_dual| for a column.
B200_HD void kkt_pass1(const KktTolerances& t, bool is_col, double cost, double lower, double upper, double value, double dual_in,
                       double sense, double norm_bounds, double norm_costs, double activity_residual, KktSums& s) {
  const double dual = dual_in * sense;
  const VarKkt v = variable_kkt(t, lower, upper, value, dual);
  if (v.primal_infeasibility > 0) {
    if (v.primal_infeasibility > t.primal_feasibility) s.num_primal_infeasibility++;
    s.max_primal_infeasibility = kmax(s.max_primal_infeasibility, v.primal_infeasibility);
    s.sum_primal_infeasibility += v.primal_infeasibility;
    double measure = norm_bounds;
    if (v.at_status == kAtNo) {
      if (v.mid_status == kAtNo || v.mid_status == kAtLo) measure = kmax(std::fabs(lower), measure);
      else measure = kmax(std::fabs(upper), measure);
    }
    const double rel = v.primal_infeasibility / (1.0 + measure);
    if (rel > t.primal_feasibility) s.num_relative_primal_infeasibility++;
    s.max_relative_primal_infeasibility = kmax(s.max_relative_primal_infeasibility, rel);
  }
  if (v.dual_infeasibility > 0) {
    if (v.dual_infeasibility > t.dual_feasibility) s.num_dual_infeasibility++;
    s.max_dual_infeasibility = kmax(s.max_dual_infeasibility, v.dual_infeasibility);
    s.sum_dual_infeasibility += v.dual_infeasibility;
    double measure = norm_costs;
    if (is_col && cost != 0.0 && dual * dual >= t.dual_feasibility) measure = kmax(std::fabs(cost), measure);
    const double rel = v.dual_infeasibility / (1.0 + measure);
    if (rel > t.dual_feasibility) s.num_relative_dual_infeasibility++;
    s.max_relative_dual_infeasibility = kmax(s.max_relative_dual_infeasibility, rel);
  }
  if (!is_col) {
    const double rel = activity_residual / (1.0 + norm_bounds);
    if (activity_residual > t.primal_residual) s.num_primal_residual_error++;
    s.max_primal_residual_error = kmax(s.max_primal_residual_error, activity_residual);
    if (rel > t.primal_residual) s.num_relative_primal_residual_error++;
    s.max_relative_primal_residual_error = kmax(s.max_relative_primal_residual_error, rel);
  } else {
    const double rel = activity_residual / (1.0 + norm_costs);
    if (activity_residual > t.dual_residual) s.num_dual_residual_error++;
    s.max_dual_residual_error = kmax(s.max_dual_residual_error, activity_residual);
    if (rel > t.dual_residual) s.num_relative_dual_residual_error++;
    s.max_relative_dual_residual_error = kmax(s.max_relative_dual_residual_error, rel);
    s.objective += cost * value;
  }
  // getComplementarityViolations and computeDualObjectiveValue use the UNFLIPPED dual
  double primal_residual, bound;
  if (lower <= -INFINITY && upper >= INFINITY) { primal_residual = 1; bound = 1; }
  else {
    const double mid = (lower + upper) * 0.5;
    primal_residual = value < mid ? std::fabs(lower - value) : std::fabs(upper - value);
    bound = value < mid ? lower : upper;
  }
  const double cv = primal_residual * std::fabs(dual_in);
  if (cv > t.optimality) s.num_complementarity_violation++;
  s.max_complementarity_violation = kmax(s.max_complementarity_violation, cv);
  s.dual_objective += bound * dual_in;
}

B200_HD void kkt_merge(KktSums& a, const KktSums& b) {
  a.num_primal_infeasibility += b.num_primal_infeasibility; a.num_dual_infeasibility += b.num_dual_infeasibility;
  a.num_relative_primal_infeasibility += b.num_relative_primal_infeasibility;
  a.num_relative_dual_infeasibility += b.num_relative_dual_infeasibility;
  a.num_primal_residual_error += b.num_primal_residual_error; a.num_dual_residual_error += b.num_dual_residual_error;
  a.num_relative_primal_residual_error += b.num_relative_primal_residual_error;
  a.num_relative_dual_residual_error += b.num_relative_dual_residual_error;
  a.num_complementarity_violation += b.num_complementarity_violation;
  a.max_primal_infeasibility = kmax(a.max_primal_infeasibility, b.max_primal_infeasibility);
  a.max_dual_infeasibility = kmax(a.max_dual_infeasibility, b.max_dual_infeasibility);
  a.max_relative_primal_infeasibility = kmax(a.max_relative_primal_infeasibility, b.max_relative_primal_infeasibility);
  a.max_relative_dual_infeasibility = kmax(a.max_relative_dual_infeasibility, b.max_relative_dual_infeasibility);
  a.max_primal_residual_error = kmax(a.max_primal_residual_error, b.max_primal_residual_error);
  a.max_dual_residual_error = kmax(a.max_dual_residual_error, b.max_dual_residual_error);
  a.max_relative_primal_residual_error = kmax(a.max_relative_primal_residual_error, b.max_relative_primal_residual_error);
  a.max_relative_dual_residual_error = kmax(a.max_relative_dual_residual_error, b.max_relative_dual_residual_error);
  a.max_complementarity_violation = kmax(a.max_complementarity_violation, b.max_complementarity_violation);
  a.sum_primal_infeasibility += b.sum_primal_infeasibility; a.sum_dual_infeasibility += b.sum_dual_infeasibility;
  a.dual_objective += b.dual_objective; a.objective += b.objective;
}

}  // namespace b200
