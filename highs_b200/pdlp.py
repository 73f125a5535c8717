"""Host-side mirror of the reference's PDLP wrapper interface (the drop-in boundary).

`solveLpCupdlp` below has the argument list and the behaviour of
    HighsStatus solveLpCupdlp(const HighsOptions&, HighsTimer&, const HighsLp&, HighsBasis&,
                              HighsSolution&, HighsModelStatus&, HighsInfo&, HighsCallback&)
(/root/reference/highs/pdlp/CupdlpWrapper.cpp:30-278): it resets status/info, maps
options to solver parameters (getUserParamsFromOptions, :642-717), hot-starts from
an incoming valid HighsSolution (:194-205, cupdlp_solver.c:1217-1279), runs the B200
engine through the C ABI (include/b200pdlp.h), fills HighsSolution / HighsInfo::
pdlp_iteration_count / HighsBasis::valid and maps the termination code to a
HighsModelStatus (:220-251).  The C++ twin that links into HiGHS itself is
highs_b200/csrc/highs_shim.cpp.

There is no CPU fallback here: everything numerical happens in libb200pdlp.so.
"""
from __future__ import annotations

import dataclasses
import enum
import time

import numpy as np

from . import engine
from .lp import HighsLp, kHighsInf

kHighsIInf = 2147483647            # HConst.h
kDefaultKktTolerance = 1e-7        # HConst.h:344-345


class HighsStatus(enum.IntEnum):   # HighsStatus.h
    kError = -1
    kOk = 0
    kWarning = 1


class HighsModelStatus(enum.IntEnum):  # HConst.h:201-228
    kNotset = 0
    kLoadError = 1
    kModelError = 2
    kPresolveError = 3
    kSolveError = 4
    kPostsolveError = 5
    kModelEmpty = 6
    kOptimal = 7
    kInfeasible = 8
    kUnboundedOrInfeasible = 9
    kUnbounded = 10
    kObjectiveBound = 11
    kObjectiveTarget = 12
    kTimeLimit = 13
    kIterationLimit = 14
    kUnknown = 15
    kSolutionLimit = 16
    kInterrupt = 17
    kMemoryLimit = 18


# pdlp_features_off bits, HConst.h:417-422
kPdlpScalingOff = 1
kPdlpRestartOff = 2
kPdlpAdaptiveStepSizeOff = 4


@dataclasses.dataclass
class HighsOptions:
    """The HighsOptions members the wrapper reads (HighsOptions.h:356,403-410,1339-1385)."""
    solver: str = "pdlp"
    presolve: str = "off"
    pdlp_iteration_limit: int = kHighsIInf
    pdlp_features_off: int = 0
    pdlp_cupdlpc_restart_method: int = 1
    primal_feasibility_tolerance: float = 1e-7
    dual_feasibility_tolerance: float = 1e-7
    pdlp_optimality_tolerance: float = 1e-7
    kkt_tolerance: float = kDefaultKktTolerance
    time_limit: float = kHighsInf
    output_flag: bool = False
    log_dev_level: int = 0
    # read by the HiPDLP wrapper only (HighsOptions.h:1345-1379; PDLPSolver::setup, hipdlp/pdhg.cc:1820-1863)
    pdlp_scaling_mode: int = 5            # 1 Ruiz + 4 Pock-Chambolle (+ 2 L2)
    pdlp_ruiz_iterations: int = 10
    pdlp_step_size_strategy: int = 1      # anything but 0 (fixed) means the PID primal weight


class HighsTimer:
    """HighsTimer::read() is all the wrapper uses (CupdlpWrapper.cpp:701-705)."""

    def __init__(self):
        self._t0 = time.monotonic()

    def read(self) -> float:
        return time.monotonic() - self._t0


@dataclasses.dataclass
class HighsSolution:               # HStruct.h:20-35
    value_valid: bool = False
    dual_valid: bool = False
    col_value: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0))
    col_dual: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0))
    row_value: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0))
    row_dual: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0))


@dataclasses.dataclass
class HighsBasis:                  # HStruct.h:60-79 (only `valid` is touched)
    valid: bool = False


@dataclasses.dataclass
class HighsInfo:                   # HighsInfo.h:91-131 (the wrapper writes pdlp_iteration_count only)
    valid: bool = False
    pdlp_iteration_count: int = -1
    # engine diagnostics (not HiGHS fields)
    b200: dict = dataclasses.field(default_factory=dict)


def getCupdlpLogLevel(options: HighsOptions) -> int:   # CupdlpWrapper.cpp:839-848
    if options.output_flag:
        return 2 if options.log_dev_level else 1
    return 0


def getUserParamsFromOptions(options: HighsOptions, timer: HighsTimer) -> dict:
    """CupdlpWrapper.cpp:642-717."""
    p = {}
    p["iter_limit"] = int(min(options.pdlp_iteration_limit, kHighsIInf))
    p["log_level"] = getCupdlpLogLevel(options)
    p["scaling"] = 1 if (options.pdlp_features_off & kPdlpScalingOff) == 0 else 0
    p["adaptive_step"] = 1 if (options.pdlp_features_off & kPdlpAdaptiveStepSizeOff) == 0 else 0
    p["tol_primal"] = options.primal_feasibility_tolerance
    p["tol_dual"] = options.dual_feasibility_tolerance
    p["tol_gap"] = options.pdlp_optimality_tolerance
    if options.kkt_tolerance != kDefaultKktTolerance:
        p["tol_primal"] = p["tol_dual"] = p["tol_gap"] = options.kkt_tolerance
    # the reference computes the remaining time (:701-705) but then passes the FULL limit (:707)
    p["time_limit"] = options.time_limit if options.time_limit < kHighsInf else -1.0
    restart_on = 1 if (options.pdlp_features_off & kPdlpRestartOff) == 0 else 0
    if options.pdlp_cupdlpc_restart_method == 0:
        restart_on = 0
    p["restart"] = restart_on
    return p


def solveLpCupdlp(options: HighsOptions, timer: HighsTimer, lp: HighsLp, highs_basis: HighsBasis,
                  highs_solution: HighsSolution, highs_info: HighsInfo, callback=None,
                  **engine_params):
    """Returns (HighsStatus, HighsModelStatus); writes highs_solution / highs_info / highs_basis."""
    # resetModelStatusAndHighsInfo (:36)
    model_status = HighsModelStatus.kNotset
    highs_info.valid = False
    highs_info.pdlp_iteration_count = -1
    params = getUserParamsFromOptions(options, timer)
    params.update(engine_params)
    warm = None
    if highs_solution.value_valid and highs_solution.dual_valid:     # PDHG_PreSolve condition
        warm = (highs_solution.col_value, highs_solution.row_value, highs_solution.row_dual)
    try:
        res = engine.solve(lp, warm=warm, **params)
    except engine.EngineError:
        highs_solution.value_valid = highs_solution.dual_valid = False
        return HighsStatus.kError, HighsModelStatus.kSolveError
    highs_solution.col_value, highs_solution.col_dual = res["col_value"], res["col_dual"]
    highs_solution.row_value, highs_solution.row_dual = res["row_value"], res["row_dual"]
    highs_info.pdlp_iteration_count = res["iters"]                   # :206
    highs_info.b200 = {k: v for k, v in res.items() if not isinstance(v, np.ndarray)}
    highs_solution.value_valid = bool(res["value_valid"])
    highs_solution.dual_valid = bool(res["dual_valid"])
    highs_basis.valid = False                                         # :223
    code = res["term_code"]
    if code == 0:
        model_status = HighsModelStatus.kOptimal
    elif code == 1:
        model_status = HighsModelStatus.kInfeasible
    elif code == 2:
        model_status = HighsModelStatus.kUnbounded
    elif code == 3:
        model_status = HighsModelStatus.kUnboundedOrInfeasible
    elif code == 4:
        model_status = (HighsModelStatus.kIterationLimit if res["iters"] >= params["iter_limit"] - 1
                        else HighsModelStatus.kTimeLimit)             # :233-236
    else:
        model_status = HighsModelStatus.kUnknown
    return HighsStatus.kOk, model_status


def solveLpHiPdlp(options: HighsOptions, timer: HighsTimer, lp: HighsLp, highs_basis: HighsBasis,
                  highs_solution: HighsSolution, highs_info: HighsInfo, callback=None, **engine_params):
    """Mirror of  HighsStatus solveLpHiPdlp(const HighsOptions&, HighsTimer&, const HighsLp&, HighsBasis&, HighsSolution&,
    HighsModelStatus&, HighsInfo&, HighsCallback&)  (/root/reference/highs/pdlp/HiPdlpWrapper.cpp:26-141, solver=hipdlp):
    options as PDLPSolver::setup reads them (hipdlp/pdhg.cc:1820-1863), the engine's HiPDLP mode through the C ABI
    (b200pdlp_solve_hipdlp), status mapping of :96-127.  Returns (HighsStatus, HighsModelStatus).  The C++ twin is
    highs_b200/csrc/highs_shim_hipdlp.cpp.  No CPU fallback."""
    model_status = HighsModelStatus.kNotset
    highs_info.valid = False
    highs_info.pdlp_iteration_count = -1
    tol = options.pdlp_optimality_tolerance
    if options.kkt_tolerance != kDefaultKktTolerance:
        tol = options.kkt_tolerance
    params = dict(tolerance=tol, iter_limit=int(min(options.pdlp_iteration_limit, kHighsIInf)),
                  time_limit=options.time_limit if options.time_limit < kHighsInf else -1.0,
                  scaling_mode=options.pdlp_scaling_mode if (options.pdlp_features_off & kPdlpScalingOff) == 0 else 0,
                  ruiz_iterations=options.pdlp_ruiz_iterations,
                  step_size_strategy=0 if options.pdlp_step_size_strategy == 0 else 3,
                  log_level=(2 if options.log_dev_level else 1) if options.output_flag else 0)
    params.update(engine_params)
    try:
        res = engine.solve_hipdlp(lp, **params)
    except engine.EngineError:
        return HighsStatus.kError, HighsModelStatus.kUnknown
    highs_solution.col_value, highs_solution.col_dual = res["col_value"], res["col_dual"]
    highs_solution.row_value, highs_solution.row_dual = res["row_value"], res["row_dual"]
    highs_info.pdlp_iteration_count = res["iters"]
    highs_info.b200 = {k: v for k, v in res.items() if not isinstance(v, np.ndarray)}
    highs_solution.value_valid = highs_solution.dual_valid = True
    highs_basis.valid = False
    if res["term_code"] == 0:
        model_status = HighsModelStatus.kOptimal
    elif res["term_iterate"] == 2:
        model_status = HighsModelStatus.kTimeLimit
    else:
        model_status = HighsModelStatus.kIterationLimit
    return HighsStatus.kOk, model_status
