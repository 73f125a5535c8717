"""ctypes binding of libb200pdlp.so (C ABI: include/b200pdlp.h).

Plumbing only: it marshals numpy arrays into the C structs.  There is no CPU
fallback -- if the CUDA extension is missing or no device is visible every compute
entry point raises `EngineError`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .lp import HighsLp

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200pdlp.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

TERM_NAMES = {0: "OPTIMAL", 1: "INFEASIBLE", 2: "UNBOUNDED", 3: "INFEASIBLE_OR_UNBOUNDED",
              4: "TIMELIMIT_OR_ITERLIMIT", 5: "FEASIBLE"}


class EngineError(RuntimeError):
    pass


class CLp(C.Structure):
    _fields_ = [("num_col", C.c_int32), ("num_row", C.c_int32), ("a_start", _ip), ("a_index", _ip),
                ("a_value", _dp), ("col_cost", _dp), ("col_lower", _dp), ("col_upper", _dp),
                ("row_lower", _dp), ("row_upper", _dp), ("sense", C.c_double), ("offset", C.c_double)]


class CParams(C.Structure):
    _fields_ = [("iter_limit", C.c_int32), ("tol_primal", C.c_double), ("tol_dual", C.c_double),
                ("tol_gap", C.c_double), ("time_limit", C.c_double), ("scaling", C.c_int32),
                ("adaptive_step", C.c_int32), ("restart", C.c_int32), ("log_level", C.c_int32),
                ("check_interval", C.c_int32), ("device", C.c_int32), ("graph_passes", C.c_int32),
                ("ordered_max", C.c_int32), ("device_scaling", C.c_int32), ("reserved", C.c_int32 * 2)]


class CHipdlpParams(C.Structure):
    _fields_ = [("tolerance", C.c_double), ("iter_limit", C.c_int32), ("scaling_mode", C.c_int32),
                ("ruiz_iterations", C.c_int32), ("step_size_strategy", C.c_int32), ("time_limit", C.c_double),
                ("ordered_max", C.c_int32), ("device", C.c_int32), ("log_level", C.c_int32), ("reserved", C.c_int32 * 3)]


class CWarm(C.Structure):
    _fields_ = [("col_value", _dp), ("row_value", _dp), ("row_dual", _dp)]


class CResult(C.Structure):
    _fields_ = [("col_value", _dp), ("col_dual", _dp), ("row_value", _dp), ("row_dual", _dp),
                ("value_valid", C.c_int32), ("dual_valid", C.c_int32), ("term_code", C.c_int32),
                ("term_iterate", C.c_int32), ("iters", C.c_int32), ("passes", C.c_int32),
                ("restarts", C.c_int32), ("kernel_launches", C.c_int32),
                ("primal_obj", C.c_double), ("dual_obj", C.c_double), ("primal_feas", C.c_double),
                ("dual_feas", C.c_double), ("gap", C.c_double), ("rel_gap", C.c_double),
                ("setup_seconds", C.c_double), ("solve_seconds", C.c_double), ("iter_device_ms", C.c_double), ("loop_device_ms", C.c_double),
                ("form_cols", C.c_int32), ("form_rows", C.c_int32), ("form_nnz", C.c_int32),
                ("form_neq", C.c_int32), ("trace", _dp), ("trace_cap", C.c_int32), ("trace_len", C.c_int32)]


# every symbol include/b200pdlp.h declares (tests/test_abi.py checks the library exports them all)
ABI_SYMBOLS = [
    "b200pdlp_default_params", "b200pdlp_solve", "b200pdlp_problem_create", "b200pdlp_problem_destroy",
    "b200pdlp_problem_dims", "b200pdlp_problem_get_vector", "b200pdlp_problem_get_csr", "b200pdlp_spmv_ax",
    "b200pdlp_spmv_aty", "b200pdlp_bench_spmv", "b200pdlp_bench_pass", "b200pdlp_p2p_export", "b200pdlp_p2p_import", "b200pdlp_p2p_timeline", "b200pdlp_p2p_release", "b200pdlp_problem_solve", "b200pdlp_nccl_unique_id",
    "b200pdlp_comm_init", "b200pdlp_partition_rows", "b200pdlp_last_error", "b200pdlp_version",
    "b200pdlp_device_count", "b200pdlp_form_create", "b200pdlp_form_destroy", "b200pdlp_form_dims",
    "b200pdlp_form_get_vector", "b200pdlp_form_get_csc", "b200pdlp_form_get_csr", "b200pdlp_form_get_row_map",
    "b200pdlp_form_layout_eval", "b200pdlp_p2p_link_local", "b200pdlp_solve_multi",
    "b200pdlp_hipdlp_form_create", "b200pdlp_hipdlp_power_method", "b200pdlp_hipdlp_default_params",
    "b200pdlp_solve_hipdlp", "b200pdlp_hipdlp_controller_replay", "b200pdlp_debug_prep_compare", "b200pdlp_release_cache",
    "b200pdlp_host_register", "b200pdlp_host_unregister", "b200pdlp_primal_step", "b200pdlp_dual_step", "b200pdlp_residuals",
    "b200pdlp_kkt_check", "b200pdlp_kkt_check_host", "b200pdlp_kkt_default_tolerances",
]

_lib = None


def _preload_nccl():
    """The engine binds NCCL at run time by SONAME (dlopen "libnccl.so.2").  If PyTorch's bundled,
    newer NCCL is installed, load THAT copy first so that the engine and a later `import torch` share
    one NCCL (the system libnccl lacks symbols torch needs)."""
    import importlib.util
    import sys
    if "torch" in sys.modules:
        return
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
    except Exception:
        pass


def lib():
    """Load the CUDA extension; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"{LIB_PATH} is missing: run `python -m highs_b200.build` (nvcc, sm_100a). "
                              "There is no CPU fallback.")
        _preload_nccl()
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L.b200pdlp_last_error.restype = C.c_char_p
        L.b200pdlp_default_params.argtypes = [C.POINTER(CParams)]
        L.b200pdlp_default_params.restype = None
        L.b200pdlp_solve.argtypes = [C.POINTER(CLp), C.POINTER(CParams), C.POINTER(CWarm), C.POINTER(CResult)]
        L.b200pdlp_problem_create.argtypes = [C.POINTER(CLp), C.POINTER(CParams), C.c_int32, C.c_int32,
                                              C.POINTER(C.c_void_p)]
        L.b200pdlp_problem_destroy.argtypes = [C.c_void_p]
        L.b200pdlp_problem_destroy.restype = None
        L.b200pdlp_problem_dims.argtypes = [C.c_void_p, _ip]
        L.b200pdlp_problem_get_vector.argtypes = [C.c_void_p, C.c_int32, _dp, C.c_int32]
        L.b200pdlp_problem_get_csr.argtypes = [C.c_void_p, _ip, _ip, _dp]
        L.b200pdlp_spmv_ax.argtypes = [C.c_void_p, _dp, _dp]
        L.b200pdlp_spmv_aty.argtypes = [C.c_void_p, _dp, _dp]
        L.b200pdlp_bench_spmv.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_float)]
        L.b200pdlp_bench_pass.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_float)]
        L.b200pdlp_problem_solve.argtypes = [C.c_void_p, C.POINTER(CParams), C.POINTER(CWarm), C.POINTER(CResult)]
        L.b200pdlp_nccl_unique_id.argtypes = [C.POINTER(C.c_uint8)]
        L.b200pdlp_comm_init.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
        L.b200pdlp_p2p_export.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
        L.b200pdlp_p2p_import.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
        L.b200pdlp_p2p_timeline.argtypes = [C.c_void_p, _dp]
        L.b200pdlp_p2p_release.argtypes = [C.c_void_p]
        L.b200pdlp_p2p_link_local.argtypes = [C.POINTER(C.c_void_p), C.c_int32]
        L.b200pdlp_partition_rows.argtypes = [C.POINTER(CLp), C.c_int32, _ip]
        L.b200pdlp_form_create.argtypes = [C.POINTER(CLp), C.c_int32, C.POINTER(C.c_void_p)]
        L.b200pdlp_form_destroy.argtypes = [C.c_void_p]
        L.b200pdlp_form_destroy.restype = None
        L.b200pdlp_form_dims.argtypes = [C.c_void_p, _ip, _dp]
        L.b200pdlp_form_get_vector.argtypes = [C.c_void_p, C.c_int32, _dp, C.c_int32]
        L.b200pdlp_form_get_csc.argtypes = [C.c_void_p, _ip, _ip, _dp]
        L.b200pdlp_form_get_row_map.argtypes = [C.c_void_p, _ip, _ip]
        L.b200pdlp_form_get_csr.argtypes = [C.c_void_p, _ip, _ip, _dp]
        L.b200pdlp_form_layout_eval.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp, _dp]
        L.b200pdlp_debug_prep_compare.argtypes = [C.POINTER(CLp), C.c_int32, _dp]
        L.b200pdlp_release_cache.argtypes = []
        L.b200pdlp_release_cache.restype = None
        L.b200pdlp_host_register.argtypes = [C.c_void_p, C.c_size_t]
        L.b200pdlp_host_unregister.argtypes = [C.c_void_p]
        L.b200pdlp_primal_step.argtypes = [C.c_void_p, _dp, _dp, C.c_double, _dp, _dp]
        L.b200pdlp_dual_step.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double, _dp, _dp, _dp]
        L.b200pdlp_residuals.argtypes = [C.c_void_p, _dp, _dp, _dp]
        _lib = L
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        raise EngineError(f"{what} failed ({rc}): {lib().b200pdlp_last_error().decode()}")


def _p(a, t):
    return a.ctypes.data_as(t)


def make_clp(lp: HighsLp):
    a = lp.a_matrix_
    keep = (a.start_, a.index_, a.value_, lp.col_cost_, lp.col_lower_, lp.col_upper_, lp.row_lower_, lp.row_upper_)
    c = CLp(lp.num_col_, lp.num_row_, _p(a.start_, _ip), _p(a.index_, _ip), _p(a.value_, _dp), _p(lp.col_cost_, _dp),
            _p(lp.col_lower_, _dp), _p(lp.col_upper_, _dp), _p(lp.row_lower_, _dp), _p(lp.row_upper_, _dp),
            float(lp.sense_), float(lp.offset_))
    return c, keep


def make_params(**kw) -> CParams:
    p = CParams()
    lib().b200pdlp_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown parameter {k}")
        setattr(p, k, v)
    return p


def _result_dict(res: CResult, arrays) -> dict:
    cv, cd, rv, rd = arrays[:4]
    skip = ("col_value", "col_dual", "row_value", "row_dual", "trace")
    out = {k: getattr(res, k) for k, _ in CResult._fields_ if k not in skip}
    out.update(col_value=cv, col_dual=cd, row_value=rv, row_dual=rd, term_name=TERM_NAMES.get(res.term_code, "?"))
    if len(arrays) > 4:
        out["trace"] = arrays[4][: res.trace_len].copy()
    return out


def _mk_result(lp: HighsLp, trace_cap: int = 0):
    n, m = lp.num_col_, lp.num_row_
    arrays = (np.zeros(n), np.zeros(n), np.zeros(m), np.zeros(m))
    res = CResult()
    res.col_value, res.col_dual, res.row_value, res.row_dual = (_p(a, _dp) for a in arrays)
    if trace_cap > 0:
        tr = np.zeros((trace_cap, 16))
        res.trace, res.trace_cap = _p(tr, _dp), trace_cap
        arrays = arrays + (tr,)
    return res, arrays


def _mk_warm(warm):
    if warm is None:
        return None, None
    arrs = tuple(np.ascontiguousarray(w, dtype=np.float64) for w in warm)  # (col_value, row_value, row_dual)
    return CWarm(_p(arrs[0], _dp), _p(arrs[1], _dp), _p(arrs[2], _dp)), arrs


def solve(lp: HighsLp, warm=None, trace_cap: int = 0, out_arrays=None, **params) -> dict:
    """b200pdlp_solve: host buffers in, host buffers out (prologue + upload + PDHG + download).
    out_arrays = (col_value, col_dual, row_value, row_dual): caller-owned (e.g. page-locked) result storage."""
    L = lib()
    clp, keep = make_clp(lp)
    prm = make_params(**params)
    res, arrays = _mk_result(lp, trace_cap)
    if out_arrays is not None:
        arrays = tuple(out_arrays) + arrays[4:]
        res.col_value, res.col_dual, res.row_value, res.row_dual = (_p(a, _dp) for a in arrays[:4])
    w, wk = _mk_warm(warm)
    _check(L.b200pdlp_solve(C.byref(clp), C.byref(prm), C.byref(w) if w else None, C.byref(res)), "b200pdlp_solve")
    return _result_dict(res, arrays)


PREP_ITEMS = ["n", "m", "nnz", "neq", "cbeg", "cidx", "cval", "cost", "lower", "upper", "col_scale", "rhs", "row_scale", "rptr",
              "rpos", "row_new_idx", "row_class", "rperm", "cperm", "A.slices", "A.col", "A.val", "AT.slices", "AT.col", "AT.val",
              "A.long", "AT.long", "amax", "norm_cost_relerr", "norm_rhs_relerr", "device_order_vectors", "cols_sorted"]


def prep_compare(lp: HighsLp, scaling: int = 1) -> dict:
    """b200pdlp_debug_prep_compare: the device-resident prologue against its host twin, mismatches per array."""
    clp, keep = make_clp(lp)
    rep = np.zeros(32)
    _check(lib().b200pdlp_debug_prep_compare(C.byref(clp), scaling, _p(rep, _dp)), "b200pdlp_debug_prep_compare")
    return dict(zip(PREP_ITEMS, rep.tolist()))


def lp_arrays(lp: HighsLp):
    a = lp.a_matrix_
    return [a.start_, a.index_, a.value_, lp.col_cost_, lp.col_lower_, lp.col_upper_, lp.row_lower_, lp.row_upper_]


def pin_arrays(arrays) -> list:
    """cudaHostRegister the given numpy arrays (b200pdlp_host_register); returns the ones that were pinned."""
    done = []
    for a in arrays:
        if a.nbytes and lib().b200pdlp_host_register(a.ctypes.data_as(C.c_void_p), a.nbytes) == 0:
            done.append(a)
    return done


def unpin_arrays(arrays):
    for a in arrays:
        lib().b200pdlp_host_unregister(a.ctypes.data_as(C.c_void_p))


class CKktTolerances(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("primal_feasibility_tolerance", "dual_feasibility_tolerance", "primal_residual_tolerance",
                                          "dual_residual_tolerance", "optimality_tolerance")]


class CKktInfo(C.Structure):
    _fields_ = ([(k, C.c_double) for k in (
        "objective_function_value", "dual_objective_value", "primal_dual_objective_error", "max_primal_infeasibility",
        "sum_primal_infeasibilities", "max_dual_infeasibility", "sum_dual_infeasibilities", "max_relative_primal_infeasibility",
        "max_relative_dual_infeasibility", "max_primal_residual_error", "max_dual_residual_error",
        "max_relative_primal_residual_error", "max_relative_dual_residual_error", "max_complementarity_violation",
        "norm_bounds", "norm_costs")] + [(k, C.c_int32) for k in (
            "num_primal_infeasibilities", "num_dual_infeasibilities", "num_relative_primal_infeasibilities",
            "num_relative_dual_infeasibilities", "num_primal_residual_errors", "num_dual_residual_errors",
            "num_relative_primal_residual_errors", "num_relative_dual_residual_errors", "num_complementarity_violations",
            "primal_solution_status", "dual_solution_status", "model_status")])


def kkt_check(lp: HighsLp, solution, kkt_tolerance: float = 0.0, model_status: int = 7, on_device: bool = True) -> dict:
    """lpKktCheck of a HighsSolution (b200pdlp_kkt_check on the GPU, b200pdlp_kkt_check_host = its host twin).
    `solution`: mapping with col_value, col_dual, row_value, row_dual; model_status: HighsModelStatus code from the solver."""
    L = lib()
    clp, keep = make_clp(lp)
    tol = CKktTolerances()
    L.b200pdlp_kkt_default_tolerances.argtypes = [C.POINTER(CKktTolerances), C.c_double]
    L.b200pdlp_kkt_default_tolerances.restype = None
    L.b200pdlp_kkt_default_tolerances(C.byref(tol), float(kkt_tolerance))
    info = CKktInfo()
    info.model_status = int(model_status)
    arrs = [np.ascontiguousarray(solution[k], dtype=np.float64) for k in ("col_value", "col_dual", "row_value", "row_dual")]
    fn = L.b200pdlp_kkt_check if on_device else L.b200pdlp_kkt_check_host
    fn.argtypes = [C.POINTER(CLp), _dp, _dp, _dp, _dp, C.POINTER(CKktTolerances), C.POINTER(CKktInfo)]
    rc = fn(C.byref(clp), *(_p(a, _dp) for a in arrs), C.byref(tol), C.byref(info))
    if rc != 0:
        raise EngineError(f"b200pdlp_kkt_check{'_host' if not on_device else ''} failed ({rc})")
    return {k: getattr(info, k) for k, _ in CKktInfo._fields_}


def hipdlp_controller_replay(norm_cost, norm_rhs, op_norm_sq, tolerance, strategy, sums, restart_sums) -> np.ndarray:
    """Host-only: replay the HiPDLP host control over recorded per-block sums (b200pdlp_hipdlp_controller_replay)."""
    L = lib()
    L.b200pdlp_hipdlp_controller_replay.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32, _dp, _dp, _dp]
    sums = np.ascontiguousarray(sums, dtype=np.float64).reshape(-1, 9)
    rs = np.ascontiguousarray(restart_sums, dtype=np.float64).reshape(-1, 3)
    nb = sums.shape[0]
    out = np.zeros((max(nb, 1), 8))
    _check(L.b200pdlp_hipdlp_controller_replay(norm_cost, norm_rhs, op_norm_sq, tolerance, strategy, nb,
                                               _p(sums if nb else np.zeros(9), _dp), _p(rs if nb else np.zeros(3), _dp), _p(out, _dp)),
           "b200pdlp_hipdlp_controller_replay")
    return out[:nb]


def solve_hipdlp(lp: HighsLp, **params) -> dict:
    """b200pdlp_solve_hipdlp: the HiPDLP mode (reflected Halpern PDHG, solver=hipdlp) on one GPU; parameters are the
    HighsOptions of the same meaning (tolerance, iter_limit, scaling_mode, ruiz_iterations, step_size_strategy ...)."""
    L = lib()
    L.b200pdlp_hipdlp_default_params.argtypes = [C.POINTER(CHipdlpParams)]
    L.b200pdlp_hipdlp_default_params.restype = None
    L.b200pdlp_solve_hipdlp.argtypes = [C.POINTER(CLp), C.POINTER(CHipdlpParams), C.POINTER(CResult)]
    clp, keep = make_clp(lp)
    prm = CHipdlpParams()
    L.b200pdlp_hipdlp_default_params(C.byref(prm))
    for k, v in params.items():
        if not hasattr(prm, k):
            raise TypeError(f"unknown parameter {k}")
        setattr(prm, k, v)
    res, arrays = _mk_result(lp, 0)
    _check(L.b200pdlp_solve_hipdlp(C.byref(clp), C.byref(prm), C.byref(res)), "b200pdlp_solve_hipdlp")
    return _result_dict(res, arrays)


def solve_multi(lp: HighsLp, ngpus: int, devices=None, warm=None, **params) -> dict:
    """b200pdlp_solve_multi: the whole solve on `ngpus` devices of this process (in-process peer linking, no NCCL)."""
    L = lib()
    L.b200pdlp_solve_multi.argtypes = [C.POINTER(CLp), C.POINTER(CParams), C.POINTER(CWarm), C.POINTER(CResult), C.c_int32, _ip]
    clp, keep = make_clp(lp)
    prm = make_params(**params)
    res, arrays = _mk_result(lp, 0)
    w, wk = _mk_warm(warm)
    dev = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
    _check(L.b200pdlp_solve_multi(C.byref(clp), C.byref(prm), C.byref(w) if w else None, C.byref(res), ngpus,
                                  _p(dev, _ip) if dev is not None else None), "b200pdlp_solve_multi")
    return _result_dict(res, arrays)


class Problem:
    """Persistent device-resident problem (b200pdlp_problem_*)."""

    def __init__(self, lp: HighsLp, rank: int = 0, world: int = 1, **params):
        L = lib()
        self.lp = lp
        self._clp, self._keep = make_clp(lp)
        self._prm = make_params(**params)
        self._h = C.c_void_p()
        _check(L.b200pdlp_problem_create(C.byref(self._clp), C.byref(self._prm), rank, world, C.byref(self._h)),
               "b200pdlp_problem_create")
        d = (C.c_int32 * 8)()
        _check(L.b200pdlp_problem_dims(self._h, d), "b200pdlp_problem_dims")
        (self.n, self.m, self.nnz, self.neq, self.m_local, self.row_offset, self.nnz_local, self.n_orig) = list(d)

    def close(self):
        if self._h:
            lib().b200pdlp_problem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def vector(self, which: str) -> np.ndarray:
        idx = {"cost": 0, "lower": 1, "upper": 2, "rhs": 3, "col_scale": 4, "row_scale": 5}[which]
        cap = self.n if idx in (0, 1, 2, 4) else self.m
        out = np.zeros(cap)
        k = lib().b200pdlp_problem_get_vector(self._h, idx, _p(out, _dp), cap)
        if k < 0:
            raise EngineError("b200pdlp_problem_get_vector failed")
        return out[:k]

    def csr(self):
        rp = np.zeros(self.m_local + 1, dtype=np.int32)
        col = np.zeros(max(self.nnz_local, 1), dtype=np.int32)
        val = np.zeros(max(self.nnz_local, 1))
        _check(lib().b200pdlp_problem_get_csr(self._h, _p(rp, _ip), _p(col, _ip), _p(val, _dp)), "get_csr")
        return rp, col[: self.nnz_local], val[: self.nnz_local]

    def spmv_ax(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.shape == (self.n,)
        out = np.zeros(self.m_local)
        _check(lib().b200pdlp_spmv_ax(self._h, _p(x, _dp), _p(out, _dp)), "b200pdlp_spmv_ax")
        return out

    def spmv_aty(self, y: np.ndarray) -> np.ndarray:
        y = np.ascontiguousarray(y, dtype=np.float64)
        assert y.shape == (self.m_local,)
        out = np.zeros(self.n)
        _check(lib().b200pdlp_spmv_aty(self._h, _p(y, _dp), _p(out, _dp)), "b200pdlp_spmv_aty")
        return out

    def primal_step(self, x, aty, tau):
        """K1 on host vectors (standard-form order): returns (x_new, |x - x_new|^2)"""
        x, aty = (np.ascontiguousarray(v, dtype=np.float64) for v in (x, aty))
        out, d = np.zeros(self.n), C.c_double()
        _check(lib().b200pdlp_primal_step(self._h, _p(x, _dp), _p(aty, _dp), float(tau), _p(out, _dp), C.byref(d)), "b200pdlp_primal_step")
        return out, d.value

    def dual_step(self, x_new, y, ax, sigma):
        """K2 on host vectors: returns (y_new, ax_new = A x_new, |y - y_new|^2)"""
        x_new, y, ax = (np.ascontiguousarray(v, dtype=np.float64) for v in (x_new, y, ax))
        yn, axn, d = np.zeros(self.m), np.zeros(self.m), C.c_double()
        _check(lib().b200pdlp_dual_step(self._h, _p(x_new, _dp), _p(y, _dp), _p(ax, _dp), float(sigma), _p(yn, _dp), _p(axn, _dp),
                                        C.byref(d)), "b200pdlp_dual_step")
        return yn, axn, d.value

    def residuals(self, x, y) -> dict:
        x, y = (np.ascontiguousarray(v, dtype=np.float64) for v in (x, y))
        out = np.zeros(10)
        _check(lib().b200pdlp_residuals(self._h, _p(x, _dp), _p(y, _dp), _p(out, _dp)), "b200pdlp_residuals")
        return dict(zip(("pobj", "dobj", "pfeas", "dfeas", "gap", "relgap", "pinf_obj", "pinf_res", "dinf_obj", "dinf_res"), out.tolist()))

    def bench_spmv(self, which: int, reps: int) -> float:
        ms = C.c_float()
        _check(lib().b200pdlp_bench_spmv(self._h, which, reps, C.byref(ms)), "b200pdlp_bench_spmv")
        return float(ms.value)

    def bench_pass(self, reps: int):
        ms = (C.c_float * 4)()
        _check(lib().b200pdlp_bench_pass(self._h, reps, ms), "b200pdlp_bench_pass")
        return [float(v) for v in ms]

    def comm_init(self, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(lib().b200pdlp_comm_init(self._h, buf), "b200pdlp_comm_init")

    def p2p_timeline(self) -> dict:
        out = np.zeros(8)
        _check(lib().b200pdlp_p2p_timeline(self._h, _p(out, _dp)), "b200pdlp_p2p_timeline")
        keys = ["primal_shard_phase", "barrier0_total", "barrier0_wait", "ax_aty_phase", "barrier1_total", "barrier1_wait", "passes"]
        return dict(zip(keys, out[:7].tolist()))

    def p2p_release(self):
        _check(lib().b200pdlp_p2p_release(self._h), "b200pdlp_p2p_release")

    def p2p_export(self) -> bytes:
        buf = (C.c_uint8 * 256)()
        _check(lib().b200pdlp_p2p_export(self._h, buf), "b200pdlp_p2p_export")
        return bytes(buf)

    def p2p_import(self, all_handles: bytes):
        buf = (C.c_uint8 * len(all_handles)).from_buffer_copy(all_handles)
        _check(lib().b200pdlp_p2p_import(self._h, buf), "b200pdlp_p2p_import")

    def solve(self, warm=None, trace_cap: int = 0, **params) -> dict:
        prm = make_params(**params)
        res, arrays = _mk_result(self.lp, trace_cap)
        w, wk = _mk_warm(warm)
        _check(lib().b200pdlp_problem_solve(self._h, C.byref(prm), C.byref(w) if w else None, C.byref(res)),
               "b200pdlp_problem_solve")
        return _result_dict(res, arrays)


def nccl_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    _check(lib().b200pdlp_nccl_unique_id(buf), "b200pdlp_nccl_unique_id")
    return bytes(buf)


def partition_rows(lp: HighsLp, world: int) -> np.ndarray:
    clp, keep = make_clp(lp)
    out = np.zeros(world + 1, dtype=np.int32)
    _check(lib().b200pdlp_partition_rows(C.byref(clp), world, _p(out, _ip)), "b200pdlp_partition_rows")
    return out


def host_form(lp: HighsLp, scaling: int = 1) -> dict:
    """Host-only standard form (formulate + scale) as numpy arrays; needs no GPU."""
    L = lib()
    clp, keep = make_clp(lp)
    h = C.c_void_p()
    _check(L.b200pdlp_form_create(C.byref(clp), scaling, C.byref(h)), "b200pdlp_form_create")
    try:
        d = (C.c_int32 * 5)()
        sc = (C.c_double * 3)()
        _check(L.b200pdlp_form_dims(h, d, sc), "b200pdlp_form_dims")
        n, m, nnz, neq, n_orig = list(d)
        out = dict(n=n, m=m, nnz=nnz, neq=neq, n_orig=n_orig, norm_cost=sc[0], norm_rhs=sc[1], amax=sc[2])
        for idx, (name, ln) in enumerate([("cost", n), ("lower", n), ("upper", n), ("rhs", m), ("col_scale", n), ("row_scale", m)]):
            v = np.zeros(max(ln, 1))
            k = L.b200pdlp_form_get_vector(h, idx, _p(v, _dp), ln)
            out[name] = v[:k].copy()
        cbeg = np.zeros(n + 1, dtype=np.int32)
        cidx = np.zeros(max(nnz, 1), dtype=np.int32)
        cval = np.zeros(max(nnz, 1))
        _check(L.b200pdlp_form_get_csc(h, _p(cbeg, _ip), _p(cidx, _ip), _p(cval, _dp)), "form_get_csc")
        rni = np.zeros(max(m, 1), dtype=np.int32)
        rcl = np.zeros(max(m, 1), dtype=np.int32)
        _check(L.b200pdlp_form_get_row_map(h, _p(rni, _ip), _p(rcl, _ip)), "form_get_row_map")
        rbeg = np.zeros(m + 1, dtype=np.int32)
        ridx = np.zeros(max(nnz, 1), dtype=np.int32)
        rval = np.zeros(max(nnz, 1))
        _check(L.b200pdlp_form_get_csr(h, _p(rbeg, _ip), _p(ridx, _ip), _p(rval, _dp)), "form_get_csr")
        out.update(cbeg=cbeg, cidx=cidx[:nnz], cval=cval[:nnz], row_new_idx=rni[:m], row_type=rcl[:m],
                   rbeg=rbeg, ridx=ridx[:nnz], rval=rval[:nnz])
        return out
    finally:
        L.b200pdlp_form_destroy(h)


def host_form_hipdlp(lp: HighsLp, scaling_mode: int = 5, ruiz_iterations: int = 10) -> dict:
    """Host-only HiPDLP standard form (preprocess + scale) and its power-method estimate; needs no GPU."""
    L = lib()
    L.b200pdlp_hipdlp_form_create.argtypes = [C.POINTER(CLp), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    L.b200pdlp_hipdlp_power_method.argtypes = [C.c_void_p, _dp]
    clp, keep = make_clp(lp)
    h = C.c_void_p()
    _check(L.b200pdlp_hipdlp_form_create(C.byref(clp), scaling_mode, ruiz_iterations, C.byref(h)), "b200pdlp_hipdlp_form_create")
    try:
        d = (C.c_int32 * 5)()
        sc = (C.c_double * 3)()
        _check(L.b200pdlp_form_dims(h, d, sc), "b200pdlp_form_dims")
        n, m, nnz, neq, n_orig = list(d)
        out = dict(n=n, m=m, nnz=nnz, neq=neq, n_orig=n_orig, c_norm=sc[0], rhs_norm=sc[1], amax=sc[2])
        for idx, (name, ln) in enumerate([("cost", n), ("lower", n), ("upper", n), ("rlo", m), ("col_scale", n), ("row_scale", m), ("rup", m)]):
            v = np.zeros(max(ln, 1))
            k = L.b200pdlp_form_get_vector(h, idx, _p(v, _dp), ln)
            out[name] = v[:k].copy()
        cbeg = np.zeros(n + 1, dtype=np.int32)
        cidx = np.zeros(max(nnz, 1), dtype=np.int32)
        cval = np.zeros(max(nnz, 1))
        _check(L.b200pdlp_form_get_csc(h, _p(cbeg, _ip), _p(cidx, _ip), _p(cval, _dp)), "form_get_csc")
        rni = np.zeros(max(m, 1), dtype=np.int32)
        rcl = np.zeros(max(m, 1), dtype=np.int32)
        _check(L.b200pdlp_form_get_row_map(h, _p(rni, _ip), _p(rcl, _ip)), "form_get_row_map")
        lam = np.zeros(1)
        _check(L.b200pdlp_hipdlp_power_method(h, _p(lam, _dp)), "b200pdlp_hipdlp_power_method")
        out.update(cbeg=cbeg, cidx=cidx[:nnz], cval=cval[:nnz], new_idx=rni[:m], ctype=rcl[:m], op_norm_sq=float(lam[0]))
        return out
    finally:
        L.b200pdlp_form_destroy(h)


def host_layout_eval(lp: HighsLp, world: int = 1, ordered_max: int = 0, x=None, y=None, scaling: int = 1, seed: int = 0) -> dict:
    """Host-only: build every rank's device layout of `lp` (as b200pdlp_problem_create would) and evaluate the
    sliced-ELL data on the host.  Returns ax (assembled over ranks), aty (summed over ranks in rank order), the
    per-rank partial A_g'y, the per-rank stats and the standard form's column-wise matrix for checking."""
    L = lib()
    clp, keep = make_clp(lp)
    h = C.c_void_p()
    _check(L.b200pdlp_form_create(C.byref(clp), scaling, C.byref(h)), "b200pdlp_form_create")
    try:
        d = (C.c_int32 * 5)()
        sc = (C.c_double * 3)()
        _check(L.b200pdlp_form_dims(h, d, sc), "b200pdlp_form_dims")
        n, m, nnz = d[0], d[1], d[2]
        rng = np.random.default_rng(seed)
        x = rng.standard_normal(max(n, 1)) if x is None else np.ascontiguousarray(x, dtype=np.float64)
        y = rng.standard_normal(max(m, 1)) if y is None else np.ascontiguousarray(y, dtype=np.float64)
        ax = np.full(max(m, 1), np.nan)
        aty = np.zeros(max(n, 1))
        parts, stats = [], []
        for g in range(world):
            part = np.full(max(n, 1), np.nan)
            st = np.zeros(12)
            _check(L.b200pdlp_form_layout_eval(h, g, world, ordered_max, _p(x, _dp), _p(y, _dp), _p(ax, _dp), _p(part, _dp),
                                               _p(st, _dp)), "b200pdlp_form_layout_eval")
            parts.append(part[:n].copy())
            stats.append(st)
            aty[:n] += part[:n]
        cbeg = np.zeros(n + 1, dtype=np.int32)
        cidx = np.zeros(max(nnz, 1), dtype=np.int32)
        cval = np.zeros(max(nnz, 1))
        _check(L.b200pdlp_form_get_csc(h, _p(cbeg, _ip), _p(cidx, _ip), _p(cval, _dp)), "form_get_csc")
        return dict(n=n, m=m, nnz=nnz, x=x[:n], y=y[:m], ax=ax[:m], aty=aty[:n], parts=parts, stats=stats,
                    cbeg=cbeg, cidx=cidx[:nnz], cval=cval[:nnz])
    finally:
        L.b200pdlp_form_destroy(h)


def solve_logical_shards(lp: HighsLp, world: int, device: int = -1, **params) -> list:
    """the multi-GPU code path with `world` logical shards inside this process (normally all on one
    device): one problem per rank, buffers wired to each other with b200pdlp_p2p_link_local, one host thread per rank
    (ctypes releases the GIL during the solve).  Returns the per-rank result dicts (identical by construction)."""
    import threading
    probs = [Problem(lp, rank=g, world=world, device=device, **params) for g in range(world)]
    try:
        arr = (C.c_void_p * world)(*[q._h for q in probs])
        _check(lib().b200pdlp_p2p_link_local(arr, world), "b200pdlp_p2p_link_local")
        out, err = [None] * world, [None] * world

        def run(g):
            try:
                out[g] = probs[g].solve(device=device, **params)
            except Exception as e:   # noqa: BLE001 -- reported to the caller below
                err[g] = e

        threads = [threading.Thread(target=run, args=(g,), daemon=True) for g in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in err:
            if e is not None:
                raise e
        return out
    finally:
        for q in probs:
            try:
                q.p2p_release()
            except Exception:
                pass
        for q in probs:
            q.close()


def device_count() -> int:
    return int(lib().b200pdlp_device_count())
