"""HiGHS's post-solve KKT assessment on the GPU (b200pdlp_kkt_check, highs_b200/csrc/kkt_check.cu)
against its host twin (checked on the CPU against the reference's lpKktCheck: tests/test_kkt_host.py) and against the
reference itself (oracle/_ref/ref_driver --kkt-of), on solutions produced by the engine."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from test_kkt_host import FLOAT_FIELDS, INT_FIELDS, compare

pytestmark = pytest.mark.gpu


def _same(dev, host):
    for k in FLOAT_FIELDS + ("norm_bounds", "norm_costs", "dual_objective_value"):
        tol = 1e-9 * (1 + abs(host[k])) if "residual" in k else 1e-12 * (1 + abs(host[k]))
        assert abs(dev[k] - host[k]) <= tol, (k, dev[k], host[k])
    for k in INT_FIELDS + ("num_relative_primal_infeasibilities", "num_relative_dual_infeasibilities", "model_status"):
        assert dev[k] == host[k], (k, dev[k], host[k])


@pytest.mark.parametrize("name", ["afiro", "adlittle", "e226", "standata", "stair", "scrs8", "25fv47", "80bau3b"])
def test_device_kkt_matches_host_twin_and_reference(engine_lib, oracle, name):
    from highs_b200 import engine
    from highs_b200.lp import read_b2lp
    lp = read_b2lp(os.path.join(GOLDEN, "instances", name + ".b2lp"))
    sol = engine.solve(lp, tol_primal=1e-5, tol_dual=1e-5, tol_gap=1e-5, iter_limit=20000)
    status = 7 if sol["term_code"] == 0 else 14
    dev = engine.kkt_check(lp, sol, kkt_tolerance=1e-5, model_status=status)
    host = engine.kkt_check(lp, sol, kkt_tolerance=1e-5, model_status=status, on_device=False)
    _same(dev, host)
    if oracle.ref_available():
        compare(dev, oracle.reference_kkt(lp, sol, model_status_code=status, options={"kkt_tolerance": 1e-5}))


def test_device_kkt_large(engine_lib, oracle):
    """config S2's size (100k x 100k, 1M nonzeros) with a dense column: the engine's converged solution"""
    from highs_b200 import engine
    from highs_b200.lp import synthetic_lp
    lp = synthetic_lp(100_000, 100_000, 10, seed=12345, dense_col_nnz=20_000)
    sol = engine.solve(lp, tol_primal=1e-4, tol_dual=1e-4, tol_gap=1e-4, iter_limit=200000)
    assert sol["term_code"] == 0
    dev = engine.kkt_check(lp, sol, kkt_tolerance=1e-4, model_status=7)
    host = engine.kkt_check(lp, sol, kkt_tolerance=1e-4, model_status=7, on_device=False)
    _same(dev, host)
    assert dev["model_status"] == 7 and dev["max_primal_residual_error"] < 1e-9
    if oracle.ref_available():
        compare(dev, oracle.reference_kkt(lp, sol, model_status_code=7, options={"kkt_tolerance": 1e-4}))


def test_kkt_has_no_cpu_fallback_symbol_split(engine_lib):
    """the device entry point and the host twin are separate symbols: the product path never silently drops to the CPU"""
    import ctypes
    from highs_b200 import engine
    L = engine.lib()
    assert ctypes.cast(L.b200pdlp_kkt_check, ctypes.c_void_p).value != ctypes.cast(L.b200pdlp_kkt_check_host, ctypes.c_void_p).value
