"""BASELINE config 1 through the real drop-in: the reference's own Highs::run() -> solveLp() with
highs_b200/csrc/highs_shim.cpp linked in place of pdlp/CupdlpWrapper.cpp (oracle/_ref/ref_driver_b200,
built by `python oracle/build_ref.py --shim`).  Everything around the boundary -- option handling,
HighsLpSolverObject, lpKktCheck, HighsInfo -- is the reference's unmodified code."""
import numpy as np
import pytest

from conftest import case_id, golden_lp, golden_solution, load_golden

pytestmark = pytest.mark.gpu
CASES = [c for c in load_golden() if c["warm_from"] is None and "synthetic" not in c
         and c["name"] in ("avgas", "afiro", "distillation", "boxed_row", "infeasible", "unbounded", "threed", "restart_lp")]


@pytest.mark.parametrize("case", CASES, ids=case_id)
def test_highs_run_with_b200_shim(oracle, case):
    if not oracle.dropin_available():
        pytest.skip("oracle/_ref/ref_driver_b200 not built")
    lp = golden_lp(case)
    res = oracle.run_reference(lp=lp, options=case["options"], want_solution=True, driver=oracle.DROPIN_DRIVER)
    assert res["model_status_code"] == case["model_status_code"], (res["model_status"], case["model_status"])
    assert res["pdlp_iteration_count"] == case["pdlp_iteration_count"]
    if case["model_status_code"] in (7, 14):
        assert res["objective_function_value"] == pytest.approx(case["objective_function_value"], rel=1e-12, abs=1e-12)
        for fld in ("max_primal_infeasibility", "max_dual_infeasibility", "primal_dual_objective_error",
                    "max_complementarity_violation"):
            assert res[fld] == pytest.approx(case[fld], rel=1e-9, abs=1e-15), fld
        gold = golden_solution(case)
        for k in ("col_value", "col_dual", "row_value", "row_dual"):
            assert np.array_equal(res[k], gold[k]), k
