"""params.device_scaling = 1: PDHG_Scale_Data (10 Ruiz passes + Pock-Chambolle) on the GPU
(highs_b200/csrc/setup_kernels.cu) must leave exactly the data the host passes leave -- and therefore exactly the
reference's (tests/test_host_prep.py pins the host passes against the oracle bit for bit).

STATUS: written after this round's GPU budget was spent; compiled, not yet run on hardware, hence
xfail(strict=False): a pass shows up as XPASS, a failure does not turn the suite red.  The marker goes away after the
first hardware run."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="device scaling not yet run on hardware (written after the GPU budget was spent)")]

VECS = ["cost", "lower", "upper", "rhs", "col_scale", "row_scale"]


def _compare(lp, **prm):
    from highs_b200 import engine
    a = engine.Problem(lp, **prm)
    b = engine.Problem(lp, device_scaling=1, **prm)
    try:
        for k in VECS:
            assert np.array_equal(a.vector(k), b.vector(k)), k
        ra, rb = a.csr(), b.csr()
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
        sa = a.solve(trace_cap=64, **prm)
        sb = b.solve(trace_cap=64, **prm)
        assert sa["iters"] == sb["iters"] and sa["term_code"] == sb["term_code"]
        assert np.array_equal(sa["trace"], sb["trace"])          # same amax -> same first step -> same trajectory
        for k in ("col_value", "col_dual", "row_value", "row_dual"):
            assert np.array_equal(sa[k], sb[k]), k
    finally:
        a.close()
        b.close()


@pytest.mark.parametrize("name", ["avgas", "afiro", "adlittle", "sctest", "boxed_row", "e226", "stair"])
def test_device_scaling_golden_lps(name):
    from highs_b200.lp import read_b2lp
    _compare(read_b2lp(os.path.join(GOLDEN, name + ".b2lp")), iter_limit=400)


def test_device_scaling_synthetic_with_dense_column():
    from highs_b200.lp import synthetic_lp
    _compare(synthetic_lp(120000, 90000, 8, 4, dense_col_nnz=30000), iter_limit=200)


def test_device_scaling_ragged():
    # empty rows and columns (norm 0 -> factor 1), explicit zero entries
    from highs_b200.lp import HighsLp, HighsSparseMatrix, kHighsInf
    n, m = 7, 4
    start = np.array([0, 0, 2, 2, 3, 3, 4, 5], dtype=np.int32)
    index = np.array([0, 2, 2, 3, 0], dtype=np.int32)
    value = np.array([1.0, -2.0, 3.0, 0.0, 4.0])
    lp = HighsLp(n, m, np.ones(n), np.zeros(n), np.full(n, kHighsInf), np.array([1.0, -kHighsInf, 0.0, -1.0]),
                 np.array([kHighsInf, 5.0, 2.0, 1.0]), HighsSparseMatrix(n, m, start, index, value), 1, 0.0, "ragged")
    _compare(lp, iter_limit=50)
