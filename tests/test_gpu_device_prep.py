"""The device-resident prologue (highs_b200/csrc/device_prep.cu: formulateLP_highs, PDHG_Scale_Data, csc2csr, the device
orderings and both sliced-ELL layouts as kernels; /root/reference/highs/pdlp/CupdlpWrapper.cpp:280-448,
cupdlp/cupdlp_scaling.c:233-425, cupdlp_utils.c:1222-1254) against its host twin (host_prep.cpp, itself bit-exact against
the oracle on every instance: tests/test_instances_host.py): every array bit for bit.  Then whole solves through it."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

FILES = sorted(glob.glob(os.path.join(GOLDEN, "instances", "*.b2lp")))
EXACT = ["n", "m", "nnz", "neq", "cbeg", "cidx", "cval", "cost", "lower", "upper", "col_scale", "rhs", "row_scale", "rptr", "rpos",
         "row_new_idx", "row_class", "rperm", "cperm", "A.slices", "A.col", "A.val", "AT.slices", "AT.col", "AT.val", "A.long",
         "AT.long", "amax", "device_order_vectors"]


def _check(rep):
    bad = {k: rep[k] for k in EXACT if rep[k] != 0}
    assert not bad, bad
    assert rep["norm_cost_relerr"] < 1e-13 and rep["norm_rhs_relerr"] < 1e-13, rep


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-5] for f in FILES])
def test_prologue_matches_host_twin_on_instances(engine_lib, path):
    """every LP instance of the reference's own tests: EQ / LEQ / GEQ / ranged / free rows, slack columns, empty rows"""
    from highs_b200 import engine
    from highs_b200.lp import read_b2lp
    _check(engine.prep_compare(read_b2lp(path)))


@pytest.mark.parametrize("scaling", [1, 0])
def test_prologue_matches_host_twin_synthetic(engine_lib, scaling):
    from highs_b200 import engine
    from highs_b200.lp import synthetic_lp
    _check(engine.prep_compare(synthetic_lp(100_000, 80_000, 9, seed=4), scaling))
    _check(engine.prep_compare(synthetic_lp(30_000, 30_000, 6, seed=5, dense_col_nnz=15_000), scaling))   # long rows of A'


def test_prologue_unsorted_columns(engine_lib):
    """columns NOT stored with ascending rows: A' must still add in ascending-row order (second stable sort pass)"""
    from highs_b200 import engine
    from highs_b200.lp import synthetic_lp
    lp = synthetic_lp(20_000, 15_000, 7, seed=8, dense_col_nnz=3000)
    a = lp.a_matrix_
    rng = np.random.default_rng(1)
    for j in range(0, lp.num_col_, 3):   # shuffle the entries of every third column
        b, e = a.start_[j], a.start_[j + 1]
        p = rng.permutation(e - b)
        a.index_[b:e] = a.index_[b:e][p]
        a.value_[b:e] = a.value_[b:e][p]
    rep = engine.prep_compare(lp)
    assert rep["cols_sorted"] == 0
    _check(rep)


def test_prologue_mixed_row_types(engine_lib):
    """a large LP with every row class (EQ, LEQ, GEQ, ranged -> slack columns, free)"""
    from highs_b200 import engine
    from highs_b200.lp import synthetic_lp
    lp = synthetic_lp(40_000, 30_000, 6, seed=12)
    rng = np.random.default_rng(3)
    kind = rng.integers(0, 5, lp.num_row_)
    lo, up = lp.row_lower_.copy(), lp.row_upper_.copy()
    for i in range(lp.num_row_):
        k = kind[i]
        if k == 0:
            up[i] = lo[i]                       # EQ
        elif k == 1:
            up[i], lo[i] = lo[i] + 1.0, -np.inf  # LEQ
        elif k == 2:
            up[i] = lo[i] + 2.5                 # ranged -> BOUND
        elif k == 3:
            lo[i], up[i] = -np.inf, np.inf      # free -> BOUND
    lp.row_lower_, lp.row_upper_ = lo, up
    _check(engine.prep_compare(lp))


def test_solve_through_device_prologue_matches_host_prologue(engine_lib, oracle):
    """the same LP solved through the device prologue and through the host prologue: same optimum (the two differ only in
    the summation order of |c|^2, |b|^2), and both agree with the oracle"""
    from highs_b200 import engine
    from highs_b200.lp import synthetic_lp
    lp = synthetic_lp(12_000, 9_000, 6, seed=21)
    kw = dict(tol_primal=1e-5, tol_dual=1e-5, tol_gap=1e-5, iter_limit=400000)
    dev = engine.solve(lp, **kw)
    host = engine.solve(lp, device_scaling=-1, **kw)
    orc = oracle.solve(lp, **kw)
    assert dev["term_code"] == host["term_code"] == orc["term_code"] == 0
    o_dev, o_host, o_orc = (lp.objectiveValue(r["col_value"]) for r in (dev, host, orc))
    assert abs(o_dev - o_orc) <= 1e-4 * (1 + abs(o_orc)) and abs(o_host - o_orc) <= 1e-4 * (1 + abs(o_orc))   # all three at kkt 1e-5
    assert abs(dev["iters"] - orc["iters"]) <= 0.35 * orc["iters"] + 80   # restarts fall differently once trajectories part
    # a fixed number of iterations: identical up to the propagation of two last-bit differences
    d2 = engine.solve(lp, iter_limit=120)
    h2 = engine.solve(lp, device_scaling=-1, iter_limit=120)
    o2 = oracle.solve(lp, iter_limit=120)
    for k in ("col_value", "row_value", "col_dual", "row_dual"):
        s = 1 + np.abs(o2[k]).max()
        assert np.abs(d2[k] - o2[k]).max() <= 1e-6 * s, k
        assert np.abs(h2[k] - o2[k]).max() <= 1e-6 * s, k


def test_hot_start_through_device_prologue(engine_lib, oracle):
    """PDHG_PreSolve on the device: hot start from a loose solution (permuted, scaled, projected in HBM)"""
    from highs_b200 import engine
    from highs_b200.lp import synthetic_lp
    lp = synthetic_lp(9_000, 7_000, 6, seed=33)
    first = oracle.solve(lp, tol_primal=1e-3, tol_dual=1e-3, tol_gap=1e-3)
    warm = (first["col_value"], first["row_value"], first["row_dual"])
    kw = dict(tol_primal=1e-5, tol_dual=1e-5, tol_gap=1e-5, iter_limit=400000)
    dev = engine.solve(lp, warm=warm, **kw)
    orc = oracle.solve(lp, warm=warm, **kw)
    assert dev["term_code"] == orc["term_code"] == 0
    assert abs(dev["iters"] - orc["iters"]) <= 0.35 * orc["iters"] + 80
    o_dev, o_orc = lp.objectiveValue(dev["col_value"]), lp.objectiveValue(orc["col_value"])
    assert abs(o_dev - o_orc) <= 1e-4 * (1 + abs(o_orc))
    # after ONE iteration from the hot start the two must agree to rounding (no long reductions have acted yet)
    d1 = engine.solve(lp, warm=warm, iter_limit=2)
    o1 = oracle.solve(lp, warm=warm, iter_limit=2)
    for k in ("col_value", "row_dual"):
        assert np.abs(d1[k] - o1[k]).max() <= 1e-9 * (1 + np.abs(o1[k]).max()), k
