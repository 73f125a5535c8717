"""Host-side preparation of the product (formulate + scale, highs_b200/csrc/host_prep.cpp) against
the oracle, bit for bit; row partition properties.  No GPU."""
import numpy as np
import pytest

from conftest import golden_lp, load_golden

KEYS = ["cost", "lower", "upper", "rhs", "col_scale", "row_scale", "cbeg", "cidx", "cval", "row_new_idx", "row_type",
        "rbeg", "ridx", "rval"]
NAMES = sorted({c["name"] for c in load_golden() if "synthetic" not in c})


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("scaling", [1, 0])
def test_standard_form_bit_exact(engine_lib, oracle, name, scaling):
    import os
    from conftest import GOLDEN
    from highs_b200 import engine
    from highs_b200.lp import read_b2lp
    lp = read_b2lp(os.path.join(GOLDEN, name + ".b2lp"))
    a, b = engine.host_form(lp, scaling), oracle.formulate_and_scale(lp, scaling)
    for k in ("n", "m", "nnz", "neq", "n_orig", "norm_cost", "norm_rhs", "amax"):
        assert a[k] == b[k], k
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k


def test_standard_form_synthetic(engine_lib, oracle):
    from highs_b200 import engine
    from highs_b200.lp import synthetic_lp
    lp = synthetic_lp(20000, 15000, 7, 3)
    a, b = engine.host_form(lp, 1), oracle.formulate_and_scale(lp, 1)
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k
    assert a["neq"] == 0 and a["n"] == lp.num_col_    # all rows are GEQ: no slack columns


def test_row_classes_and_slacks(engine_lib):
    """formulateLP_highs semantics (CupdlpWrapper.cpp:280-448) on a hand-made LP"""
    from highs_b200 import engine
    from highs_b200.lp import HighsLp, HighsSparseMatrix
    inf = np.inf
    # rows: GEQ, EQ, BOUND(ranged), LEQ, free
    lp = HighsLp(2, 5, [1, 2], [0, -1e30], [inf, 5], [1, 3, 2, -inf, -1e25], [inf, 3, 10, 5, 1e21],
                 HighsSparseMatrix(2, 5, [0, 5, 10], [0, 1, 2, 3, 4] * 2, [1, 2, 3, 4, 5, 6, 7, 8, 9, 10.0]), -1, 0.0)
    f = engine.host_form(lp, 0)
    assert list(f["row_type"]) == [2, 0, 3, 1, 3]
    assert f["neq"] == 3 and f["n"] == 4 and f["nnz"] == 12
    assert list(f["row_new_idx"]) == [3, 0, 1, 4, 2]          # EQ/BOUND first (original order), then LEQ/GEQ
    assert list(f["rhs"]) == [3, 0, 0, 1, -5]                 # BOUND rhs 0, LEQ negated
    assert list(f["cost"]) == [-1, -2, 0, 0]                  # c * sense, zero slack cost
    assert f["lower"][1] == -inf and list(f["lower"][2:]) == [2, -inf] and list(f["upper"][2:]) == [10, inf]
    # column 0: EQ/BOUND entries first (rows 1,2,4 -> 0,1,2), then GEQ row 0 -> 3 and LEQ row 3 -> 4 negated
    assert list(f["cidx"][:5]) == [0, 1, 2, 3, 4] and list(f["cval"][:5]) == [2, 3, 5, 1, -4]
    assert list(f["cidx"][10:]) == [1, 2] and list(f["cval"][10:]) == [-1, -1]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_rows(engine_lib, world):
    from highs_b200 import engine
    from highs_b200.lp import synthetic_lp
    lp = synthetic_lp(30000, 20000, 6, 1, dense_col_nnz=15000)
    b = engine.partition_rows(lp, world)
    assert b[0] == 0 and b[-1] == lp.num_row_ and np.all(np.diff(b) >= 0)
    rows = np.bincount(lp.a_matrix_.index_, minlength=lp.num_row_)   # all rows GEQ: no permutation
    load = [rows[b[g]:b[g + 1]].sum() + 5 * (b[g + 1] - b[g]) for g in range(world)]
    assert max(load) <= 1.02 * (sum(load) / world) + 64


def test_standard_form_threaded_path(engine_lib, oracle):
    """> 2^18 nonzeros: the multi-threaded scaling sweeps must still be bit-identical to the oracle"""
    from highs_b200 import engine
    from highs_b200.lp import synthetic_lp
    lp = synthetic_lp(120000, 90000, 8, 4, dense_col_nnz=30000)
    a, b = engine.host_form(lp, 1), oracle.formulate_and_scale(lp, 1)
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k
    assert a["amax"] == b["amax"]
