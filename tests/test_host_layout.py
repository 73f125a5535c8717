"""Host evaluation of the device layouts (no GPU): the sliced-ELL data, the length-sorted device orderings, the
long-row segments, the segmented column positions and the A_g' output map that b200pdlp_problem_create uploads
are built by host_prep.cpp::build_layout; b200pdlp_form_layout_eval walks them on the host in the kernels'
traversal order.  Here the result is checked against a plain scipy product of the standard form, for one GPU and
for every rank of a 2/3/8-way row partition (the multi-GPU layouts cannot be exercised on a one-GPU box)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN
from highs_b200 import engine
from highs_b200.lp import read_b2lp, synthetic_lp


def _check(lp, world, ordered_max=0, exact_rows=False):
    r = engine.host_layout_eval(lp, world=world, ordered_max=ordered_max, seed=world)
    A = sp.csc_matrix((r["cval"], r["cidx"], r["cbeg"]), shape=(r["m"], r["n"]))
    ax_ref = A @ r["x"]
    aty_ref = A.T @ r["y"]
    assert not np.isnan(r["ax"]).any(), "a row was not produced by any rank"
    for part in r["parts"]:
        assert not np.isnan(part).any(), "an A_g' output position was never written"
    scale_ax = np.abs(A) @ np.abs(r["x"]) + 1e-300
    scale_aty = np.abs(A).T @ np.abs(r["y"]) + 1e-300
    assert np.all(np.abs(r["ax"] - ax_ref) <= 1e-13 * scale_ax + 1e-300)
    assert np.all(np.abs(r["aty"] - aty_ref) <= 1e-13 * scale_aty + 1e-300)
    if exact_rows:
        # ordered mode keeps the reference's entry order inside every row: sequential sums, bit for bit
        csr = A.tocsr()
        csr.sort_indices()
        for i in range(r["m"]):
            acc = 0.0
            for q in range(csr.indptr[i], csr.indptr[i + 1]):
                acc += csr.data[q] * r["x"][csr.indices[q]]
            assert acc == r["ax"][i]
    # the row blocks tile [0, m)
    bounds = [(int(s[0]), int(s[1])) for s in r["stats"]]
    assert bounds[0][0] == 0 and bounds[-1][1] == r["m"]
    for (a0, a1), (b0, b1) in zip(bounds[:-1], bounds[1:]):
        assert a1 == b0 and a0 <= a1
    return r


@pytest.mark.parametrize("name", ["avgas", "afiro", "adlittle", "sctest", "chip", "boxed_row", "restart_lp", "e226", "stair", "galenet"])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_layout_golden_lps(name, world):
    path = os.path.join(GOLDEN, name + ".b2lp")
    if not os.path.exists(path):
        pytest.skip("no such golden")
    lp = read_b2lp(path)
    r = _check(lp, world, exact_rows=(world == 1))
    assert int(r["stats"][0][2]) == (1 if world == 1 else 0)
    # and the sorted (tree-mode) layout of the same small LP
    _check(lp, world, ordered_max=-1)


@pytest.mark.parametrize("world", [1, 2, 8])
def test_layout_synthetic_sorted(world):
    lp = synthetic_lp(20000, 17000, 6, seed=7)
    r = _check(lp, world)
    for st in r["stats"]:
        assert st[2] == 0
        assert st[4] == 0 and st[7] == 0          # no long rows
    # padding of the sorted sliced-ELL body stays small
    padded = sum(st[3] for st in r["stats"])
    assert padded <= 1.25 * r["nnz"]


@pytest.mark.parametrize("world", [1, 2, 4])
def test_layout_long_rows_and_columns(world):
    # one dense column (long row of A') and, through a transposed copy, one dense row of A
    lp = synthetic_lp(30000, 9000, 5, seed=3, dense_col_nnz=7000)
    r = _check(lp, world)
    assert sum(st[7] for st in r["stats"]) >= 1       # A' has a split row on some rank
    assert sum(st[8] for st in r["stats"]) >= 2


def test_layout_ragged():
    # empty rows, empty columns, fewer rows than ranks
    from highs_b200.lp import HighsLp, HighsSparseMatrix, kHighsInf
    n, m = 7, 3
    start = np.array([0, 0, 2, 2, 3, 3, 3, 4], dtype=np.int32)
    index = np.array([0, 2, 2, 0], dtype=np.int32)
    value = np.array([1.0, -2.0, 3.0, 4.0])
    lp = HighsLp(n, m, np.ones(n), np.zeros(n), np.full(n, kHighsInf), np.array([1.0, -kHighsInf, 0.0]),
                 np.array([kHighsInf, 5.0, 2.0]), HighsSparseMatrix(n, m, start, index, value), 1, 0.0, "ragged")
    for world in (1, 2, 4):
        _check(lp, world)
        _check(lp, world, ordered_max=-1)


@pytest.mark.parametrize("shape", ["no_rows", "no_cols", "no_nnz", "empty"])
def test_layout_degenerate(shape):
    from highs_b200.lp import HighsLp, HighsSparseMatrix, kHighsInf
    n, m = {"no_rows": (3, 0), "no_cols": (0, 2), "no_nnz": (2, 2), "empty": (0, 0)}[shape]
    lp = HighsLp(n, m, np.ones(n), np.zeros(n), np.full(n, kHighsInf), np.zeros(m), np.full(m, kHighsInf),
                 HighsSparseMatrix(n, m, np.zeros(n + 1, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0)), 1, 0.0, shape)
    f = engine.host_form(lp, 1)
    assert (f["n"], f["m"], f["nnz"]) == (n, m, 0) and f["amax"] == 0.0
    for world in (1, 2, 3):
        r = engine.host_layout_eval(lp, world=world)
        assert np.all(r["ax"] == 0.0) and np.all(r["aty"] == 0.0)
        assert int(r["stats"][0][0]) == 0 and int(r["stats"][-1][1]) == m
