"""The product's host prologue of the HiPDLP mode (highs_b200/csrc/host_prep_hipdlp.cpp: preprocess, Ruiz / Pock-Chambolle /
L2 scaling, power method) against the pinned HiPDLP oracle, bit for bit, on the reference's own LP instances and on LPs with
free / ranged rows, unsorted columns and infinite bounds.  No GPU."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN

FILES = sorted(glob.glob(os.path.join(GOLDEN, "instances", "*.b2lp")))
VEC = ["cost", "lower", "upper", "rlo", "rup", "col_scale", "row_scale", "cbeg", "cidx", "cval", "new_idx"]
ORACLE_TO_PRODUCT_CLASS = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4}   # EQ, LEQ, GEQ, BOUND, FREE


def _compare(engine, oracle, lp, mode, ruiz):
    a = engine.host_form_hipdlp(lp, mode, ruiz)
    b = oracle.hipdlp_form(lp, mode, ruiz)
    for k in ("n", "m", "nnz", "neq"):
        assert a[k] == b[k], k
    for k in VEC:
        assert np.array_equal(a[k], b[k], equal_nan=a[k].dtype.kind == "f"), k
    assert np.array_equal(a["ctype"], b["ctype"])
    assert a["c_norm"] == b["c_norm"] and a["rhs_norm"] == b["rhs_norm"]
    assert a["op_norm_sq"] == b["op_norm_sq"] or (np.isnan(a["op_norm_sq"]) and np.isnan(b["op_norm_sq"]))


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-5] for f in FILES])
def test_hipdlp_host_prologue_instances(engine_lib, oracle, path):
    from highs_b200 import engine
    from highs_b200.lp import read_b2lp
    lp = read_b2lp(path)
    _compare(engine, oracle, lp, 5, 10)        # the default: Ruiz x10 + Pock-Chambolle
    _compare(engine, oracle, lp, 7, 3)         # + L2, three Ruiz passes
    _compare(engine, oracle, lp, 0, 10)        # no scaling


def test_hipdlp_host_prologue_row_kinds(engine_lib, oracle):
    from highs_b200 import engine
    from highs_b200.lp import HighsLp, HighsSparseMatrix
    inf = np.inf
    # rows: GEQ, EQ, ranged, LEQ, free; column 1 stored with descending rows; a 1e25 bound is NOT infinite for HiPDLP
    lp = HighsLp(2, 5, [1, 2], [0, -1e25], [inf, 5], [1, 3, 2, -inf, -inf], [inf, 3, 10, 5, inf],
                 HighsSparseMatrix(2, 5, [0, 5, 10], [0, 1, 2, 3, 4, 4, 3, 2, 1, 0], [1, 2, 3, 4, 5, 6, 7, 8, 9, 10.0]), -1, 0.5)
    for mode in (0, 1, 4, 5, 7):
        _compare(engine, oracle, lp, mode, 4)
    f = engine.host_form_hipdlp(lp, 0, 0)
    assert list(f["ctype"]) == [2, 0, 3, 1, 4]
    assert f["n"] == 4 and f["neq"] == 3
    assert list(f["new_idx"]) == [3, 0, 1, 4, 2]
    assert list(f["cost"]) == [1, 2, 0, 0]                     # NOT multiplied by the sense (pdhg.cc:226-230)
    assert list(f["rlo"]) == [3, 0, 0, 1, -5] and list(f["rup"]) == [3, 0, 0, inf, inf]
    assert list(f["cidx"][:5]) == [0, 1, 2, 3, 4] and list(f["cval"][:5]) == [2, 3, 5, 1, -4]
    assert list(f["cidx"][5:10]) == [0, 1, 2, 3, 4]            # re-sorted by new row


@pytest.mark.parametrize("name,kw", [("afiro", {}), ("adlittle", {}), ("sctest", dict(step_size_strategy=0)),
                                     ("e226", dict(tolerance=1e-4)), ("25fv47", {}), ("stair", dict(scaling_mode=7, ruiz_iterations=3))])
def test_hipdlp_host_control_replay(engine_lib, oracle, name, kw):
    """The product's host control of the Halpern loop (HipController: fixed-point error, convergence test, restart criteria,
    PID primal weight, step sizes) replayed over the sums the ORACLE recorded block by block must take the oracle's decisions
    and reproduce its weights and step sizes bit for bit."""
    from highs_b200 import engine
    from highs_b200.lp import read_b2lp
    lp = read_b2lp(os.path.join(GOLDEN, "instances", name + ".b2lp"))
    prm = dict(tolerance=1e-7, scaling_mode=5, ruiz_iterations=10, step_size_strategy=3)
    prm.update(kw)
    ref = oracle.hipdlp_solve(lp, max_iterations=6000, trace_cap=200, **prm)
    tr = ref["trace"]
    assert len(tr) >= 5
    form = oracle.hipdlp_form(lp, prm["scaling_mode"], prm["ruiz_iterations"])
    out = engine.hipdlp_controller_replay(form["c_norm"], form["rhs_norm"], form["op_norm_sq"], prm["tolerance"],
                                          prm["step_size_strategy"], tr[:, 1:10], tr[:, 10:13])
    assert np.array_equal(out[:, 6], tr[:, 0])                       # iteration counts
    assert np.array_equal(out[:, 4], tr[:, 18])                      # fixed-point errors
    assert np.array_equal(out[:, 5], tr[:, 19])                      # convergence decisions
    live = tr[:, 19] == 0
    assert np.array_equal(out[live, 0], tr[live, 14])                # restart decisions
    did = live & (tr[:, 14] == 1)
    assert did.sum() >= 2
    for col_out, col_tr in ((1, 15), (2, 16), (3, 17)):              # primal weight, primal / dual step after each restart
        assert np.array_equal(out[did, col_out], tr[did, col_tr])
    # a block that follows a restart carries the reference sums of its first step
    assert np.array_equal(tr[1:, 13] == 1, tr[:-1, 14] == 1)
