"""The reference's own test instances (tests/golden/instances/*.b2lp, dumped by tests/golden/make_instances.py through the
unmodified reference's MPS reader): the product's host prologue against the oracle bit for bit, and the device layouts of
one and two ranks evaluated on the host.  No GPU, no reference tree needed."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN

FILES = sorted(glob.glob(os.path.join(GOLDEN, "instances", "*.b2lp")))
KEYS = ["cost", "lower", "upper", "rhs", "col_scale", "row_scale", "cbeg", "cidx", "cval", "row_new_idx", "row_type",
        "rbeg", "ridx", "rval"]


def test_fixture_set_is_there():
    assert len(FILES) >= 70


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-5] for f in FILES])
def test_instance_host_prologue_and_layouts(engine_lib, oracle, path):
    from highs_b200 import engine
    from highs_b200.lp import read_b2lp
    lp = read_b2lp(path)
    a, b = engine.host_form(lp, 1), oracle.formulate_and_scale(lp, 1)
    for k in ("n", "m", "nnz", "neq", "n_orig", "norm_cost", "norm_rhs", "amax"):
        assert a[k] == b[k], k
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k
    for world, omax in ((1, 0), (2, 0), (1, -1)):
        r = engine.host_layout_eval(lp, world=world, ordered_max=omax, seed=5)
        A = sp.csc_matrix((r["cval"], r["cidx"], r["cbeg"]), shape=(r["m"], r["n"]))
        sax = np.abs(A) @ np.abs(r["x"]) + 1e-300
        say = np.abs(A).T @ np.abs(r["y"]) + 1e-300
        assert np.all(np.abs(r["ax"] - A @ r["x"]) <= 1e-12 * sax)
        assert np.all(np.abs(r["aty"] - A.T @ r["y"]) <= 1e-12 * say)
