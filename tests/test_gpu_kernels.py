"""GPU parity of the SpMV kernels against the oracle's scatter SpMV (bit-exact)."""
import numpy as np
import pytest

from conftest import golden_lp, load_golden

pytestmark = pytest.mark.gpu


def _oracle_spmv(ob, lp, x, y):
    import ctypes as C
    L = ob.lib()
    clp, keep = ob._mk_lp(lp)
    f = ob.OrcForm()
    L.orc_formulate(C.byref(clp), C.byref(f))
    L.orc_scale(C.byref(f), 1)
    L.orc_build_csr(C.byref(f))
    ax, aty = np.zeros(max(f.m, 1)), np.zeros(max(f.n, 1))
    dp = C.POINTER(C.c_double)
    L.orc_ax(C.byref(f), x.ctypes.data_as(dp), ax.ctypes.data_as(dp))
    L.orc_aty(C.byref(f), y.ctypes.data_as(dp), aty.ctypes.data_as(dp))
    out = ax[: f.m].copy(), aty[: f.n].copy()
    L.orc_form_free(C.byref(f))
    return out


def _check(engine, ob, lp, seed=0):
    prob = engine.Problem(lp)
    rng = np.random.default_rng(seed)
    x, y = rng.standard_normal(prob.n), rng.standard_normal(prob.m)
    ax_o, aty_o = _oracle_spmv(ob, lp, x, y)
    ax, aty = prob.spmv_ax(x), prob.spmv_aty(y)
    prob.close()
    return ax, ax_o, aty, aty_o


@pytest.mark.parametrize("name", ["avgas", "afiro", "adlittle", "boxed_row", "restart_lp", "e226", "stair"])
def test_spmv_bit_exact_goldens(engine_lib, oracle, name):
    from highs_b200 import engine
    from highs_b200.lp import read_b2lp
    import os
    from conftest import GOLDEN
    lp = read_b2lp(os.path.join(GOLDEN, name + ".b2lp"))
    ax, ax_o, aty, aty_o = _check(engine, oracle, lp)
    assert np.array_equal(ax, ax_o)
    assert np.array_equal(aty, aty_o)


@pytest.mark.parametrize("m,n,k", [(5000, 4000, 7), (100000, 100000, 10), (20000, 300, 3), (300, 20000, 2)])
def test_spmv_bit_exact_synthetic(engine_lib, oracle, m, n, k):
    from highs_b200 import engine
    from highs_b200.lp import synthetic_lp
    lp = synthetic_lp(m, n, k, seed=5)
    ax, ax_o, aty, aty_o = _check(engine, oracle, lp)
    assert np.array_equal(ax, ax_o)
    assert np.array_equal(aty, aty_o)


def test_spmv_long_rows(engine_lib, oracle):
    """One 50%-dense column (config S5 in miniature): the A^T row of 10 000 nonzeros is split into
    segment blocks; their tree-summed partials are not in the oracle's sequential order, so the long
    row is compared to 1e-13 relative, everything else bit-exactly."""
    from highs_b200 import engine
    from highs_b200.lp import synthetic_lp
    lp = synthetic_lp(20000, 20000, 6, seed=9, dense_col_nnz=10000)
    ax, ax_o, aty, aty_o = _check(engine, oracle, lp)
    assert np.array_equal(ax, ax_o)
    assert np.array_equal(aty[1:], aty_o[1:])
    assert abs(aty[0] - aty_o[0]) <= 1e-13 * max(1.0, np.abs(aty_o).max())
