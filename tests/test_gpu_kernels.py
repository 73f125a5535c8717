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


# ---- fused hot-path kernels through their C-ABI entry points (_primal_step, _dual_step, _residuals)
def _form(ob, lp):
    return ob.formulate_and_scale(lp)


@pytest.mark.parametrize("m,n,k", [(3000, 2500, 6), (60000, 50000, 8)])
def test_primal_and_dual_step_kernels(engine_lib, oracle, m, n, k):
    """K1 / K2 against PDHG_primalGradientStep / PDHG_dualGradientStep (cupdlp_step.c:16-69) evaluated in the reference's
    CPU operation order on the oracle's scaled form: element-wise results bit for bit, the two norms to rounding"""
    import scipy.sparse as sp
    from highs_b200 import engine
    from highs_b200.lp import synthetic_lp
    lp = synthetic_lp(m, n, k, seed=17)
    lp.col_upper_[::7] = 1.5    # some finite upper bounds
    f = _form(oracle, lp)
    prob = engine.Problem(lp)
    rng = np.random.default_rng(2)
    x, aty = rng.standard_normal(prob.n), rng.standard_normal(prob.n)
    y, ax = rng.standard_normal(prob.m), rng.standard_normal(prob.m)
    tau, sigma = 0.37, 0.81
    xn, dx2 = prob.primal_step(x, aty, tau)
    v = x + (-tau) * f["cost"]
    v = v + tau * aty
    v = np.minimum(v, f["upper"])
    v = np.maximum(v, f["lower"])
    assert np.array_equal(xn, v)
    assert abs(dx2 - np.sum((x - v) ** 2)) <= 1e-12 * (1 + dx2)
    yn, axn, dy2 = prob.dual_step(xn, y, ax, sigma)
    ax_o, _ = _oracle_spmv(oracle, lp, xn, y)   # the reference's scatter order
    assert np.array_equal(axn, ax_o)
    w = y + sigma * f["rhs"]
    w = w + (-2.0 * sigma) * ax_o
    w = w + sigma * ax
    w[f["neq"]:] = np.maximum(w[f["neq"]:], 0.0)
    assert np.array_equal(yn, w)
    assert abs(dy2 - np.sum((y - w) ** 2)) <= 1e-12 * (1 + dy2)
    prob.close()


def test_residual_kernels(engine_lib, oracle):
    """the check-iteration sums (PDHG_Compute_Residuals, cupdlp_solver.c:12-204, :473-529) of a given (x, y) against numpy
    on the oracle's scaled form"""
    import scipy.sparse as sp
    from highs_b200 import engine
    from highs_b200.lp import synthetic_lp
    lp = synthetic_lp(40000, 30000, 7, seed=23)
    f = _form(oracle, lp)
    prob = engine.Problem(lp)
    rng = np.random.default_rng(4)
    x, y = np.abs(rng.standard_normal(prob.n)), rng.standard_normal(prob.m)
    y[f["neq"]:] = np.abs(y[f["neq"]:])
    got = prob.residuals(x, y)
    A = sp.csr_matrix((f["rval"], f["ridx"], f["rbeg"]), shape=(f["m"], f["n"]))
    ax, aty = A @ x, A.T @ y
    pobj = float(f["cost"] @ x)
    r = ax - f["rhs"]
    r[f["neq"]:] = np.minimum(r[f["neq"]:], 0.0)
    pfeas = np.linalg.norm(r * f["row_scale"])
    rc = f["cost"] - aty
    hl, hu = f["lower"] > -np.inf, f["upper"] < np.inf
    spos, sneg = np.where(hl, np.maximum(rc, 0), 0.0), np.where(hu, -np.minimum(rc, 0), 0.0)
    dobj = float(f["rhs"] @ y + np.where(hl, f["lower"], 0.0) @ spos - np.where(hu, f["upper"], 0.0) @ sneg)
    dfeas = np.linalg.norm((rc - spos + sneg) * f["col_scale"])
    for k, ref in (("pobj", pobj), ("dobj", dobj), ("pfeas", pfeas), ("dfeas", dfeas)):
        assert abs(got[k] - ref) <= 1e-11 * (1 + abs(ref)), (k, got[k], ref)
    assert abs(got["gap"] - (pobj - dobj)) <= 1e-10 * (1 + abs(pobj) + abs(dobj))
    prob.close()
