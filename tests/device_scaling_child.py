"""Child process of tests/test_gpu_device_scaling.py (own process: a faulting kernel must not poison the CUDA context of
the test session).  usage: python tests/device_scaling_child.py <level>   -> one JSON line {case: "ok" | error text}"""
import json
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from highs_b200 import engine  # noqa: E402
from highs_b200.lp import HighsLp, HighsSparseMatrix, kHighsInf, read_b2lp, synthetic_lp  # noqa: E402

VECS = ["cost", "lower", "upper", "rhs", "col_scale", "row_scale"]


def compare(lp, level, **prm):
    a = engine.Problem(lp, device_scaling=-1, **prm)    # the host-thread prologue: the bit-reference of the staged variants
    b = engine.Problem(lp, device_scaling=level, **prm)
    try:
        for k in VECS:
            assert np.array_equal(a.vector(k), b.vector(k)), k
        for u, v in zip(a.csr(), b.csr()):
            assert np.array_equal(u, v), "csr"
        # the device layouts (level 2: filled on the GPU from the host's plan): same products, bit for bit
        rng = np.random.default_rng(3)
        x, y = rng.standard_normal(max(a.n, 1))[: a.n], rng.standard_normal(max(a.m, 1))[: a.m]
        assert np.array_equal(a.spmv_ax(x), b.spmv_ax(x)), "A x"
        assert np.array_equal(a.spmv_aty(y), b.spmv_aty(y)), "A' y"
        sa = a.solve(trace_cap=64, **prm)
        sb = b.solve(trace_cap=64, **prm)
        assert sa["iters"] == sb["iters"] and sa["term_code"] == sb["term_code"], "iterations"
        assert np.array_equal(sa["trace"], sb["trace"]), "trace"    # same amax -> same first step -> same trajectory
        for k in ("col_value", "col_dual", "row_value", "row_dual"):
            assert np.array_equal(sa[k], sb[k]), k
    finally:
        a.close()
        b.close()


def ragged():
    # empty rows and columns (norm 0 -> factor 1), an explicit zero entry, a ranged row
    n, m = 7, 4
    start = np.array([0, 0, 2, 2, 3, 3, 4, 5], dtype=np.int32)
    index = np.array([0, 2, 2, 3, 0], dtype=np.int32)
    value = np.array([1.0, -2.0, 3.0, 0.0, 4.0])
    return HighsLp(n, m, np.ones(n), np.zeros(n), np.full(n, kHighsInf), np.array([1.0, -kHighsInf, 0.0, -1.0]),
                   np.array([kHighsInf, 5.0, 2.0, 1.0]), HighsSparseMatrix(n, m, start, index, value), 1, 0.0, "ragged")


def main():
    level = int(sys.argv[1])
    cases = [(name, lambda name=name: read_b2lp(os.path.join(ROOT, "tests", "golden", name + ".b2lp")), dict(iter_limit=400))
             for name in ["avgas", "afiro", "adlittle", "sctest", "boxed_row", "e226", "stair"]]
    cases += [(name + "/sorted", mk, dict(iter_limit=400, ordered_max=-1)) for name, mk, _ in cases[:7]]
    cases.append(("ragged", ragged, dict(iter_limit=50)))
    cases.append(("synthetic_dense_column", lambda: synthetic_lp(120000, 90000, 8, 4, dense_col_nnz=30000), dict(iter_limit=200)))
    out = {}
    for name, mk, prm in cases:
        try:
            compare(mk(), level, **prm)
            out[name] = "ok"
        except Exception as e:   # noqa: BLE001
            out[name] = f"{type(e).__name__}: {e}"[:300]
            if "CUDA" in out[name] or "cuda" in out[name]:
                traceback.print_exc()
                break            # the context is gone: stop here
    print(json.dumps(out))
    return 0 if all(v == "ok" for v in out.values()) else 1


if __name__ == "__main__":
    sys.exit(main())
