"""Child process of tests/test_gpu_hipdlp.py (own process: a faulting kernel must not poison the CUDA context of the test
session).  The HiPDLP mode of the engine against the pinned HiPDLP oracle: iteration counts, termination and the four
HighsSolution vectors bit for bit (problems up to 4096 rows / columns run their checks in the reference's summation order;
the Halpern steps themselves contain no reduction at any size).
usage: python tests/hipdlp_child.py   -> one JSON line {case: "ok" | error text}"""
import glob
import json
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from highs_b200 import engine  # noqa: E402
from highs_b200.lp import read_b2lp  # noqa: E402
from oracle import binding as ob  # noqa: E402

VARIANTS = [dict(iter_limit=800), dict(iter_limit=4000, tolerance=1e-4), dict(iter_limit=400, step_size_strategy=0),
            dict(iter_limit=400, scaling_mode=7, ruiz_iterations=3), dict(iter_limit=1200, tolerance=1e-3, scaling_mode=0)]


def main():
    out = {}
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "instances", "*.b2lp")))
    for path in files:
        name = os.path.basename(path)[:-5]
        lp = read_b2lp(path)
        if lp.a_matrix_.numNz() > 200000:
            continue
        for vi, prm in enumerate(VARIANTS if name in ("afiro", "adlittle", "sctest", "e226", "25fv47") else VARIANTS[:2]):
            key = f"{name}#{vi}"
            try:
                okw = dict(max_iterations=prm.get("iter_limit"), tolerance=prm.get("tolerance", 1e-7),
                           scaling_mode=prm.get("scaling_mode", 5), ruiz_iterations=prm.get("ruiz_iterations", 10),
                           step_size_strategy=prm.get("step_size_strategy", 3))
                ref = ob.hipdlp_solve(lp, **okw)
                res = engine.solve_hipdlp(lp, **prm)
                assert res["iters"] == ref["iters"], ("iterations", res["iters"], ref["iters"])
                assert (res["term_code"] == 0) == (ref["term_code"] == 0), "termination"
                if max(res["form_cols"], res["form_rows"]) <= 4096:
                    for k in ("col_value", "col_dual", "row_value", "row_dual"):
                        assert np.array_equal(res[k], ref[k]), k
                else:
                    for k in ("col_value", "row_dual"):
                        assert np.allclose(res[k], ref[k], rtol=1e-7, atol=1e-9 * (1 + np.abs(ref[k]).max())), k
                out[key] = "ok"
            except Exception as e:   # noqa: BLE001
                out[key] = f"{type(e).__name__}: {e}"[:300]
                if "CUDA" in out[key] or "cuda" in out[key]:
                    traceback.print_exc()
                    print(json.dumps(out))
                    return 1
    print(json.dumps(out))
    return 0 if all(v == "ok" for v in out.values()) else 1


if __name__ == "__main__":
    sys.exit(main())
