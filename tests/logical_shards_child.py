"""Child process of tests/test_gpu_logical_shards.py: the multi-GPU code path (row blocks + column shards, the fused
peer-memory kernels, flag barriers, device-side step rule, speculative check, NCCL-free solution assembly) with G
LOGICAL shards on ONE device -- SURVEY.md 8(e): "the same code path must run with G logical shards on 1 device".
Run in its own process so that a stuck barrier can be killed by PID without taking the test session along.
usage: python tests/logical_shards_child.py <world> <case>      prints one JSON line and exits 0 on success"""
import json
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per logical rank's stream
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from highs_b200 import engine  # noqa: E402
from highs_b200.lp import read_b2lp, synthetic_lp  # noqa: E402


def main():
    world, case = int(sys.argv[1]), sys.argv[2]
    tol = 1e-5
    if case == "synthetic":
        lp = synthetic_lp(6000, 5000, 6, seed=11)
    elif case == "dense":
        lp = synthetic_lp(9000, 7000, 5, seed=5, dense_col_nnz=4000)     # long rows of A' on every rank
    else:
        lp = read_b2lp(os.path.join(ROOT, "tests", "golden", case + ".b2lp"))
        tol = 1e-7
    prm = dict(iter_limit=400000, tol_primal=tol, tol_dual=tol, tol_gap=tol)
    # the ORACLE (CPU restatement pinned to the reference) is what the shards must agree with; the single-GPU engine is
    # reported beside it
    from oracle import binding as ob
    ref = ob.solve(lp, **prm)
    one = engine.solve(lp, ordered_max=-1, **prm)
    if len(sys.argv) > 3 and sys.argv[3] == "c_entry":
        # the same thing through the one-call C entry point the HiGHS shim uses for B200PDLP_GPUS > 1, all ranks on device 0
        r = engine.solve_multi(lp, world, devices=[0] * world, **prm)
        res = [r, r]
    else:
        res = engine.solve_logical_shards(lp, world, **prm)
    o_ref = lp.objectiveValue(ref["col_value"])
    out = dict(world=world, case=case, ref_iters=ref["iters"], ref_term=ref["term_code"], ref_obj=o_ref,
               one_gpu_iters=one["iters"], one_gpu_obj=lp.objectiveValue(one["col_value"]),
               iters=[r["iters"] for r in res], term=[r["term_code"] for r in res],
               obj=[lp.objectiveValue(r["col_value"]) for r in res])
    ok = True
    for r in res[1:]:   # every rank takes the same decisions and returns the same (complete) solution
        for k in ("col_value", "col_dual", "row_value", "row_dual"):
            ok &= bool(np.array_equal(r[k], res[0][k]))
        ok &= r["iters"] == res[0]["iters"] and r["term_code"] == res[0]["term_code"]
    out["ranks_identical"] = ok
    r0 = res[0]
    ok &= r0["term_code"] == ref["term_code"] == 0
    # both runs stop at relative gap < tol: they bracket the optimum to ~tol (1 + |p| + |d|) each
    ok &= abs(out["obj"][0] - o_ref) <= 3 * tol * (1 + 2 * abs(o_ref))
    ok &= abs(r0["iters"] - ref["iters"]) <= 0.35 * ref["iters"] + 80
    # the assembled solution in the ORIGINAL space: row_value must be A x
    A = lp.a_matrix_.to_scipy()
    ok &= bool(np.allclose(A @ r0["col_value"], r0["row_value"], rtol=1e-8, atol=1e-7 * (1 + np.abs(r0["row_value"]).max())))
    out["ok"] = bool(ok)
    print(json.dumps(out))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
