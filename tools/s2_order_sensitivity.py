"""How much does the ORDER of the long sums alone move the converged iteration count on S2?  The oracle (CPU restatement of
cuPDLP-C, pinned to the reference) is run with its sequential sums (= the reference) and with blocked sums of several
widths (orc_set_sum_block: a test switch).  Output: one JSON line per run -> profiles/r02_s2_order_sensitivity.json
    python tools/s2_order_sensitivity.py <tol> <block> [<block> ...]"""
import json
import sys
import time

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from highs_b200.lp import synthetic_lp  # noqa: E402
from oracle import binding as ob  # noqa: E402

tol = float(sys.argv[1])
lp = synthetic_lp(100_000, 100_000, 10, 12345)
L = ob.lib()
for blk in [int(a) for a in sys.argv[2:]]:
    L.orc_set_sum_block(blk)
    t = time.time()
    r = ob.solve(lp, iter_limit=400000, tol_primal=tol, tol_dual=tol, tol_gap=tol)
    print(json.dumps(dict(tol=tol, sum_block=blk, iters=r["iters"], term=r["term_code"], restarts=r.get("restarts"),
                          objective=lp.objectiveValue(r["col_value"]), seconds=round(time.time() - t, 1))), flush=True)
L.orc_set_sum_block(0)
