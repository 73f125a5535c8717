"""Round-2 diagnostic: the two tree-mode instance cases that failed on hardware in round 1 -- which assertion, how far."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from highs_b200 import engine
from highs_b200.lp import read_b2lp
from oracle import binding as ob
ob.build()
for name in ("80bau3b", "greenbea", "25fv47", "ship12l"):
    p = os.path.join(ROOT, "tests", "golden", "instances", name + ".b2lp")
    if not os.path.exists(p):
        continue
    lp = read_b2lp(p)
    for lim in (40, 120, 400):
        ref = ob.solve(lp, iter_limit=lim)
        out = engine.solve(lp, iter_limit=lim)
        d = {k: float(np.abs(out[k] - ref[k]).max()) for k in ("col_value", "col_dual", "row_value", "row_dual")}
        s = {k: float(np.abs(ref[k]).max()) for k in ("col_value", "col_dual", "row_value", "row_dual")}
        print(json.dumps(dict(name=name, lim=lim, form=(out["form_rows"], out["form_cols"]), term=(out["term_code"], ref["term_code"]),
                              iters=(out["iters"], ref["iters"]), maxdiff=d, scale=s)), flush=True)
    # forced ordered mode on the same instance: is the engine bit-exact when the reductions are ordered?
    try:
        out = engine.solve(lp, iter_limit=400, ordered_max=1 << 30)
        ref = ob.solve(lp, iter_limit=400)
        print(json.dumps(dict(name=name, ordered=True, iters=(out["iters"], ref["iters"]),
                              equal={k: bool(np.array_equal(out[k], ref[k])) for k in ("col_value", "col_dual", "row_value", "row_dual")})), flush=True)
    except Exception as e:
        print("ordered failed", name, repr(e), flush=True)
