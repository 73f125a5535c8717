"""Scratch GPU run: first timings (not a bench line)."""
import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
from highs_b200 import engine
from highs_b200.lp import synthetic_lp, read_b2lp
from oracle import binding as ob

for nm, kw in [("avgas", {}), ("adlittle", {})]:
    lp = read_b2lp(f'/root/repo/tests/golden/{nm}.b2lp')
    t = time.time(); r = engine.solve(lp, trace_cap=2000, **kw); dt = time.time() - t
    o = ob.solve(lp, trace_cap=2000, **kw)
    print(nm, r["term_name"], "iters", r["iters"], "oracle", o["iters"], "obj", lp.objectiveValue(r["col_value"]), lp.objectiveValue(o["col_value"]),
          "passes", r["passes"], "restarts", r["restarts"], f"{dt:.3f}s", flush=True)
    k = min(len(r["trace"]), len(o["trace"]))
    bad = np.argwhere(r["trace"][:k, :15] != o["trace"][:k, :15])
    print("  trace rows", len(r["trace"]), len(o["trace"]), "first mismatch", bad[:3].tolist(), flush=True)
    if len(bad):
        i = bad[0][0]
        print("  gpu", r["trace"][i].tolist()); print("  orc", o["trace"][i].tolist())
