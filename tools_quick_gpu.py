"""Scratch GPU run: parity spot checks + first timings (not a bench line)."""
import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
from highs_b200 import engine
from highs_b200.lp import synthetic_lp, read_b2lp
from oracle import binding as ob

print("devices", engine.device_count(), flush=True)
for nm, kw in [("avgas", {}), ("afiro", {}), ("distillation", dict(tol_primal=1e-4, tol_dual=1e-4, tol_gap=1e-4)), ("adlittle", {})]:
    lp = read_b2lp(f'/root/repo/tests/golden/{nm}.b2lp')
    t = time.time(); r = engine.solve(lp, **kw); dt = time.time() - t
    o = ob.solve(lp, **kw)
    print(nm, r["term_name"], "iters", r["iters"], "oracle", o["iters"], "obj", lp.objectiveValue(r["col_value"]), lp.objectiveValue(o["col_value"]),
          "passes", r["passes"], "restarts", r["restarts"], f"{dt:.3f}s", flush=True)
for (m, n, k) in [(100000, 100000, 10), (1000000, 1000000, 8)]:
    t = time.time(); lp = synthetic_lp(m, n, k, 12345); print("gen", time.time() - t, flush=True)
    t = time.time(); prob = engine.Problem(lp); print("create", time.time() - t, "nnz", prob.nnz, flush=True)
    for which in (0, 1):
        prob.bench_spmv(which, 5)
        ms = prob.bench_spmv(which, 50) / 50
        B = 12 * prob.nnz + 4 * (m + 1) + 8 * n + 8 * m
        print(("Ax", "ATy")[which], f"{ms*1e3:.1f} us  {B/ms/1e6:.0f} GB/s", flush=True)
    r = prob.solve(iter_limit=2000)
    print("solve2000", r["term_name"], r["iters"], "passes", r["passes"], "dev ms", r["iter_device_ms"], "it/s", r["iters"] / (r["iter_device_ms"] / 1e3),
          "wall", r["solve_seconds"], "launches", r["kernel_launches"], "pobj", r["primal_obj"], "dobj", r["dual_obj"], flush=True)
    if m == 100000:
        o = ob.solve(lp, iter_limit=400)
        r2 = prob.solve(iter_limit=400)
        print("S2 400 its: gpu pobj", r2["primal_obj"], "oracle", o["pobj"], "dobj", r2["dual_obj"], o["dobj"], "pfeas", r2["primal_feas"], o["pfeas"], flush=True)
    prob.close()
