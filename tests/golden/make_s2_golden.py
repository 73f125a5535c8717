"""Golden for the tight-tolerance parity check of SURVEY.md 8(d): config S2 (m = n = 100 000, nnz = 1 000 000, the bench's
generator and seed) solved to kkt_tolerance = 1e-4, 1e-6 and 1e-8 by the UNMODIFIED reference (oracle/_ref, Highs::run(), solver=pdlp,
presolve=off).  Writes tests/golden/s2_converged.json (status, iterations, objective, HighsInfo KKT fields).
    python tests/golden/make_s2_golden.py          (development container: needs oracle/_ref; ~2 minutes for 1e-4 and 1e-6, ~51 minutes for 1e-8)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from highs_b200.lp import synthetic_lp  # noqa: E402
from oracle import binding as ob  # noqa: E402

M, N, K, SEED = 100_000, 100_000, 10, 12345   # bench.py WORKLOADS["S2"], SEED
lp = synthetic_lp(M, N, K, SEED)
out = {"workload": {"m": M, "n": N, "nnz_per_col": K, "seed": SEED, "nnz": lp.a_matrix_.numNz()}, "runs": {}}
for tol in (1e-4, 1e-6, 1e-8):   # (1e-8: 362 800 iterations, 51 minutes on the reference's single CPU thread)
    t = time.time()
    r = ob.run_reference(lp=lp, options={"kkt_tolerance": tol})
    r["wall_seconds"] = time.time() - t
    out["runs"][f"{tol:g}"] = r
    print(tol, r["model_status"], r["pdlp_iteration_count"], r["objective_function_value"], round(r["wall_seconds"], 1), flush=True)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "s2_converged.json"), "w"), indent=1)
