"""Golden for the converged parity check at the headline size: config S3 (m = n = 1 000 000, nnz = 8 000 000, the bench's generator
and seed) solved to kkt_tolerance = 1e-4 by the UNMODIFIED reference (oracle/_ref, Highs::run(), solver=pdlp, presolve=off).
Writes tests/golden/s3_converged.json (status, iterations, objective, HighsInfo KKT fields).
    python tests/golden/make_s3_golden.py          (development container: needs oracle/_ref; a few minutes, ~3 GB)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from highs_b200.lp import synthetic_lp  # noqa: E402
from oracle import binding as ob  # noqa: E402

M, N, K, SEED = 1_000_000, 1_000_000, 8, 12345   # bench.py WORKLOADS["S3"], SEED
lp = synthetic_lp(M, N, K, SEED)
out = {"workload": {"m": M, "n": N, "nnz_per_col": K, "seed": SEED, "nnz": lp.a_matrix_.numNz()}, "runs": {}}
for tol in (1e-4,):
    t = time.time()
    r = ob.run_reference(lp=lp, options={"kkt_tolerance": tol})
    r["wall_seconds"] = time.time() - t
    out["runs"][f"{tol:g}"] = r
    print(tol, r["model_status"], r["pdlp_iteration_count"], r["objective_function_value"], round(r["wall_seconds"], 1), flush=True)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "s3_converged.json"), "w"), indent=1)
