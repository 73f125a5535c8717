#!/usr/bin/env python3
"""Generate tests/golden/ from the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.py).

Run in the build container only (it needs /root/reference for the .mps instances):
    python oracle/build_ref.py && python tests/golden/make_golden.py
Outputs (all committed):
    tests/golden/<name>.b2lp          the HighsLp handed to Highs::run()
    tests/golden/golden.json          per (name, options): reference status, pdlp_iteration_count,
                                      objective and the HighsInfo KKT fields after lpKktCheck
    tests/golden/<name>.<tag>.npz     reference HighsSolution vectors
The special LPs are the ones the reference's own tests build in
check/TestPdlp.cpp:119-285 and check/SpecialLps.h:278-353; the instance LPs are
check/instances/*.mps read by the reference's reader and dumped through
ref_driver --dump-lp.
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from highs_b200.lp import HighsLp, HighsSparseMatrix, read_b2lp, synthetic_lp, write_b2lp  # noqa: E402
from oracle import binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
INST = "/root/reference/check/instances/"
inf = float("inf")


def mk(n, m, c, lo, up, rl, ru, start, index, value, sense=1):
    return HighsLp(n, m, np.array(c, float), np.array(lo, float), np.array(up, float), np.array(rl, float),
                   np.array(ru, float), HighsSparseMatrix(n, m, np.array(start), np.array(index), np.array(value, float)), sense)


SPECIAL = {
    # check/SpecialLps.h:278-296
    "distillation": mk(2, 3, [8, 10], [0, 0], [inf, inf], [7, 12, 6], [inf] * 3, [0, 3, 6], [0, 1, 2, 0, 1, 2], [2, 3, 2, 2, 4, 1]),
    # check/SpecialLps.h:335-353
    "threed": mk(3, 2, [1, 2, 3], [0, 0, 0], [inf] * 3, [-inf, -inf], [3, 2], [0, 1, 2, 4], [0, 1, 0, 1], [1, 1, 2, 2], -1),
    # check/TestPdlp.cpp:150-184
    "boxed_row": mk(2, 2, [-1, -2], [0, 0], [inf, 6], [3, -4], [10, 2], [0, 2, 4], [0, 1, 0, 1], [1, 1, 1, -1]),
    # check/TestPdlp.cpp:186-209
    "infeasible": mk(2, 1, [-1, -2], [0, 0], [inf, inf], [-inf], [-1], [0, 1, 2], [0, 0], [1, 1]),
    # check/TestPdlp.cpp:211-239
    "unbounded": mk(2, 1, [-1, -2], [0, 0], [inf, inf], [1], [inf], [0, 1, 2], [0, 0], [1, 1]),
    # check/TestPdlp.cpp:260-284
    "restart_lp": mk(3, 4, [1, 3, 5], [0, 0, 0], [inf] * 3, [1, 3, 2, -inf], [inf, 3, 10, 5], [0, 4, 8, 12],
                     [0, 1, 2, 3] * 3, [1, 1, 1, 1, 2, 1, 2, 2, 4, 3, 2, 3], -1),
}
INSTANCES = ["avgas", "afiro", "adlittle", "blending", "chip", "sctest", "stair", "e226", "galenet", "woodinfe"]
FIELDS = ["model_status", "model_status_code", "run_status", "pdlp_iteration_count", "objective_function_value",
          "primal_dual_objective_error", "max_primal_infeasibility", "max_dual_infeasibility",
          "max_relative_primal_infeasibility", "max_relative_dual_infeasibility", "max_primal_residual_error",
          "max_dual_residual_error", "max_complementarity_violation", "num_primal_infeasibilities",
          "num_dual_infeasibilities", "primal_solution_status", "dual_solution_status"]


def main():
    cases = []

    def add(name, lp, tag, options, warm=None, save_sol=True):
        res = ob.run_reference(lp=lp, options=options, want_solution=True, warm=warm)
        entry = {"name": name, "tag": tag, "options": options, "warm_from": None}
        entry.update({k: res[k] for k in FIELDS})
        if save_sol:
            np.savez_compressed(os.path.join(HERE, f"{name}.{tag}.npz"), col_value=res["col_value"],
                                col_dual=res["col_dual"], row_value=res["row_value"], row_dual=res["row_dual"])
        cases.append(entry)
        print(f"{name:14s} {tag:10s} {res['model_status']:24s} it={res['pdlp_iteration_count']:6d} obj={res['objective_function_value']:.12g}")
        return res

    for name, lp in SPECIAL.items():
        write_b2lp(os.path.join(HERE, name + ".b2lp"), lp)
        add(name, lp, "kkt1e-4", {"kkt_tolerance": 1e-4})
        if name not in ("infeasible", "unbounded"):
            add(name, lp, "default", {})
    # iteration-limit golden of check/TestPdlp.cpp:53-61 (limit 80 -> kIterationLimit after 79)
    add("distillation", SPECIAL["distillation"], "limit80", {"kkt_tolerance": 1e-4, "pdlp_iteration_limit": 80})
    # pdlp_features_off has no option record in 1.15.1 (HighsOptions.h:403 vs :1339-1385), so scaling /
    # adaptive step cannot be switched off through Highs::run(); only the restart switch is reachable.
    add("distillation", SPECIAL["distillation"], "norestart", {"kkt_tolerance": 1e-4, "pdlp_cupdlpc_restart_method": 0})
    # hot start (check/TestPdlp.cpp:241-285): second run seeded with the first run's HighsSolution
    first = ob.run_reference(lp=SPECIAL["restart_lp"], options={"kkt_tolerance": 1e-4}, want_solution=True)
    warm = (first["col_value"], first["col_dual"], first["row_value"], first["row_dual"])
    add("restart_lp", SPECIAL["restart_lp"], "hotstart", {"kkt_tolerance": 1e-4}, warm=warm)
    cases[-1]["warm_from"] = "restart_lp.kkt1e-4.npz"

    for name in INSTANCES:
        p = os.path.join(HERE, name + ".b2lp")
        subprocess.run([ob.REF_DRIVER, "--mps", INST + name + ".mps", "--dump-lp", p, "--opt", "solver=pdlp",
                        "--opt", "presolve=off", "--opt", "pdlp_iteration_limit=1"], check=True, capture_output=True)
        lp = read_b2lp(p)
        add(name, lp, "default", {})
        add(name, lp, "kkt1e-4", {"kkt_tolerance": 1e-4})
    add("afiro", read_b2lp(os.path.join(HERE, "afiro.b2lp")), "norestart", {"kkt_tolerance": 1e-4, "pdlp_cupdlpc_restart_method": 0})

    # small synthetic LPs of the benchmark family (regenerated from the seed, not stored)
    for (m, n, k, seed) in [(500, 400, 6, 7), (2000, 2000, 8, 12345)]:
        lp = synthetic_lp(m, n, k, seed)
        name = f"synth_{m}x{n}x{k}_s{seed}"
        add(name, lp, "kkt1e-4", {"kkt_tolerance": 1e-4}, save_sol=False)
        cases[-1]["synthetic"] = [m, n, k, seed]
        add(name, lp, "limit400", {"pdlp_iteration_limit": 400}, save_sol=False)
        cases[-1]["synthetic"] = [m, n, k, seed]

    json.dump({"reference": "HiGHS 1.15.1 CPU pdlp via oracle/_ref (solver=pdlp presolve=off)", "cases": cases},
              open(os.path.join(HERE, "golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
