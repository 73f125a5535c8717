"""Golden results of the UNMODIFIED reference's HiPDLP engine (`solver=hipdlp`, CPU) through Highs::run()
(oracle/_ref/ref_driver) on a subset of the reference's own LP instances (tests/golden/instances/) under several
option sets: iteration count, model status and the four HighsSolution vectors.  Development container only.
usage: python tests/golden/make_hipdlp_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from highs_b200.lp import read_b2lp  # noqa: E402
from oracle import binding as ob  # noqa: E402

NAMES = ["afiro", "adlittle", "avgas", "blending", "chip", "sctest", "e226", "stair", "25fv47", "shell", "bgetam", "scrs8"]
VARIANTS = {
    "limit800": {"pdlp_iteration_limit": 800},
    "kkt1e-4": {"pdlp_iteration_limit": 4000, "kkt_tolerance": 1e-4},
    "fixed_step": {"pdlp_iteration_limit": 400, "pdlp_step_size_strategy": 0},
    "scaling7_ruiz3": {"pdlp_iteration_limit": 400, "pdlp_scaling_mode": 7, "pdlp_ruiz_iterations": 3},
    "noscale_1e-3": {"pdlp_iteration_limit": 1200, "kkt_tolerance": 1e-3, "pdlp_scaling_mode": 0},
}


def main():
    out_dir = os.path.join(HERE, "hipdlp")
    os.makedirs(out_dir, exist_ok=True)
    index = []
    for name in NAMES:
        path = os.path.join(HERE, "instances", name + ".b2lp")
        if not os.path.exists(path):
            continue
        lp = read_b2lp(path)
        for tag, opts in VARIANTS.items():
            o = dict(opts)
            o["solver"] = "hipdlp"
            r = ob.run_reference(lp=lp, options=o, want_solution=True)
            f = f"{name}__{tag}.npz"
            np.savez_compressed(os.path.join(out_dir, f), col_value=r["col_value"], col_dual=r["col_dual"],
                                row_value=r["row_value"], row_dual=r["row_dual"])
            index.append(dict(name=name, tag=tag, options=opts, file=f, pdlp_iteration_count=r["pdlp_iteration_count"],
                              model_status=r.get("model_status"), objective_function_value=r.get("objective_function_value")))
    json.dump(index, open(os.path.join(out_dir, "golden.json"), "w"), indent=1)
    print(len(index), "cases")


if __name__ == "__main__":
    main()
