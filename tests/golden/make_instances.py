"""Dumps every LP instance the reference ships for its own tests (/root/reference/check/instances/*.mps) in the .b2lp
layout (oracle/ref_driver.cpp header), as read by the UNMODIFIED reference's own MPS reader (oracle/_ref/ref_driver
--dump-lp).  Development container only; the dumps travel to the GPU box as fixtures (tests/test_gpu_instances.py).
usage: python tests/golden/make_instances.py"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from highs_b200.lp import read_b2lp  # noqa: E402


def main():
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    out_dir = os.path.join(HERE, "instances")
    os.makedirs(out_dir, exist_ok=True)
    kept = 0
    for mps in sorted(glob.glob("/root/reference/check/instances/*.mps")):
        name = os.path.basename(mps)[:-4]
        out = os.path.join(out_dir, name + ".b2lp")
        subprocess.run([drv, "--mps", mps, "--dump-lp", out, "--opt", "solver=pdlp", "--opt", "pdlp_iteration_limit=1",
                        "--opt", "presolve=off"], capture_output=True, text=True, timeout=300)
        if not os.path.exists(out):
            continue   # the reference's reader rejects the file (its own negative tests)
        lp = read_b2lp(out)
        if lp.num_row_ == 0 or lp.a_matrix_.numNz() == 0 or np.isnan(lp.col_cost_).any():
            os.remove(out)   # not a PDLP problem (HighsSolve.cpp:61,68) / invalid data
            continue
        kept += 1
    print("kept", kept, "instances in", out_dir)


if __name__ == "__main__":
    main()
