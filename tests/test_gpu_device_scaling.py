"""params.device_scaling: 1 = PDHG_Scale_Data (10 Ruiz passes + Pock-Chambolle) on the GPU, 2 = additionally the
sliced-ELL bodies of A and A' filled on the GPU from the host's plan (highs_b200/csrc/setup_kernels.cu).  Both must leave
exactly what the host path leaves -- and therefore exactly the reference's scaled data (tests/test_host_prep.py pins the
host passes against the oracle bit for bit): scaled vectors, row-major matrix, SpMV products of the device layouts, and
whole solve trajectories are compared bit for bit on golden LPs (reference order and length-sorted layouts), a ragged LP
and a synthetic LP with a dense column.

The cases run in a child process: a faulting kernel must not poison the CUDA context of the test session.

First hardware run: round 1's driver GPUTEST (both levels bit-exact); plain tests since.  Round 2's default is the
fully device-resident prologue (tests/test_gpu_device_prep.py); these staged variants stay selectable (device_scaling 1/2)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = os.path.join(ROOT, "tests", "device_scaling_child.py")

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("level", [1, 2])
def test_device_setup_matches_host(level):
    try:
        r = subprocess.run([sys.executable, CHILD, str(level)], capture_output=True, text=True, timeout=300, cwd=ROOT)
    except subprocess.TimeoutExpired:
        pytest.fail("device-setup child did not finish in 300 s (killed)")
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, (r.stdout[-2000:], r.stderr[-2000:])
    out = json.loads(lines[-1])
    bad = {k: v for k, v in out.items() if v != "ok"}
    assert not bad and r.returncode == 0, bad
