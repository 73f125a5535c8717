""""the same code path must run with G logical shards on 1 device".  G problems (rank g of G) are
created in ONE process on ONE GPU, wired to each other with b200pdlp_p2p_link_local (plain pointers instead of CUDA IPC,
no NCCL), and solved by G host threads: the row/column partition, the fused peer-memory kernels, both flag barriers,
the device-side step rule, the speculative check and the NCCL-free solution assembly all run as they do on G GPUs.
This is what lets a one-GPU box exercise the multi-GPU path (tests/test_gpu_multi.py needs two devices).

Each case runs in a child process with a hard timeout (a stuck barrier must not take the session along).

STATUS: the local-link entry point and the NCCL-free assembly were written after this round's GPU budget was spent;
they have been compiled and reviewed but not yet run on hardware, hence xfail(strict=False): a pass shows up as XPASS,
a failure does not turn the suite red.  The marker goes away after the first hardware run."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = os.path.join(ROOT, "tests", "logical_shards_child.py")

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="logical-shard path not yet run on hardware (written after the GPU budget was spent)")]


_state = {"broken": False}   # after one failing case the others are not attempted (bounds the time a broken path can cost)


@pytest.mark.parametrize("world,case,mode", [(2, "synthetic", "threads"), (3, "adlittle", "threads"), (4, "dense", "threads"),
                                             (2, "synthetic", "c_entry")])
def test_logical_shards(world, case, mode):
    if _state["broken"]:
        pytest.xfail("an earlier logical-shard case failed; not attempted")
    _state["broken"] = True
    try:
        r = subprocess.run([sys.executable, CHILD, str(world), case, mode], capture_output=True, text=True, timeout=240, cwd=ROOT)
    except subprocess.TimeoutExpired:
        pytest.fail("logical-shard solve did not finish in 240 s (child killed)")
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, (r.stdout[-2000:], r.stderr[-2000:])
    out = json.loads(lines[-1])
    assert out["ok"] and out["ranks_identical"], out
    assert out["term"][0] == out["ref_term"]
    _state["broken"] = False
