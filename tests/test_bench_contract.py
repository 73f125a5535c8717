"""bench.py contract pieces that can run without a GPU: the reference arm's JSON line, the byte model."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_algorithmic_bytes_model():
    sys.path.insert(0, ROOT)
    import bench
    B = bench.algorithmic_bytes(1_000_000, 1_000_000, 8_000_000)
    assert B["ax"] == 12 * 8_000_000 + 4 * 1_000_001 + 16_000_000          # 116.0 MB
    assert abs(B["ax"] / 1e6 - 116.0) < 0.01 and B["aty"] == B["ax"]
    assert B["iter"] == B["ax"] + B["aty"] + 72 * 1_000_000 + 56 * 1_000_000  # 360 MB
    assert B["k1"] == 64_000_000 and B["k2"] == B["ax"] + 48_000_000 and B["k3"] == B["aty"] + 24_000_000


def test_reference_arm_json_line(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "S2",
                          "--steps", "40", "--warmup", "3"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pdhg_iterations_per_sec" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "S2" in d["config"]["workload"]
