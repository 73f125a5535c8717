import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden():
    return json.load(open(os.path.join(GOLDEN, "golden.json")))["cases"]


def golden_lp(case):
    from highs_b200.lp import read_b2lp, synthetic_lp
    if "synthetic" in case:
        m, n, k, seed = case["synthetic"]
        return synthetic_lp(m, n, k, seed)
    return read_b2lp(os.path.join(GOLDEN, case["name"] + ".b2lp"))


def golden_solution(case):
    p = os.path.join(GOLDEN, f"{case['name']}.{case['tag']}.npz")
    return dict(np.load(p)) if os.path.exists(p) else None


def case_id(case):
    return f"{case['name']}-{case['tag']}"


def options_to_params(options):
    """HighsOptions (as in golden.json) -> engine/oracle parameters (CupdlpWrapper.cpp:642-717)."""
    p = {}
    tol = options.get("kkt_tolerance")
    if tol is not None:
        p.update(tol_primal=tol, tol_dual=tol, tol_gap=tol)
    if "pdlp_iteration_limit" in options:
        p["iter_limit"] = options["pdlp_iteration_limit"]
    if options.get("pdlp_cupdlpc_restart_method", 1) == 0:
        p["restart"] = 0
    return p


# cuPDLP termination code -> HighsModelStatus code before lpKktCheck (CupdlpWrapper.cpp:220-251)
def term_to_model_status(term_code, iters, iter_limit=2147483647):
    return {0: 7, 1: 8, 2: 10, 3: 9}.get(term_code, 14 if iters >= iter_limit - 1 else 13)


@pytest.fixture(scope="session")
def engine_lib():
    from highs_b200 import build, engine
    build.build()   # compares mtimes: rebuilds only what changed, so the tests never run against a stale library
    return engine.lib()


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.lib()
    return binding
