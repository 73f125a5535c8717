"""Host-only timing of the engine's prologue (formulate + scale + device layouts) -- no GPU needed.
usage: python tools/host_prologue_bench.py [m n nnz_per_col]     (default: S3, 1e6 x 1e6, 8 per column)
Set B200PDLP_HOST_THREADS to vary the thread count, B200PDLP_TIMING=1 for per-stage laps on stderr."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))

from highs_b200 import engine
from highs_b200.lp import synthetic_lp

m, n, k = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1_000_000, 1_000_000, 8)
lp = synthetic_lp(m, n, k)
L = engine.lib()
clp, keep = engine.make_clp(lp)
for rep in range(int(__import__("os").environ.get("REPS", "3"))):
    h = C.c_void_p()
    t0 = time.perf_counter()
    assert L.b200pdlp_form_create(C.byref(clp), 1, C.byref(h)) == 0
    t1 = time.perf_counter()
    x = np.zeros(n); y = np.zeros(m); ax = np.zeros(m); aty = np.zeros(n); st = np.zeros(12)
    dp = C.POINTER(C.c_double)
    assert L.b200pdlp_form_layout_eval(h, 0, 1, 0, x.ctypes.data_as(dp), y.ctypes.data_as(dp), ax.ctypes.data_as(dp),
                                       aty.ctypes.data_as(dp), st.ctypes.data_as(dp)) == 0
    t2 = time.perf_counter()
    L.b200pdlp_form_destroy(h)
    t3 = time.perf_counter()
    print(f"rep {rep}: formulate+scale {1e3 * (t1 - t0):7.1f} ms   layout {st[11]:7.1f} ms   (host eval {1e3 * (t2 - t1) - st[11]:.0f} ms, destroy {1e3*(t3-t2):.0f} ms)")
