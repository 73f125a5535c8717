import sys, time, os
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
os.environ["B200PDLP_TIMING"] = "1"
from highs_b200 import engine
from highs_b200.lp import synthetic_lp
lp = synthetic_lp(1000000, 1000000, 8, 12345)
for i in range(2):
    t = time.time(); r = engine.solve(lp, iter_limit=2001); dt = time.time() - t
    print("e2e wall", dt, "setup", r["setup_seconds"], "solve", r["solve_seconds"], flush=True)
