import sys, os, subprocess, json
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
if len(sys.argv) > 1:
    from highs_b200 import engine
    from highs_b200.lp import synthetic_lp
    lp = synthetic_lp(1000000, 1000000, 8, 12345)
    prob = engine.Problem(lp)
    prob.solve(iter_limit=121)
    ms = prob.bench_pass(200)
    print(sys.argv[1], os.environ.get("B200PDLP_EXP"), os.environ.get("B200PDLP_PREFETCH"), [round(1e3 * v / 200, 1) for v in ms], flush=True)
else:
    for name, env in [("base", {}), ("noreduce", {"B200PDLP_EXP": "1"}), ("pf888", {"B200PDLP_PREFETCH": "888"}),
                      ("pf400", {"B200PDLP_PREFETCH": "400"}), ("pf2000", {"B200PDLP_PREFETCH": "2000"}),
                      ("pf888+noreduce", {"B200PDLP_EXP": "1", "B200PDLP_PREFETCH": "888"})]:
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, __file__, name], env=e)
