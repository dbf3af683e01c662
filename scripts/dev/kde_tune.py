import os, sys, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from smartstartcontinuous_b200.engine import Engine
eng = Engine(0)
kw = bench.kde_workload()
for per_sm in (12, 14, 16, 20, 24, 28, 32, 48):
    os.environ["SS_KDE_CTAS_PER_SM"] = str(per_sm)
    best = 1e9
    for i in range(6):
        eng.select_start(kw["all_states"], kw["queries"], kw["values"], kw["n"], kw["volume"])
        tm = dict(eng.last_timings()); best = min(best, tm["kde_pairs"])
    print("ctas/sm", per_sm, "pairs %.1f us -> %.3e evals/s" % (best * 1e3, 16384 * 100001 / (best * 1e-3)), {k: round(v*1e3,1) for k,v in tm.items()})
