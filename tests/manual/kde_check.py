"""KDE pair-kernel check: densities vs the float64 oracle on config-2-like data + timings."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from oracle import kde_oracle
from smartstartcontinuous_b200.engine import Engine
eng = Engine(0)
for n, m in ((100_000, 16_384), (1_000_000, 16_384)):
    kw = bench.kde_workload(n=n, m=m)
    j, u, dens, ucb = eng.select_start(kw["all_states"], kw["queries"], kw["values"], kw["n"], kw["volume"],
                                       want_density=True, want_ucb=True)
    ms = 512
    want_j, want_d, want_u = kde_oracle.select_start(kw["all_states"], kw["queries"][:ms], kw["values"][:ms], kw["n"], kw["volume"], 1.0, 2.0)
    if dens is not None:
        rel = np.abs(dens[:ms] / want_d - 1)
        print("n=%d: density rel err max %.2e median %.2e ; ucb rel err max %.2e" % (n, rel.max(), np.median(rel), np.abs(ucb[:ms] / want_u - 1).max()))
    for i in range(5):
        eng.select_start(kw["all_states"], kw["queries"], kw["values"], kw["n"], kw["volume"])
        tm = dict(eng.last_timings())
    print("   phases (us):", {k: round(v * 1e3, 1) for k, v in tm.items()}, "-> pair kernel %.3e evals/s" % (m * (n + 1) / (tm["kde_pairs"] * 1e-3)))
