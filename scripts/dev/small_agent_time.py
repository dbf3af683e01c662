"""Development helper: NND_MB_agent.get_best_sim_actions at configs 3 and 1 -- host draw / device MT19937 / Philox -- and a cProfile of the host side."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from smartstartcontinuous_b200.engine import Engine
eng = Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
eng.set_timing(False)
for name, (L, h, K, H) in (("config3", (2, 500, 4096, 20)), ("config1", (1, 32, 5000, 4))):
    wls = bench.make_workload_mountaincar(L, h)
    for host_rng in (True, False):
        np.random.seed(1)
        ag = bench.make_nav_agent(eng, wls, K, H, device_sampling=False, host_rng=host_rng)
        for _ in range(5): ag.get_best_sim_actions(wls["state"])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
        for _ in range(50): ag.get_best_sim_actions(wls["state"])
        e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
        print("%s host_rng=%s: %.1f us per decision (events), %.1f us wall" % (name, host_rng, e0.elapsed_time(e1) * 20, (t1 - t0) * 2e4))
    ag = bench.make_nav_agent(eng, wls, K, H, device_sampling=True)
    for _ in range(5): ag.get_best_sim_actions(wls["state"])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): ag.get_best_sim_actions(wls["state"])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("%s philox: %.1f us wall" % (name, (t1 - t0) * 2e4))
import cProfile, pstats
ag = bench.make_nav_agent(eng, wls, 5000, 4, device_sampling=False, host_rng=False)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): ag.get_best_sim_actions(wls["state"])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
