"""Development helper: wall time of SmartStartContinuous.get_smart_start_path with the candidates' values
computed on the device (bench leg kde.e2e_device_values), with a cProfile of the host side."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import random
import numpy as np, torch
import bench
from smartstartcontinuous_b200.engine import Engine
eng = Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
eng.set_timing(False)
kw = bench.kde_workload()
for n_ss in (16384, 2000):
    random.seed(0)
    ss = bench.make_smart_start(eng, kw, n_ss, device_values=True)
    for _ in range(3): ss.get_smart_start_path()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): ss.get_smart_start_path()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("n_ss=%d: %.1f us per get_smart_start_path" % (n_ss, (t1 - t0) * 5e4))
    if n_ss == 16384:
        import cProfile, pstats
        pr = cProfile.Profile(); pr.enable()
        for _ in range(50): ss.get_smart_start_path()
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
# phases of the mirror selection (per-phase events on)
eng.set_timing(True)
random.seed(0)
ss = bench.make_smart_start(eng, kw, 16384, device_values=True)
for _ in range(3): ss.get_smart_start_path()
print("phases (ms):", [(k, round(v, 4)) for k, v in eng.last_timings()])
rb = ss.replay_buffer
idx = rb.get_possible_smart_start_indices(16384)
ring = rb.state_ring()
eng.set_timing(False)
for _ in range(3): eng.select_start_mirror(ring, idx, None, len(rb), 1e-3, 1.0, 2.0)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): eng.select_start_mirror(ring, idx, None, len(rb), 1e-3, 1.0, 2.0)
t1 = time.perf_counter()
print("select_start_mirror alone: %.1f us" % ((t1 - t0) * 2e4))
