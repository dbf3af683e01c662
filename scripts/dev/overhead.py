"""Host-side overhead of one decision: tiny batches, CUDA-event time around the Python call."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from smartstartcontinuous_b200.distributed import ShardedPlanner
from smartstartcontinuous_b200.engine import Engine
eng = Engine(0)
stream = torch.cuda.current_stream(); eng.set_stream(stream.cuda_stream)
wl = bench.make_workload()
eng.set_model(wl["w"], wl["b"], wl["norm"]); eng.set_plan(wl["plan"]["desired_states"], wl["plan"]["distances_left"], wl["plan"]["radii"])
planner = ShardedPlanner(eng, device=torch.device("cuda", 0))
for K, H in ((256, 2), (131072, 50)):
    for name, fn in (("planner.plan", lambda i: planner.plan(wl["state"], 0, K=K, H=H, seed=i, act_low=wl["low"], act_high=wl["high"], penalty_mode="reference", precision="bf16_tc", want_path=True)),
                     ("engine.plan", lambda i: eng.plan(wl["state"], 0, K=K, H=H, seed=i, act_low=wl["low"], act_high=wl["high"], penalty_mode="reference", precision="bf16_tc", want_path=True))):
        for i in range(5): fn(i)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        t0 = time.perf_counter()
        for i in range(20):
            ev[i][0].record(stream); fn(i); ev[i][1].record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 20
        ms = np.mean([a.elapsed_time(b) for a, b in ev])
        print("K=%d H=%d %-13s event %.3f ms  wall %.3f ms  phases %s" % (K, H, name, ms, wall * 1e3, [(n, round(v, 3)) for n, v in eng.last_timings()]))
