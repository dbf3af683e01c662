"""Development helper: decision and rollout-kernel time of small batches (config 3 and neighbours), pair vs quad kernel (SS_TC_QUAD)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from smartstartcontinuous_b200.engine import Engine
eng = Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
wls = bench.make_workload_mountaincar(2, 500)
eng.set_model(wls["w"], wls["b"], wls["norm"])
eng.set_plan(wls["plan"]["desired_states"], wls["plan"]["distances_left"], wls["plan"]["radii"])
for K, H in ((4096, 20), (2048, 20), (512, 20), (4736, 20), (8192, 20)):
    for quad in ("0", "1"):
        os.environ["SS_TC_QUAD"] = quad
        ks = []
        def step(i):
            eng.plan(wls["state"], 0, K=K, H=H, seed=500 + i, act_low=wls["low"], act_high=wls["high"], penalty_mode="reference", precision="bf16_tc")
            ks.append(dict(eng.last_timings()).get("mpc_rollout", 0.0))
        for i in range(5): step(i)
        ks.clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20): step(i)
        e1.record(); torch.cuda.synchronize()
        print("K=%d H=%d quad=%s: decision %.1f us, rollout kernel %.1f us" % (K, H, quad, e0.elapsed_time(e1) * 50, 1e3 * np.mean(ks[1:])))
