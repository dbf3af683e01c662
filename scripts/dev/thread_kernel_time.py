"""Development helper: the FP32 rollout of the reference's default 1 x 32 model, thread-per-sequence kernel vs the
general CTA kernel (SS_SIMT_GENERAL=1), at K = 5000 (config 1) and larger batches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from smartstartcontinuous_b200.engine import Engine
eng = Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
wls = bench.make_workload_mountaincar(1, 32)
eng.set_model(wls["w"], wls["b"], wls["norm"])
eng.set_plan(wls["plan"]["desired_states"], wls["plan"]["distances_left"], wls["plan"]["radii"])
for K, H in ((5000, 4), (50000, 4), (500000, 4), (50000, 20)):
    for general in ("0", "1"):
        if general == "1": os.environ["SS_SIMT_GENERAL"] = "1"
        else: os.environ.pop("SS_SIMT_GENERAL", None)
        def step(i):
            eng.plan(wls["state"], 0, K=K, H=H, seed=500 + i, act_low=wls["low"], act_high=wls["high"], penalty_mode="reference", precision="fp32")
        for i in range(5): step(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20): step(i)
        e1.record(); torch.cuda.synchronize()
        print("K=%d H=%d %s: decision %.1f us (%s)" % (K, H, "general CTA kernel" if general == "1" else "thread kernel    ", e0.elapsed_time(e1) * 50, eng.last_rollout_kernel()))
os.environ.pop("SS_SIMT_GENERAL", None)
