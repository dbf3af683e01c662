"""Development helper: SS_TC_TRACE clock stamps of the quad kernel at config 3 (profiles/r02_quad_trace.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
from smartstartcontinuous_b200.engine import Engine
eng = Engine(0)
wls = bench.make_workload_mountaincar(2, 500)
eng.set_model(wls["w"], wls["b"], wls["norm"])
eng.set_plan(wls["plan"]["desired_states"], wls["plan"]["distances_left"], wls["plan"]["radii"])
for i in range(3):
    eng.plan(wls["state"], 0, K=4096, H=20, seed=500 + i, act_low=wls["low"], act_high=wls["high"], penalty_mode="reference", precision="bf16_tc")
