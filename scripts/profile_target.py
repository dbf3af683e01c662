"""Short, fixed workload for ncu: 2 MPC decisions on the BASELINE config-4 shard + 2 KDE
selections on config 2 (same code path as bench.py, no timing)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from smartstartcontinuous_b200.engine import Engine

eng = Engine(0)
eng.set_timing(True)
if os.environ.get("SS_PROFILE_CFG") == "c3":
    # BASELINE config 3: MountainCar, K=4096, H=20, MLP 2x500 -- the small-K decision
    wl = bench.make_workload_mountaincar(2, 500)
    eng.set_model(wl["w"], wl["b"], wl["norm"])
    eng.set_plan(wl["plan"]["desired_states"], wl["plan"]["distances_left"], wl["plan"]["radii"])
    for i in range(6):
        r = eng.plan(wl["state"], 0, K=bench.C3_K, H=bench.C3_H, seed=i, act_low=wl["low"], act_high=wl["high"],
                     penalty_mode="reference", precision="bf16_tc")
    print("mpc c3", r["best_k"], r["best_score"], eng.last_timings())
    sys.exit(0)
if os.environ.get("SS_PROFILE_CFG") == "c1":
    # BASELINE config 1 (the reference's own example): MountainCar, K=5000, H=4, MLP 1x32 -- the FP32 kernel
    wl = bench.make_workload_mountaincar(1, 32)
    eng.set_model(wl["w"], wl["b"], wl["norm"])
    eng.set_plan(wl["plan"]["desired_states"], wl["plan"]["distances_left"], wl["plan"]["radii"])
    for i in range(6):
        r = eng.plan(wl["state"], 0, K=bench.C1_K, H=bench.C1_H, seed=i, act_low=wl["low"], act_high=wl["high"],
                     penalty_mode="reference", precision="auto")
    print("mpc c1", r["best_k"], r["best_score"], eng.last_timings())
    sys.exit(0)
if os.environ.get("SS_PROFILE_CFG") == "mt":
    # numpy's MT19937 stream on the device at the bench shape (K = 131072, H = 50) and at config 3's
    rs = np.random.RandomState(1)
    rs.random_sample(100)
    for n in (bench.K_PER_GPU * bench.HORIZON, bench.C3_K * bench.C3_H):
        for i in range(2):
            eng.mt19937_uniform(rs.get_state(), n, [-2.0], [2.0])
            st = eng.mt19937_state()
    print("mt19937 pos", st[2])
    sys.exit(0)
if os.environ.get("SS_PROFILE_CFG") == "sel":
    # the selection call with the candidates' values computed on the device (row f4) from the device mirror
    import random
    random.seed(0)
    kw = bench.kde_workload()
    ss = bench.make_smart_start(eng, kw, 16384, device_values=True)
    for i in range(3):
        path = ss.get_smart_start_path()
    print("selection", ss.last_selection, len(path), eng.last_timings())
    sys.exit(0)
if os.environ.get("SS_PROFILE_CFG") == "dyn":
    # row f1: a few Adam steps of the device trainer at the BASELINE shapes (2x500, batch 512)
    from oracle import dyn_train_oracle as dto
    tw = bench.dyn_train_workload()
    eng.set_model(tw["w"], tw["b"], tw["norm"])
    eng.dyn_set_data(0, tw["X_old"], tw["Z_old"])
    eng.dyn_set_data(1, tw["X_new"], tw["Z_new"])
    np.random.seed(0)
    io, inw = dto.epoch_batches(len(tw["X_old"]), len(tw["X_new"]), tw["batch"], tw["frac"])
    losses = eng.dyn_train_batches(io[:6], inw[:6], tw["lr"])
    eng.dyn_commit()
    print("dyn losses", losses)
    sys.exit(0)
wl = bench.make_workload()
eng.set_model(wl["w"], wl["b"], wl["norm"])
eng.set_plan(wl["plan"]["desired_states"], wl["plan"]["distances_left"], wl["plan"]["radii"])
prec = "bf16_tc" if eng.tc_supported() else "fp32"
K = int(os.environ.get("SS_PROFILE_K", bench.K_PER_GPU))
for i in range(2):
    r = eng.plan(wl["state"], 0, K=K, H=bench.HORIZON, seed=i, act_low=wl["low"], act_high=wl["high"],
                 penalty_mode="reference", precision=prec)
print("mpc", r["best_k"], r["best_score"], eng.last_timings())
kw = bench.kde_workload()
for i in range(2):
    j, u, _, _ = eng.select_start(kw["all_states"], kw["queries"], kw["values"], kw["n"], kw["volume"])
print("kde", j, u, eng.last_timings())
