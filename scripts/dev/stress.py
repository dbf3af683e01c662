"""Determinism stress: the same decision (same seed) must give bit-identical scores every time --
a race in the tcgen05 pipelines (TMEM in-place conversion, slot hand-over, ring reuse) would show up
as an occasional mismatch.  Also the KDE (tcgen05 pair kernel): identical densities every time."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from smartstartcontinuous_b200.engine import Engine
eng = Engine(0)
wl = bench.make_workload()
eng.set_model(wl["w"], wl["b"], wl["norm"]); eng.set_plan(wl["plan"]["desired_states"], wl["plan"]["distances_left"], wl["plan"]["radii"])
bad = 0
for K, H, reps in ((131072, 50, 40), (40000, 13, 100), (700, 50, 200)):
    for mode in ("reference", "per_sample"):
        ref = None
        for i in range(reps):
            r = eng.plan(wl["state"], 0, K=K, H=H, seed=5, act_low=wl["low"], act_high=wl["high"], penalty_mode=mode,
                         precision="bf16_tc", want_scores=True)
            if ref is None:
                ref = r
            elif not (np.array_equal(r["scores"], ref["scores"]) and r["best_k"] == ref["best_k"] and
                      np.array_equal(r["best_path"], ref["best_path"])):
                bad += 1
                print("MISMATCH K=%d H=%d %s rep %d: %d scores differ" % (K, H, mode, i, int((r["scores"] != ref["scores"]).sum())))
        print("K=%d H=%d %s: %d repetitions identical=%s" % (K, H, mode, reps, bad == 0))
kw = bench.kde_workload()
ref = None
for i in range(100):
    r = eng.select_start(kw["all_states"], kw["queries"], kw["values"], kw["n"], kw["volume"], want_density=True)
    if ref is None:
        ref = r
    elif not (np.array_equal(r[2], ref[2]) and r[0] == ref[0]):
        bad += 1
        print("KDE MISMATCH rep", i, int((r[2] != ref[2]).sum()))
print("KDE: 100 repetitions identical=%s" % (bad == 0))
print("STRESS", "PASS" if bad == 0 else "FAIL")
