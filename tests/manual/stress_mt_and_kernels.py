"""Manual stress run (GPU): random batch shapes through Engine.plan with numpy's stream generated on the device
vs the same decision on the host draw -- bit-identical scores, sequences and generator states every time, across the
FP32 / pair / quad rollout kernels, single-CTA / shared-jump / per-SM MT19937 layouts, interleaved with selections.
  python tests/manual/stress_mt_and_kernels.py [seconds]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from smartstartcontinuous_b200.engine import Engine

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
eng = Engine(0)
rng = np.random.default_rng(0)
kw = bench.kde_workload(n=20000, m=2048)
models = {"2x500": bench.make_workload_mountaincar(2, 500), "1x32": bench.make_workload_mountaincar(1, 32)}
t0, n, kernels = time.time(), 0, {}
cur = None
while time.time() - t0 < budget:
    name = "2x500" if rng.random() < 0.7 else "1x32"
    wl = models[name]
    if cur != name:
        eng.set_model(wl["w"], wl["b"], wl["norm"])
        eng.set_plan(wl["plan"]["desired_states"], wl["plan"]["distances_left"], wl["plan"]["radii"])
        cur = name
    K = int(rng.choice([300, 1000, 4096, 5000, 12000, 40000, 100000]))
    H = int(rng.choice([3, 8, 20, 33]))
    mode = "reference" if rng.random() < 0.6 else "per_sample"
    rs = np.random.RandomState(int(rng.integers(1 << 30)))
    rs.randint(0, 2 ** 31, size=int(rng.integers(0, 700)))              # any word position, odd ones included
    st = rs.get_state()
    acts = rs.uniform(np.array(wl["low"], dtype=np.float64), np.array(wl["high"], dtype=np.float64), (K, H, 1))
    after = rs.get_state()
    prec = "auto" if name == "2x500" else "fp32"
    want = eng.plan(wl["state"], 0, actions=acts, penalty_mode=mode, precision=prec, want_scores=True)
    got = eng.plan(wl["state"], 0, K=K, H=H, act_low=wl["low"], act_high=wl["high"], rng_state=st, penalty_mode=mode,
                   precision=prec, want_scores=True)
    kernels[eng.last_rollout_kernel()] = kernels.get(eng.last_rollout_kernel(), 0) + 1
    assert got["best_k"] == want["best_k"] and np.array_equal(got["scores"], want["scores"]), (name, K, H, mode)
    assert np.array_equal(got["best_sequence"], acts[got["best_k"]]) and np.array_equal(got["best_path"], want["best_path"])
    assert got["rng_state"][2] == after[2] and np.array_equal(got["rng_state"][1], after[1]), (name, K, H)
    if n % 5 == 0:
        a = eng.select_start(kw["all_states"], kw["queries"], kw["values"], kw["n"], kw["volume"])
        b = eng.select_start(kw["all_states"], kw["queries"], kw["values"], kw["n"], kw["volume"])
        assert a[0] == b[0] and a[1] == b[1]
    n += 1
# ---- results through the tagged host slots: back-to-back small decisions without any stream synchronisation in
# between (no scores requested); a stale or torn package would differ from the decision recomputed with the scores
t1, m = time.time(), 0
wl = models["2x500"]
eng.set_model(wl["w"], wl["b"], wl["norm"])
eng.set_plan(wl["plan"]["desired_states"], wl["plan"]["distances_left"], wl["plan"]["radii"])
while time.time() - t1 < budget / 3:
    K = int(rng.choice([64, 300, 1000, 4096]))
    H = int(rng.choice([3, 8, 20, 50]))
    seeds = [int(s) for s in rng.integers(1 << 40, size=40)]
    fast = [eng.plan(wl["state"], 0, K=K, H=H, seed=s, act_low=wl["low"], act_high=wl["high"]) for s in seeds]
    sel = [eng.select_start(kw["all_states"], kw["queries"][:256 + 8 * i], kw["values"][:256 + 8 * i], kw["n"], kw["volume"])
           for i in range(8)]
    for s, f in zip(seeds[::8], fast[::8]):
        full = eng.plan(wl["state"], 0, K=K, H=H, seed=s, act_low=wl["low"], act_high=wl["high"], want_scores=True)
        assert f["best_k"] == full["best_k"] == int(np.argmax(full["scores"])) and f["best_score"] == full["best_score"]
        assert np.array_equal(f["best_sequence"], full["best_sequence"]) and np.array_equal(f["best_path"], full["best_path"])
    for i, r in enumerate(sel[::4]):
        again = eng.select_start(kw["all_states"], kw["queries"][:256 + 32 * i], kw["values"][:256 + 32 * i], kw["n"], kw["volume"])
        assert r[0] == again[0] and r[1] == again[1]
    m += len(seeds) + len(sel)
print("stress ok: %d decisions in %.0f s, kernels %s; %d back-to-back results through the host slots" %
      (n, time.time() - t0, kernels, m))
