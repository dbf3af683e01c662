"""Quick device-side timings of the two stages (development helper, not the bench)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smartstartcontinuous_b200 import synthetic as syn
from smartstartcontinuous_b200.engine import Engine
from smartstartcontinuous_b200.nnd_mb_agent import plan_from_path

eng = Engine(0)
eng.set_timing(True)
print(eng.device_info())
for n, m in ((100_000, 16_384), (1_000_000, 16_384), (100_000, 2_000)):
    all_states, s2, _ = syn.pendulum_buffer(n, seed=0)
    rng = np.random.default_rng(0)
    q = s2[rng.choice(n, m, replace=False)]
    vals = syn.critic_like_values(q)
    for it in range(4):
        t0 = time.time()
        eng.select_start(all_states, q, vals, n, 1e-3)
        wall = time.time() - t0
        tm = eng.last_timings()
    pairs = dict(tm).get("kde_pairs", 0)
    print("KDE n=%d m=%d wall=%.2fms phases=%s -> %.3e evals/s (pairs kernel)" %
          (n, m, wall * 1e3, tm, (n + 1) * m / (pairs * 1e-3)))

rng = np.random.default_rng(3)
for (d, env, K, H) in ((2, "mc", 4096, 20), (3, "pend", 32768, 20)):
    if env == "mc":
        roll = [syn.mountaincar_rollout(rng, 200) for _ in range(8)]
        st = [r[0] for r in roll]; ac = [r[1] for r in roll]; lo, hi = [-1.0], [1.0]
    else:
        obs, act = syn.pendulum_rollouts(rng, 8, 200)
        st = list(obs); ac = list(act); lo, hi = [-2.0], [2.0]
    norm = syn.normalisation_stats(np.concatenate(st), np.concatenate([np.concatenate([a, a[-1:]]) for a in ac]))
    w, b = syn.xavier_mlp(rng, d, 1, 2, 500, scale=0.5)
    plan = plan_from_path(list(st[0][:60]), mean_per_stepsize=1, std_per_stepsize=1,
                          stepsizes_in_waypoint_radii=1, path_shortcutting=True, theta=1, steps_per_waypoint=1)
    eng.set_model(w, b, norm)
    eng.set_plan(plan["desired_states"], plan["distances_left"], plan["radii"])
    precs = ["fp32"] + (["bf16_tc"] if eng.tc_supported() else [])
    for prec in precs:
        for mode in ("per_sample", "reference"):
            for it in range(3):
                t0 = time.time()
                res = eng.plan(st[0][0], 0, K=K, H=H, seed=1, act_low=lo, act_high=hi, penalty_mode=mode,
                               precision=prec, want_path=False)
                wall = time.time() - t0
                tm = eng.last_timings()
            roll_ms = dict(tm).get("mpc_rollout", 0)
            print("MPC %s K=%d H=%d %s %s wall=%.2fms phases=%s -> %.3e rollout-steps/s (rollout kernel)" %
                  (env, K, H, prec, mode, wall * 1e3, tm, K * H / (roll_ms * 1e-3)))
