import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from conftest import golden_model, load_golden
from oracle import mpc_oracle
from smartstartcontinuous_b200 import synthetic as syn
from smartstartcontinuous_b200.engine import Engine
from smartstartcontinuous_b200.nnd_mb_agent import plan_from_path

eng = Engine(0)
g = load_golden("mpc_mountaincar_L2.npz")
w, b, norm = golden_model(g)
eng.set_model(w, b, norm)
eng.set_plan(g["out_desired_states"], g["out_distances_left"], g["out_radii"])
print("tc supported:", eng.tc_supported())
for prec in ("fp32", "bf16_tc"):
    st = eng.forward_sim(g["in_start_state"], g["in_actions"], precision=prec)
    err = np.abs(st - g["out_states"])
    print(prec, "state max abs err per dim", err.max(axis=(0, 1)), "state scale", np.abs(g["out_states"]).max(axis=(0, 1)),
          "max |delta| per dim", np.abs(np.diff(g["out_states"], axis=0)).max(axis=(0, 1)))
    for mode in ("reference", "per_sample"):
        res = eng.plan(g["in_start_state"], int(g["in_wp_index"]), actions=g["in_actions"], penalty_mode=mode,
                       precision=prec, want_scores=True)
        want = g["out_scores"] if mode == "reference" else mpc_oracle.score_add_delta(
            g["out_states"], g["out_desired_states"], g["out_distances_left"], g["out_radii"], int(g["in_wp_index"]),
            .75, .5, penalty_mode=1)
        e = np.abs(res["scores"] - want)
        print("  ", mode, "score err max %.3g median %.3g  best %d want %d" % (e.max(), np.median(e), res["best_k"], int(np.argmax(want))))

# big shapes
rng = np.random.default_rng(3)
for (d, env, K, H) in ((2, "mc", 4096, 20), (3, "pend", 131072, 50)):
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
    out = {}
    for prec in ("fp32", "bf16_tc"):
        if prec == "fp32" and K > 40000:
            continue
        for mode in ("per_sample", "reference"):
            for it in range(3):
                res = eng.plan(st[0][0], 0, K=K, H=H, seed=1, act_low=lo, act_high=hi, penalty_mode=mode,
                               precision=prec, want_scores=True, want_path=False)
                tm = eng.last_timings()
            out[(prec, mode)] = res["scores"]
            roll_ms = dict(tm).get("mpc_rollout", 0)
            print("MPC %s K=%d H=%d %s %s phases=%s -> %.3e rollout-steps/s  (%.1f TFLOP/s)" %
                  (env, K, H, prec, mode, [(n, round(v, 3)) for n, v in tm], K * H / (roll_ms * 1e-3),
                   K * H / (roll_ms * 1e-3) * 507e3 / 1e12))
    if ("fp32", "per_sample") in out:
        e = np.abs(out[("fp32", "per_sample")] - out[("bf16_tc", "per_sample")])
        s = np.abs(out[("fp32", "per_sample")])
        print("   tc vs fp32 score diff: max %.3g median %.3g (score scale median %.3g)" % (e.max(), np.median(e), np.median(s)))
