"""Measured error of the tcgen05 rollout against the float64 oracle at the configurations the bench
runs: the two golden cases, BASELINE config 3 (MountainCar, K=4096, H=20, 2x500) and config 4's
per-GPU shard (Pendulum, K=131072, H=50, 2x500), reference penalty.  Prints one line per case and
writes gpurun_out/tc_error_report.json; DESIGN.md section 3.4 and TC_SCORE_ATOL in
tests/test_gpu_mpc.py quote these numbers.

    python tests/manual/tc_error_report.py [--small]      (--small: K=16384 for the H=50 case)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden_model, load_golden  # noqa: E402
from oracle import mpc_oracle, philox  # noqa: E402
from smartstartcontinuous_b200 import synthetic as syn  # noqa: E402
from smartstartcontinuous_b200.engine import Engine  # noqa: E402
from smartstartcontinuous_b200.nnd_mb_agent import plan_from_path  # noqa: E402


def stats(name, eng, start, wp, acts, w, b, norm, plan, want=None, want_states=None):
    t0 = time.time()
    if want is None:
        o = mpc_oracle.plan(start, acts, w, b, norm, plan["desired_states"], plan["distances_left"], plan["radii"],
                            wp, .75, .5, penalty_mode=0)
        want, want_states = o["scores"], o["states"]
    t_or = time.time() - t0
    out = {"case": name, "K": int(acts.shape[0]), "H": int(acts.shape[1]), "oracle_s": round(t_or, 2)}
    order = np.argsort(-want, kind="stable")
    out["oracle_top2_gap"] = float(want[order[0]] - want[order[1]])
    out["oracle_score_range"] = [float(want.min()), float(want.max())]
    scale = np.abs(want_states).max(axis=(0, 1))
    for prec in ("fp32", "bf16_tc"):
        res = eng.plan(start, wp, actions=acts, penalty_mode="reference", precision=prec, want_scores=True)
        st = eng.get_states()
        serr = np.abs(st - want_states).max(axis=(0, 1)) / scale
        e = np.abs(res["scores"] - want)
        out[prec] = {"state_err_over_scale_max": float(serr.max()),
                     "state_err_last_step": float((np.abs(st[-1] - want_states[-1]).max(axis=0) / scale).max()),
                     "score_err_max": float(e.max()), "score_err_p999": float(np.quantile(e, .999)),
                     "score_err_p99": float(np.quantile(e, .99)), "score_err_median": float(np.median(e)),
                     "frac_gt_1e-2": float((e > 1e-2).mean()), "frac_gt_5e-2": float((e > 5e-2).mean()),
                     "best_k": int(res["best_k"]), "oracle_best_k": int(order[0]),
                     "oracle_score_of_chosen_minus_top": float(want[res["best_k"]] - want[order[0]])}
    print(json.dumps(out), flush=True)
    return out


def main():
    small = "--small" in sys.argv
    eng = Engine(0)
    rows = []
    for name in ("mpc_mountaincar_L2.npz", "mpc_pendulum_2x500.npz"):
        g = load_golden(name)
        w, b, norm = golden_model(g)
        eng.set_model(w, b, norm)
        plan = dict(desired_states=g["out_desired_states"], distances_left=g["out_distances_left"], radii=g["out_radii"])
        eng.set_plan(plan["desired_states"], plan["distances_left"], plan["radii"])
        rows.append(stats("golden " + name, eng, g["in_start_state"], int(g["in_wp_index"]), g["in_actions"], w, b, norm,
                          plan, want=g["out_scores"], want_states=g["out_states"]))
    # BASELINE config 3
    rng = np.random.default_rng(3)
    roll = [syn.mountaincar_rollout(rng, 200) for _ in range(8)]
    norm = syn.normalisation_stats(np.concatenate([r[0] for r in roll]),
                                   np.concatenate([np.concatenate([r[1], r[1][-1:]]) for r in roll]))
    w, b = syn.xavier_mlp(rng, 2, 1, 2, 500, scale=0.5)
    plan = plan_from_path(list(roll[0][0][:60]), mean_per_stepsize=1, std_per_stepsize=1,
                          stepsizes_in_waypoint_radii=1, path_shortcutting=True, theta=1, steps_per_waypoint=1)
    eng.set_model(w, b, norm)
    eng.set_plan(plan["desired_states"], plan["distances_left"], plan["radii"])
    acts = np.random.RandomState(1).uniform(-1, 1, (4096, 20, 1))
    rows.append(stats("config3 mountaincar K=4096 H=20", eng, roll[0][0][0], 0, acts, w, b, norm, plan))
    # BASELINE config 4 shard (the bench workload)
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.make_workload()
    eng.set_model(wl["w"], wl["b"], wl["norm"])
    eng.set_plan(wl["plan"]["desired_states"], wl["plan"]["distances_left"], wl["plan"]["radii"])
    for K in ((16384,) if small else (16384, 131072)):
        acts = philox.sample_actions(K, 50, 1, 1001, wl["low"], wl["high"])
        rows.append(stats("config4 shard pendulum K=%d H=50" % K, eng, wl["state"], 0, acts, wl["w"], wl["b"],
                          wl["norm"], wl["plan"]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tc_error_report.json"), "w") as fh:
        json.dump(rows, fh, indent=1)


if __name__ == "__main__":
    main()
