"""Measured precision of the tcgen05 rollout against the reference's own float64 outputs
(tests/golden/mpc_pendulum_2x500.npz: Pendulum, 2x500, K=600, H=20)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from conftest import golden_model, load_golden
from smartstartcontinuous_b200.engine import Engine
eng = Engine(0)
g = load_golden("mpc_pendulum_2x500.npz")
w, b, norm = golden_model(g)
eng.set_model(w, b, norm)
eng.set_plan(g["out_desired_states"], g["out_distances_left"], g["out_radii"])
scale = np.abs(g["out_states"]).max(axis=(0, 1))
for prec in ("fp32", "bf16_tc"):
    st = eng.forward_sim(g["in_start_state"], g["in_actions"], precision=prec)
    err = np.abs(st - g["out_states"])
    res = eng.plan(g["in_start_state"], int(g["in_wp_index"]), actions=g["in_actions"], penalty_mode="reference",
                   precision=prec, want_scores=True)
    e = np.abs(res["scores"] - g["out_scores"])
    print("%-8s state err / scale: max %s (final step %s) | score err: max %.3g median %.3g (|score| median %.3g) | best %d want %d"
          % (prec, (err.max(axis=(0, 1)) / scale).round(6), (err[-1].max(axis=0) / scale).round(6), e.max(), np.median(e),
             np.median(np.abs(g["out_scores"])), res["best_k"], int(g["out_best_k"])))
