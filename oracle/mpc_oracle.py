"""CPU oracle for SmartStart stage 2: NND_MB random-shooting MPC (float64 numpy).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates, in vectorised float64 numpy:
  * Dyn_Model.do_forward_sim, many_in_parallel branch  (dynamics_model.py:204-240)
  * the MLP of feedforward_network.py:3-23 (what sess.run evaluates; TF itself is a
    third-party dependency that cannot be installed here: tensorflow 1.5, Pipfile.lock:850)
  * NND_MB_agent.generate_scores_add_delta            (NND_MB_agent.py:566-628)
    with move_to_next (:491-496), the elliptical distance (numerical.py:116-124) and
    dist_line_seg_to_point / projection_of_a_onto_b (numerical.py:74-98) *including* the
    axis-less np.sum that makes the projection coefficient global over all K samples.
  * selection                                           (NND_MB_agent.py:516-518)

Pinned against the reference's own functions through tests/golden/mpc_*.npz
(produced by oracle/make_golden.py from the reference's get_best_sim_actions; checked in
tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np

PENALTY_REFERENCE = 0    # reference-exact: one projection coefficient per time step (global)
PENALTY_PER_SAMPLE = 1   # per-sample projection (the evidently intended maths)


def mlp_forward(x, weights, biases):
    """feedforward_network.py:12-23: hidden layers Linear+ReLU, output layer Linear."""
    h = np.asarray(x, dtype=np.float64)
    last = len(weights) - 1
    for i, (w, b) in enumerate(zip(weights, biases)):
        h = h @ w + b
        if i != last:
            h = np.maximum(h, 0.0)
    return h


def forward_sim(start_state, actions, weights, biases, norm):
    """dynamics_model.py:204-240.  actions [K,H,da] -> states [H+1,K,d] (float64)."""
    actions = np.asarray(actions, dtype=np.float64)
    K, H, _ = actions.shape
    s = np.tile(np.asarray(start_state, dtype=np.float64), (K, 1))
    out = [s.copy()]
    with np.errstate(divide="ignore", invalid="ignore"):
        for t in range(H):
            xs = np.nan_to_num((s - norm["mean_x"]) / norm["std_x"])            # :228
            ys = np.nan_to_num((actions[:, t, :] - norm["mean_y"]) / norm["std_y"])  # :229
            z = mlp_forward(np.concatenate([xs, ys], axis=1), weights, biases)  # :230-233
            s = s + (z * norm["std_z"] + norm["mean_z"])                        # :234-237
            out.append(s.copy())
    return np.stack(out)


def score_add_delta(states, desired_states, distances_left, radii, wp_index, gamma, hpf,
                    penalty_mode=PENALTY_REFERENCE, lambdas_out=None):
    """NND_MB_agent.py:566-628.  states [H+1,K,d] -> scores [K] (higher is better)."""
    states = np.asarray(states, dtype=np.float64)
    DS = np.asarray(desired_states, dtype=np.float64)
    DL = np.asarray(distances_left, dtype=np.float64)
    r = np.asarray(radii, dtype=np.float64)
    T, K, _ = states.shape
    W = len(DS)

    def dist(a, b):                                   # numerical.py:116-124
        return np.sqrt((((a - b) / r) ** 2).sum(-1))

    idx = np.full(K, int(wp_index), dtype=np.int64)
    prev = DL[idx] + dist(states[0], DS[idx])         # :574-577
    scores = np.zeros(K)
    with np.errstate(divide="ignore", invalid="ignore"):
        for t in range(T):
            x = states[t]
            dc = dist(DS[idx], x)                                         # :588-592
            dn = dist(DS[np.minimum(idx + 1, W - 1)], x)
            mv = np.logical_and(np.logical_or(dc <= 1.0, dn <= dc), idx != W - 1)  # :491-496
            idx = idx + mv                                                # :602
            dc = np.where(mv, dn, dc)                                     # :605
            to_end = DL[idx] + dc                                         # :608
            scores += (prev - to_end) * gamma ** t                        # :611
            prev = to_end
            b0 = np.maximum(idx - 1, 0)                                   # :615
            a_ = (x - DS[b0]) / r                                         # numerical.py:76-88
            b_ = (DS[b0 + 1] - DS[b0]) / r
            if penalty_mode == PENALTY_REFERENCE:
                lam = (a_ * b_).sum() / (b_ * b_).sum()                   # numerical.py:89-93 (no axis)
                if lambdas_out is not None:
                    lambdas_out.append(((a_ * b_).sum(), (b_ * b_).sum()))
            else:
                lam = ((a_ * b_).sum(-1) / (b_ * b_).sum(-1))[:, None]
            pen = np.sqrt(((lam * b_ - a_) ** 2).sum(-1))                 # numerical.py:80-81
            scores -= pen * hpf * gamma                                   # :622 (gamma^1, not gamma^t)
    return scores


def plan(start_state, actions, weights, biases, norm, desired_states, distances_left, radii,
         wp_index, gamma, hpf, penalty_mode=PENALTY_REFERENCE):
    """get_best_sim_actions (NND_MB_agent.py:498-520) for given action samples.

    Returns dict(best_k, best_score, scores, best_sequence [H,da], best_path [H+1,d],
    best_action [da]).  best_k is the *first* argmax (np.argmax; NaN counts as max).
    """
    states = forward_sim(start_state, actions, weights, biases, norm)
    scores = score_add_delta(states, desired_states, distances_left, radii, wp_index, gamma, hpf,
                             penalty_mode)
    best = int(np.argmax(scores))
    return dict(best_k=best, best_score=float(scores[best]), scores=scores,
                best_sequence=np.asarray(actions, dtype=np.float64)[best].copy(),
                best_action=np.asarray(actions, dtype=np.float64)[best, 0].copy(),
                best_path=states[:, best].copy(), states=states)
