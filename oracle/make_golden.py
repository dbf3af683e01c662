"""Generate tests/golden/*.npz by running the reference's own code (build container only).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.   Usage:  python -m oracle.make_golden

Every array stored under an ``out_`` key was produced by a reference function executed
through oracle/ref_harness.py (intermediate values are captured by temporarily wrapping
``np.argmax`` / ``scipy.stats.gaussian_kde`` / ``generate_scores_add_delta`` inside the
reference modules -- the reference code itself is not modified).  ``in_`` keys are the seeded
synthetic inputs.  The GPU box never runs this script; it only reads the committed files.
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh                      # noqa: E402
from smartstartcontinuous_b200 import synthetic as syn     # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _capture_selection(ref, ss):
    """Run reference get_smart_start_path, capturing densities (kernel output) and ucb_list."""
    cap = {}
    real_kde = ref.ssc.scipy.stats.gaussian_kde
    real_argmax = np.argmax

    class SpyKde(real_kde):
        def __call__(self, pts):
            out = super().__call__(pts)
            cap["queries"] = np.array(pts).T.copy()
            cap["density"] = np.array(out).copy()
            return out

    def spy_argmax(a, *args, **kw):
        cap["ucb"] = np.array(a, dtype=np.float64).reshape(-1).copy()
        return real_argmax(a, *args, **kw)

    ref.ssc.scipy.stats.gaussian_kde = SpyKde
    np.argmax = spy_argmax
    try:
        path = ss.get_smart_start_path()
    finally:
        ref.ssc.scipy.stats.gaussian_kde = real_kde
        np.argmax = real_argmax
    cap["path"] = np.array(path)
    return cap


def golden_kde(ref, name, *, env, n_transitions, n_ss, with_radii, seed):
    rng = np.random.default_rng(seed)
    steps = 100
    episodes = []
    if env == "pendulum":
        obs, act = syn.pendulum_rollouts(rng, n_transitions // steps, steps)
        episodes = [(obs[e], act[e]) for e in range(len(obs))]
        low, high = [-2.0], [2.0]
    else:
        for _ in range(n_transitions // steps):
            episodes.append(syn.mountaincar_rollout(rng, steps))
        low, high = [-1.0], [1.0]
    d = episodes[0][0].shape[1]
    rb = rh.make_replay_buffer(ref, episodes, max_size=n_transitions - 37)   # forces FIFO eviction
    w, b = syn.xavier_mlp(rng, d, 1, 1, 8)
    norm = syn.normalisation_stats(episodes[0][0], episodes[0][1])
    nnd = rh.make_nnd_agent(ref, w, b, norm, low, high, horizon=3, num_control_samples=4,
                            replay_buffer=rb)
    if with_radii:
        nnd.start_new_episode_plan(episodes[1][0][0], list(episodes[1][0][:40]))
    vseed = seed + 7
    ss = rh.make_smartstart(ref, rb, nnd, lambda s: syn.critic_like_values(s, vseed), n_ss=n_ss,
                            exploitation_param=1.0, exploration_param=2.0)
    random.seed(seed)
    idx = rb.get_possible_smart_start_indices(n_ss)
    all_states = rb.get_all_states()
    random.seed(seed)
    cap = _capture_selection(ref, ss)
    volume = ref.numerical.volume_of_n_dimensional_hyperellipsoid(nnd.radii) if with_radii else 1.0
    values = syn.critic_like_values(cap["queries"], vseed)
    chosen = int(idx[int(np.argmax(cap["ucb"]))])
    np.savez_compressed(
        os.path.join(GOLDEN, name),
        in_all_states=all_states, in_queries=cap["queries"], in_values=values,
        in_indices=idx, in_n_transitions=len(rb), in_volume=volume, in_alpha=1.0, in_beta=2.0,
        in_radii=np.array(nnd.radii if with_radii else []),
        out_density=cap["density"], out_ucb=cap["ucb"], out_best_j=int(np.argmax(cap["ucb"])),
        out_chosen_buffer_index=chosen, out_path=cap["path"])
    print(name, "n+1=%d m=%d d=%d best_j=%d path_len=%d" %
          (len(all_states), len(idx), d, int(np.argmax(cap["ucb"])), len(cap["path"])))


def golden_mpc(ref, name, *, env, L, h, K, H, wp_index, seed, fit_epochs, weight_seed=None, weight_scale=1.0):
    rng = np.random.default_rng(seed)
    if env == "pendulum":
        obs, act = syn.pendulum_rollouts(rng, 12, 120)
        states_list, actions_list = list(obs), list(act)
        low, high = [-2.0], [2.0]
    else:
        roll = [syn.mountaincar_rollout(rng, 150) for _ in range(12)]
        states_list, actions_list = [r[0] for r in roll], [r[1] for r in roll]
        low, high = [-1.0], [1.0]
    d = states_list[0].shape[1]
    if fit_epochs:
        w, b, norm = syn.fit_dynamics_mlp(states_list, actions_list, L, h, seed=seed, epochs=fit_epochs,
                                          batch=128)
    else:
        # weight_seed: the weights are regenerated from the seed by the tests (conftest.golden_model)
        # instead of being stored -- keeps the fixture of a 2x500 network small
        wrng = rng if weight_seed is None else np.random.default_rng(weight_seed)
        w, b = syn.xavier_mlp(wrng, d, 1, L, h, scale=weight_scale)
        norm = syn.normalisation_stats(np.concatenate(states_list), np.concatenate(
            [np.concatenate([a, a[-1:]]) for a in actions_list])[:len(np.concatenate(states_list))])
    nnd = rh.make_nnd_agent(ref, w, b, norm, low, high, horizon=H, num_control_samples=K)
    path = [np.array(s) for s in states_list[0][:60]]
    start_state = np.array(states_list[0][0]) + 0.01 * rng.standard_normal(d) * norm["std_x"]
    nnd.start_new_episode_plan(start_state, path)
    nnd.current_desired_state_index = wp_index
    cap = {}
    real_score = nnd.generate_scores_add_delta

    def spy(resulting_states):
        out = real_score(resulting_states)
        cap["states"] = np.array(resulting_states).copy()
        cap["scores"] = np.array(out[0]).copy()
        return out

    nnd.generate_scores_add_delta = spy
    np.random.seed(seed)
    actions = np.random.uniform(nnd.env.action_space.low, nnd.env.action_space.high, (K, H, 1))
    np.random.seed(seed)
    best_action, best_k, best_seq, best_path = nnd.get_best_sim_actions(start_state)
    assert np.array_equal(best_seq, actions[best_k])
    flat = {}
    if weight_seed is None:
        for i, (wi, bi) in enumerate(zip(w, b)):
            flat["in_w%d" % i] = wi
            flat["in_b%d" % i] = bi
    else:
        flat.update(in_weight_seed=weight_seed, in_weight_scale=weight_scale, in_hidden=h, in_d=d)
    np.savez_compressed(
        os.path.join(GOLDEN, name),
        in_path=np.array(path), in_start_state=start_state, in_actions=actions,
        in_act_low=np.array(low), in_act_high=np.array(high), in_wp_index=wp_index,
        in_gamma=nnd.gamma, in_hpf=nnd.horizontal_penalty_factor, in_num_layers=L,
        **{"in_" + k: v for k, v in norm.items()}, **flat,
        out_radii=np.array(nnd.radii), out_path_to_follow=np.array(nnd.path_to_follow),
        out_desired_states=np.array(nnd.desired_states), out_distances_left=np.array(nnd.distances_left),
        out_states=cap["states"], out_scores=cap["scores"], out_best_k=int(best_k),
        out_best_action=np.array(best_action), out_best_path=np.array(best_path))
    s = np.sort(cap["scores"])
    print(name, "d=%d L=%d h=%d K=%d H=%d W=%d best_k=%d top2 gap=%.3g" %
          (d, L, h, K, H, len(nnd.desired_states), int(best_k), s[-1] - s[-2]))


def main():
    assert rh.available(), "reference tree not mounted"
    os.makedirs(GOLDEN, exist_ok=True)
    ref = rh.load_reference()
    golden_kde(ref, "kde_pendulum.npz", env="pendulum", n_transitions=3000, n_ss=300,
               with_radii=True, seed=0)
    golden_kde(ref, "kde_mountaincar.npz", env="mountaincar", n_transitions=2000, n_ss=5000,
               with_radii=False, seed=1)
    golden_mpc(ref, "mpc_mountaincar_L2.npz", env="mountaincar", L=2, h=64, K=256, H=10, wp_index=0,
               seed=1, fit_epochs=15)
    golden_mpc(ref, "mpc_pendulum_L1.npz", env="pendulum", L=1, h=32, K=128, H=6, wp_index=3,
               seed=2, fit_epochs=15)
    golden_mpc(ref, "mpc_mountaincar_L3_xavier.npz", env="mountaincar", L=3, h=40, K=96, H=4,
               wp_index=1, seed=3, fit_epochs=0)
    # the BASELINE network shape (Pendulum, 2x500) through the reference's own get_best_sim_actions
    golden_mpc(ref, "mpc_pendulum_2x500.npz", env="pendulum", L=2, h=500, K=600, H=20, wp_index=0,
               seed=4, fit_epochs=0, weight_seed=1004, weight_scale=0.5)


if __name__ == "__main__":
    main()
