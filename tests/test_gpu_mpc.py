"""Stage 2 on the B200 through the C ABI vs the oracle / the reference's golden vectors."""
import numpy as np
import pytest

from conftest import golden_model, load_golden
from oracle import mpc_oracle, philox
from smartstartcontinuous_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

MPC_CASES = ["mpc_mountaincar_L2.npz", "mpc_pendulum_L1.npz", "mpc_mountaincar_L3_xavier.npz"]
STATE_RTOL = 1e-4     # fp32 path, north-star tolerance
SCORE_TOL = 1e-4


def _setup(engine, g):
    w, b, norm = golden_model(g)
    engine.set_model(w, b, norm)
    engine.set_plan(g["out_desired_states"], g["out_distances_left"], g["out_radii"])
    return w, b, norm


def _score_close(got, want, tol, max_outlier_frac=0.01):
    """Scores are continuous except at waypoint switches (dc <= 1, dn <= dc): a sample sitting on a
    switch boundary may flip in fp32.  Everything else must agree to `tol` (relative to the score
    scale); the flips must be rare."""
    scale = np.maximum(np.abs(want), 1.0)
    bad = np.abs(got - want) > tol * scale
    assert bad.mean() <= max_outlier_frac, "%.2f%% of scores off by more than %g" % (100 * bad.mean(), tol)
    return bad


def _check_best(best_k, want_scores, tol):
    order = np.argsort(-want_scores, kind="stable")
    top, second = want_scores[order[0]], want_scores[order[1]]
    if top - second > 2 * tol * max(abs(top), 1.0):
        assert best_k == int(np.argmax(want_scores))
    else:
        assert want_scores[best_k] >= top - 2 * tol * max(abs(top), 1.0)


@pytest.mark.parametrize("name", MPC_CASES)
def test_golden_forward_sim_fp32(engine, name):
    g = load_golden(name)
    _setup(engine, g)
    states = engine.forward_sim(g["in_start_state"], g["in_actions"], precision="fp32")
    scale = np.abs(g["out_states"]).max(axis=(0, 1))
    assert np.abs(states - g["out_states"]).max(axis=(0, 1)).max() <= STATE_RTOL * scale.max()
    np.testing.assert_allclose(states, g["out_states"], rtol=STATE_RTOL, atol=STATE_RTOL * scale.min())


@pytest.mark.parametrize("name", MPC_CASES)
@pytest.mark.parametrize("mode", ["reference", "per_sample"])
def test_golden_plan_fp32(engine, name, mode):
    g = load_golden(name)
    w, b, norm = _setup(engine, g)
    res = engine.plan(g["in_start_state"], int(g["in_wp_index"]), actions=g["in_actions"],
                      gamma=float(g["in_gamma"]), horizontal_penalty_factor=float(g["in_hpf"]),
                      penalty_mode=mode, precision="fp32", want_scores=True)
    if mode == "reference":
        want = g["out_scores"]
    else:
        want = mpc_oracle.score_add_delta(g["out_states"], g["out_desired_states"], g["out_distances_left"],
                                          g["out_radii"], int(g["in_wp_index"]), float(g["in_gamma"]),
                                          float(g["in_hpf"]), penalty_mode=mpc_oracle.PENALTY_PER_SAMPLE)
    _score_close(res["scores"], want, SCORE_TOL)
    _check_best(res["best_k"], want, SCORE_TOL)
    assert res["best_score"] == pytest.approx(res["scores"][res["best_k"]], rel=1e-6)
    np.testing.assert_allclose(res["best_sequence"], g["in_actions"][res["best_k"]].astype(np.float32), rtol=1e-7)
    np.testing.assert_allclose(res["best_path"], g["out_states"][:, res["best_k"]], rtol=STATE_RTOL,
                               atol=STATE_RTOL * np.abs(g["out_states"]).max())
    if mode == "reference" and res["best_k"] == int(g["out_best_k"]):
        np.testing.assert_allclose(res["best_sequence"][0], g["out_best_action"], rtol=1e-6)


def test_device_sampler_matches_numpy_philox(engine):
    got = engine.sample_actions(300, 7, 2, 0xDEADBEEF12345, [-1.0, 0.0], [1.0, 3.0], k_offset=5_000_000_000)
    want = philox.sample_actions(300, 7, 2, 0xDEADBEEF12345, [-1.0, 0.0], [1.0, 3.0], k_offset=5_000_000_000)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("mode", ["reference", "per_sample"])
def test_device_sampled_plan_matches_oracle(engine, mode):
    """Plan with Philox actions generated on the GPU, re-scored by the oracle on the same actions."""
    g = load_golden("mpc_mountaincar_L2.npz")
    w, b, norm = _setup(engine, g)
    K, H, seed = 1000, 12, 77
    res = engine.plan(g["in_start_state"], 0, K=K, H=H, seed=seed, act_low=[-1.0], act_high=[1.0],
                      penalty_mode=mode, precision="fp32", want_scores=True)
    acts = philox.sample_actions(K, H, 1, seed, [-1.0], [1.0])
    o = mpc_oracle.plan(g["in_start_state"], acts, w, b, norm, g["out_desired_states"],
                        g["out_distances_left"], g["out_radii"], 0, .75, .5,
                        penalty_mode=0 if mode == "reference" else 1)
    _score_close(res["scores"], o["scores"], SCORE_TOL)
    _check_best(res["best_k"], o["scores"], SCORE_TOL)
    np.testing.assert_array_equal(res["best_sequence"], acts[res["best_k"]])


def test_sharded_plan_equals_single(engine):
    """Two K-shards (k_offset) + summed projection sums reproduce the single-shard result: the
    multi-GPU path emulated on one GPU (sequential shards, no waiting kernels)."""
    g = load_golden("mpc_mountaincar_L2.npz")
    _setup(engine, g)
    K, H, seed = 512, 8, 5
    kw = dict(seed=seed, act_low=[-1.0], act_high=[1.0], precision="fp32")
    full = engine.plan(g["in_start_state"], 0, K=K, H=H, penalty_mode="per_sample", want_scores=True, **kw)
    parts = [engine.plan(g["in_start_state"], 0, K=K // 2, H=H, penalty_mode="per_sample", want_scores=True,
                         k_offset=off, K_global=K, **kw) for off in (0, K // 2)]
    np.testing.assert_array_equal(np.concatenate([p["scores"] for p in parts]), full["scores"])
    best = max(parts, key=lambda p: (p["best_score"], -p["best_k"]))
    assert best["best_k"] == full["best_k"]
    with pytest.raises(ValueError):
        engine.plan(g["in_start_state"], 0, K=K // 2, H=H, penalty_mode="reference", k_offset=0, K_global=K, **kw)


def test_errors(engine):
    g = load_golden("mpc_pendulum_L1.npz")
    w, b, norm = golden_model(g)
    engine.set_model(w, b, norm)
    with pytest.raises(ValueError):
        engine.set_plan(g["out_desired_states"], g["out_distances_left"], np.array([1.0, 0.0, 1.0]))
    with pytest.raises(ValueError):
        engine.set_plan(g["out_desired_states"][:1], g["out_distances_left"][:1], g["out_radii"])
    engine.set_plan(g["out_desired_states"], g["out_distances_left"], g["out_radii"])
    with pytest.raises(ValueError):
        engine.plan(g["in_start_state"], 10_000, actions=g["in_actions"])
    with pytest.raises(ValueError):
        engine.plan(g["in_start_state"][:2], 0, actions=g["in_actions"])


def test_c3_size_fp32_properties(engine):
    """BASELINE config 3 (MountainCar, K=4096, H=20, MLP 2x500): a 256-sequence slice is checked
    against the oracle; the full batch through size-independent properties."""
    rng = np.random.default_rng(3)
    roll = [syn.mountaincar_rollout(rng, 200) for _ in range(8)]
    norm = syn.normalisation_stats(np.concatenate([r[0] for r in roll]),
                                   np.concatenate([np.concatenate([r[1], r[1][-1:]]) for r in roll]))
    w, b = syn.xavier_mlp(rng, 2, 1, 2, 500, scale=0.5)
    from smartstartcontinuous_b200.nnd_mb_agent import plan_from_path
    plan = plan_from_path(list(roll[0][0][:60]), mean_per_stepsize=1, std_per_stepsize=1,
                          stepsizes_in_waypoint_radii=1, path_shortcutting=True, theta=1, steps_per_waypoint=1)
    engine.set_model(w, b, norm)
    engine.set_plan(plan["desired_states"], plan["distances_left"], plan["radii"])
    K, H = 4096, 20
    acts = rng.uniform(-1, 1, (K, H, 1))
    start = roll[0][0][0]
    res = engine.plan(start, 0, actions=acts, penalty_mode="per_sample", precision="fp32", want_scores=True)
    o = mpc_oracle.plan(start, acts[:256], w, b, norm, plan["desired_states"], plan["distances_left"],
                        plan["radii"], 0, .75, .5, penalty_mode=1)
    _score_close(res["scores"][:256], o["scores"], 2e-4, max_outlier_frac=0.02)
    assert res["best_k"] == int(np.argmax(res["scores"]))
    # permuting the sequences permutes the scores
    perm = rng.permutation(K)
    res2 = engine.plan(start, 0, actions=acts[perm], penalty_mode="per_sample", precision="fp32", want_scores=True)
    np.testing.assert_array_equal(res2["scores"], res["scores"][perm])
