"""numpy's legacy RandomState stream generated on the device (csrc/mt19937.cu) against numpy itself:
the samples of npr.uniform(low, high, (K, H, da)) (NND_MB_agent.py:500-501) and the generator state
after the draw, bit for bit, for every position of the key block, sharded and unsharded."""
import numpy as np
import pytest

from smartstartcontinuous_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _rs(seed, burn):
    rs = np.random.RandomState(seed)
    if burn:
        rs.random_sample(burn)
    return rs


def _check(engine, rs, n_rows, low, high, shards=1):
    low, high = np.asarray(low, dtype=np.float64), np.asarray(high, dtype=np.float64)
    da = low.shape[0]
    state = rs.get_state()
    want = rs.uniform(low, high, (n_rows, da)).reshape(-1)
    after = rs.get_state()
    n = n_rows * da
    bounds = np.linspace(0, n_rows, shards + 1).astype(np.int64) * da
    for s in range(shards):
        first, count = int(bounds[s]), int(bounds[s + 1] - bounds[s])
        if count == 0:
            continue
        ptr = engine.mt19937_uniform(state, n, low, high, first=first, count=count)
        got = engine.read_device_doubles(ptr, count)
        assert np.array_equal(got, want[first:first + count]), (n_rows, da, s, shards)
        st = engine.mt19937_state()
        assert st[0] == "MT19937" and st[2] == after[2] and np.array_equal(st[1], after[1]), (n_rows, da, s)
        assert st[3:] == state[3:]


@pytest.mark.parametrize("burn", [0, 1, 2, 155, 311, 312, 313, 1000])
def test_uniform_matches_numpy_every_block_position(engine, burn):
    """burn = doubles consumed before the draw: pos 624 (fresh seed), even positions, block ends."""
    for n_rows in (1, 2, 5, 311, 312, 313, 624, 700, 4097):
        _check(engine, _rs(11 + burn, burn), n_rows, [-2.0], [2.0])


def test_odd_positions_and_multi_dimensional_actions(engine):
    """32-bit draws leave the generator on an odd word: doubles then straddle blocks."""
    for seed in range(4):
        rs = np.random.RandomState(seed)
        rs.randint(0, 2 ** 31, size=2 * seed + 1)            # odd number of words consumed
        assert rs.get_state()[2] % 2 == 1
        _check(engine, rs, 2000 + seed, [-1.0, 0.0, -3.5], [1.0, 0.25, 7.0])
    rs = np.random.RandomState(99)
    st = list(rs.get_state())
    st[2] = 0                                                 # a state numpy accepts but never leaves behind
    rs.set_state(tuple(st))
    _check(engine, rs, 1000, [-0.3, 0.1], [0.4, 0.9])


@pytest.mark.parametrize("n_rows,da", [(99840, 1), (100000, 1), (163840, 2), (4096 * 20, 1), (5000 * 4, 1)])
def test_jump_ahead_segments(engine, n_rows, da):
    """Streams long enough to be cut into per-SM segments (>= 320 blocks of 624 words), and the
    BASELINE config 1 / 3 sizes that stay sequential."""
    low, high = [-2.0, -1.0][:da], [2.0, 3.0][:da]
    _check(engine, _rs(5, 77), n_rows, low, high)
    rs = np.random.RandomState(6)
    rs.randint(0, 2 ** 31, size=3)
    _check(engine, rs, n_rows, low, high)


def test_config4_size_and_shards(engine):
    """K = 131072, H = 50 (6.5 M doubles, 21 009 blocks): whole draw, and the 8 shards of it."""
    _check(engine, _rs(1, 12345), 131072 * 50, [-2.0], [2.0])
    rs = np.random.RandomState(3)
    rs.randint(0, 2 ** 31, size=1)
    _check(engine, rs, 131072 * 50, [-2.0], [2.0], shards=8)
    _check(engine, _rs(2, 5), 20000 * 12, [-2.0], [2.0], shards=3)


def test_cached_gaussian_travels_with_the_state(engine):
    rs = np.random.RandomState(8)
    rs.normal()                                               # leaves has_gauss = 1
    assert rs.get_state()[3] == 1
    _check(engine, rs, 3000, [-2.0], [2.0])


def test_errors(engine):
    st = np.random.RandomState(0).get_state()
    with pytest.raises(ValueError):
        engine.mt19937_uniform(st, 0, [-1.0], [1.0])
    with pytest.raises(ValueError):
        engine.mt19937_uniform(st, 10, [-1.0], [1.0], first=8, count=5)
    with pytest.raises(ValueError):
        engine.mt19937_uniform(("PCG64",) + tuple(st[1:]), 10, [-1.0], [1.0])


def _planner_setup(engine, L=2, h=64, seed=0):
    rng = np.random.default_rng(seed)
    s, a = syn.mountaincar_rollout(rng, 600)
    weights, biases = syn.xavier_mlp(rng, 2, 1, L, h, scale=0.3)
    norm = syn.normalisation_stats(s, a)
    engine.set_model(weights, biases, norm)
    ds = s[::40][:8]
    dl = np.linspace(3.0, 0.0, len(ds))
    engine.set_plan(ds, dl, np.array([0.1, 0.01]))
    return s[0]


@pytest.mark.parametrize("mode", ["reference", "per_sample"])
def test_plan_with_device_generated_numpy_stream_equals_host_draw(engine, mode):
    """Engine.plan(rng_state=...) = the decision on the host draw of the same generator, and the state
    it returns is the host generator's after that draw."""
    state0 = _planner_setup(engine)
    for K, H in ((300, 7), (4096, 20), (20000, 12)):
        rs = _rs(K, 17)
        st = rs.get_state()
        acts = rs.uniform(np.array([-1.0]), np.array([1.0]), (K, H, 1))
        want = engine.plan(state0, 0, actions=acts, penalty_mode=mode, precision="fp32", want_scores=True)
        got = engine.plan(state0, 0, K=K, H=H, act_low=[-1.0], act_high=[1.0], rng_state=st, penalty_mode=mode,
                          precision="fp32", want_scores=True)
        assert got["best_k"] == want["best_k"]
        assert np.array_equal(got["scores"], want["scores"])
        assert np.array_equal(got["best_sequence"], want["best_sequence"])
        assert np.array_equal(got["best_path"], want["best_path"])
        after = rs.get_state()
        assert got["rng_state"][2] == after[2] and np.array_equal(got["rng_state"][1], after[1])


def test_agent_default_path_draws_numpys_stream_on_the_device(engine):
    """NND_MB_agent.get_best_sim_actions (default) = the host_rng=True agent: same decisions, and the
    global numpy stream continues identically afterwards."""
    from smartstartcontinuous_b200.nnd_mb_agent import NND_MB_agent

    class Box:
        low, high, shape = np.array([-1.0]), np.array([1.0]), (1,)

    class Env:
        action_space = Box()

    rng = np.random.default_rng(4)
    s, a = syn.mountaincar_rollout(rng, 900)
    data = dict(dataX=s[:-1], dataY=a, dataZ=s[1:] - s[:-1])
    # the library reaches numpy's global state struct directly (no get_state / set_state round trip)
    assert engine.global_rng_address() is not None
    outs = []
    for host_rng in (True, False, None):
        np.random.seed(77)
        ag = NND_MB_agent(Env(), None, horizon=9, num_control_samples=700, num_fc_layers=2, depth_fc_layers=64,
                          nEpoch=1, verbose=False, training_data=data, engine=engine, seed=0, host_rng=host_rng,
                          precision="fp32")
        ag.start_new_episode_plan(s[0], s[:120])
        res = []
        for t in range(3):
            act, k, seq, path = ag.get_best_sim_actions(s[t])
            res.append((act.copy(), k, seq.copy(), path.copy()))
        outs.append((res, np.concatenate([np.random.random_sample(5), np.random.normal(size=3), np.random.randint(0, 1000, 4)])))
    for other in (1, 2):
        for (a0, k0, s0, p0), (a1, k1, s1, p1) in zip(outs[0][0], outs[other][0]):
            # (every agent trains its own model: the device trainer's Adam slots persist across agents of one
            # Engine, so the predicted paths agree to rounding, the drawn sequences exactly)
            assert k0 == k1 and np.array_equal(a0, a1) and np.array_equal(s0, s1)
            assert np.allclose(p0, p1, rtol=1e-4, atol=1e-6)
        assert np.array_equal(outs[0][1], outs[other][1])


@pytest.mark.parametrize("force_p", [2, 5, 9, 40, 148])
def test_forced_segment_layouts(engine, force_p):
    """Every layout the planner can pick -- few segments whose jumps are shared by many CTAs, many segments
    with one CTA each -- produces numpy's stream (SS_MT_FORCE_P overrides the cost model; a fresh stream
    length per layout keeps the plan cache from answering)."""
    import os
    old = os.environ.get("SS_MT_FORCE_P")
    os.environ["SS_MT_FORCE_P"] = str(force_p)
    try:
        n_rows = 150000 + 1000 * force_p
        _check(engine, _rs(40 + force_p, 321), n_rows, [-2.0], [2.0])
        rs = np.random.RandomState(force_p)
        rs.randint(0, 2 ** 31, size=5)
        _check(engine, rs, n_rows + 7, [-2.0, 0.5], [2.0, 0.75], shards=2)
    finally:
        if old is None:
            os.environ.pop("SS_MT_FORCE_P", None)
        else:
            os.environ["SS_MT_FORCE_P"] = old


@pytest.mark.parametrize("name,seed", [("mpc_pendulum_2x500.npz", 4), ("mpc_mountaincar_L2.npz", 1)])
def test_default_agent_reproduces_the_reference_decision_from_the_seed(engine, name, seed):
    """The goldens were produced by the reference's own get_best_sim_actions under np.random.seed(seed)
    (oracle/make_golden.py): from the same seed the drop-in agent -- samples generated on the GPU from
    numpy's generator state -- draws the very same K x H samples (the returned sequence is the golden's
    row, bit for bit), picks the same sequence wherever the fp32 tolerance allows, returns the reference's
    first action, and leaves numpy's generator where the reference's draw leaves it."""
    from conftest import golden_model, load_golden
    from smartstartcontinuous_b200.nnd_mb_agent import NND_MB_agent

    g = load_golden(name)
    w, b, norm = golden_model(g)
    K, H, _ = g["in_actions"].shape
    d = w[-1].shape[1]

    class Box:
        low, high, shape = g["in_act_low"], g["in_act_high"], (1,)

    class Env:
        action_space = Box()

    rng = np.random.default_rng(0)
    td = dict(dataX=rng.normal(size=(64, d)), dataY=rng.normal(size=(64, 1)), dataZ=rng.normal(size=(64, d)))
    ag = NND_MB_agent(Env(), None, horizon=H, num_control_samples=K, num_fc_layers=len(w) - 1,
                      depth_fc_layers=w[0].shape[1], gamma=float(g["in_gamma"]),
                      horizontal_penalty_factor=float(g["in_hpf"]), verbose=False, training_data=td, engine=engine,
                      precision="fp32", host_rng=False)
    for k in ("mean_x", "std_x", "mean_y", "std_y", "mean_z", "std_z"):
        setattr(ag, k, np.asarray(g["in_" + k], dtype=np.float64))
        setattr(ag.dyn_model, k, getattr(ag, k))
    ag.dyn_model.set_weights(w, b)
    engine.set_plan(g["out_desired_states"], g["out_distances_left"], g["out_radii"])
    ag.desired_states, ag.distances_left, ag.radii = g["out_desired_states"], g["out_distances_left"], g["out_radii"]
    ag.current_desired_state_index = int(g["in_wp_index"])

    np.random.seed(seed)
    best_action, best_k, best_seq, best_path = ag.get_best_sim_actions(g["in_start_state"])
    after = np.random.get_state()
    np.random.seed(seed)
    np.random.uniform(g["in_act_low"], g["in_act_high"], (K, H, 1))
    want_after = np.random.get_state()
    assert after[2] == want_after[2] and np.array_equal(after[1], want_after[1])
    np.testing.assert_array_equal(best_seq, g["in_actions"][best_k])          # the reference's own draw
    scores = g["out_scores"]
    top2 = np.sort(scores)[-2:]
    if top2[1] - top2[0] > 2e-4 * max(1.0, abs(top2[1])):
        assert best_k == int(g["out_best_k"])
        np.testing.assert_array_equal(best_action, g["out_best_action"])
        np.testing.assert_allclose(best_path, g["out_best_path"], rtol=1e-4, atol=1e-4 * np.abs(g["out_best_path"]).max())
    else:
        assert scores[best_k] >= top2[1] - 2e-4 * max(1.0, abs(top2[1]))
