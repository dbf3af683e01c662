"""Stage 1 on the B200 through the C ABI vs the oracle / the reference's golden vectors."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import kde_oracle
from smartstartcontinuous_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

RTOL = 1e-4      # north-star tolerance for the fp32 path (densities, ucb)


def _ucb_close(ucb, oucb, values, alpha=1.0):
    """ucb = alpha * V + bonus: each term is good to RTOL, so the sum is good to
    RTOL * (|alpha V| + bonus) -- not to RTOL * |ucb| when the two terms cancel."""
    bonus = oucb - alpha * values.astype(np.float64)
    bound = RTOL * (np.abs(alpha * values) + np.abs(bonus))
    finite = np.isfinite(oucb)
    assert np.all(np.abs(ucb[finite] - oucb[finite]) <= bound[finite])
    assert np.array_equal(ucb[~finite], oucb[~finite], equal_nan=True)


def _check_choice(best_j, ucb_oracle, rtol=RTOL):
    """Index must match wherever the oracle's top-2 gap exceeds the tolerance."""
    order = np.argsort(-ucb_oracle, kind="stable")
    top, second = ucb_oracle[order[0]], ucb_oracle[order[1]] if len(order) > 1 else -np.inf
    if top - second > 2 * rtol * abs(top):
        assert best_j == int(np.argmax(ucb_oracle))
    else:
        assert ucb_oracle[best_j] >= top - 2 * rtol * abs(top)


@pytest.mark.parametrize("name", ["kde_pendulum.npz", "kde_mountaincar.npz"])
def test_golden_selection(engine, name):
    g = load_golden(name)
    best, best_ucb, dens, ucb = engine.select_start(
        g["in_all_states"], g["in_queries"], g["in_values"], int(g["in_n_transitions"]),
        float(g["in_volume"]), float(g["in_alpha"]), float(g["in_beta"]), want_density=True, want_ucb=True)
    np.testing.assert_allclose(dens, g["out_density"], rtol=RTOL)
    _ucb_close(ucb, g["out_ucb"], g["in_values"])
    _check_choice(best, g["out_ucb"])
    assert best_ucb == pytest.approx(ucb[best], rel=1e-12)


@pytest.mark.parametrize("d,n,m", [(1, 700, 33), (2, 5000, 1000), (3, 20001, 2049), (4, 3000, 257),
                                   (6, 4000, 100), (8, 2500, 64), (11, 3000, 50), (17, 2000, 40)])
def test_random_dims_vs_oracle(engine, d, n, m):
    rng = np.random.default_rng(d * 1000 + n)
    A = rng.normal(size=(d, d)) + 2 * np.eye(d)
    data = rng.normal(size=(n, d)) @ A + rng.normal(size=d) * 5
    q = data[rng.choice(n, m, replace=False)] + 0.05 * rng.normal(size=(m, d))
    vals = rng.normal(size=m).astype(np.float32)
    best, _, dens, ucb = engine.select_start(data, q, vals, n - 1, 0.37, 1.0, 2.0, want_density=True,
                                             want_ucb=True)
    obest, odens, oucb = kde_oracle.select_start(data, q, vals, n - 1, 0.37, 1.0, 2.0)
    np.testing.assert_allclose(dens, odens, rtol=RTOL)
    _ucb_close(ucb, oucb, vals)
    _check_choice(best, oucb)


@pytest.mark.parametrize("d,n,m,no_tc", [(24, 5000, 300, False), (32, 4000, 200, False), (3, 20001, 2049, True),
                                         (8, 3000, 100, True)])
def test_cuda_core_expanded_variant_vs_oracle(engine, monkeypatch, d, n, m, no_tc):
    """kde_pairs_kernel<D, EXPANDED>: the pair kernel on the CUDA cores -- taken for 20 < d <= 32,
    and for any d with SS_KDE_NO_TC=1 (the library reads the variable at every call)."""
    if no_tc:
        monkeypatch.setenv("SS_KDE_NO_TC", "1")
    rng = np.random.default_rng(d * 77 + n)
    A = rng.normal(size=(d, d)) / np.sqrt(d) + 2 * np.eye(d)
    data = rng.normal(size=(n, d)) @ A + rng.normal(size=d) * 3
    q = data[rng.choice(n, m, replace=False)] + 0.05 * rng.normal(size=(m, d))
    vals = rng.normal(size=m).astype(np.float32)
    engine.set_timing(True)
    try:
        best, _, dens, ucb = engine.select_start(data, q, vals, n - 1, 0.37, 1.0, 2.0, want_density=True, want_ucb=True)
        assert dict(engine.last_timings()).get("kde_pairs") is not None
    finally:
        engine.set_timing(False)
    obest, odens, oucb = kde_oracle.select_start(data, q, vals, n - 1, 0.37, 1.0, 2.0)
    np.testing.assert_allclose(dens, odens, rtol=RTOL)
    _ucb_close(ucb, oucb, vals)
    _check_choice(best, oucb)
    if no_tc:
        monkeypatch.delenv("SS_KDE_NO_TC")
        _, _, dens_tc, _ = engine.select_start(data, q, vals, n - 1, 0.37, 1.0, 2.0, want_density=True)
        np.testing.assert_allclose(dens_tc, dens, rtol=5e-5)       # tcgen05 and CUDA-core variants agree


def test_far_queries_are_rescued_in_fp64(engine):
    """Queries far from every data point underflow the fp32 sum; the kernel recomputes them in
    fp64 exactly like scipy so the (huge) exploration bonus still ranks them correctly."""
    rng = np.random.default_rng(5)
    data = rng.normal(size=(4000, 3))
    q = np.concatenate([data[:30], np.array([[40.0, 0, 0], [0, 45.0, 0], [30.0, 30.0, 0.0]])])
    vals = np.zeros(len(q), dtype=np.float32)
    best, _, dens, ucb = engine.select_start(data, q, vals, 3999, 1.0, 1.0, 2.0, want_density=True,
                                             want_ucb=True)
    obest, odens, oucb = kde_oracle.select_start(data, q, vals, 3999, 1.0, 1.0, 2.0,
                                                 density_fn=kde_oracle.kde_density_direct)
    ok = odens > 0
    np.testing.assert_allclose(dens[ok], odens[ok], rtol=1e-4)
    assert np.all(dens[~ok] == 0) and np.all(np.isinf(ucb[~ok]))
    assert best == obest


def test_expanded_and_difference_pair_kernels_agree(engine):
    """The pair kernel builds the exponent as -|q|^2 - |x|^2 + 2 q.x while max|y|^2 is small and
    from exact differences otherwise (decided on the device).  One far-away extra query flips the
    whole call to the difference variant: the densities of the other queries must not move."""
    rng = np.random.default_rng(6)
    data, s2, _ = syn.pendulum_buffer(30000, seed=3)
    q = s2[rng.choice(len(s2), 2000, replace=False)]
    vals = np.zeros(len(q) + 1, dtype=np.float32)
    _, _, dens_a, _ = engine.select_start(data, q, vals[:-1], len(data) - 1, 1.0, 1.0, 2.0, want_density=True)
    far = np.concatenate([q, [[300.0, -250.0, 900.0]]])
    _, _, dens_b, _ = engine.select_start(data, far, vals, len(data) - 1, 1.0, 1.0, 2.0, want_density=True)
    assert dens_b[-1] == 0.0
    np.testing.assert_allclose(dens_a, dens_b[:-1], rtol=2e-5)
    _, odens, _ = kde_oracle.select_start(data, q[:300], vals[:300], len(data) - 1, 1.0, 1.0, 2.0)
    np.testing.assert_allclose(dens_a[:300], odens, rtol=RTOL)
    np.testing.assert_allclose(dens_b[:300], odens, rtol=RTOL)


def test_first_max_and_nan_semantics(engine):
    """np.argmax: first maximum wins; a NaN counts as the maximum."""
    rng = np.random.default_rng(6)
    data = rng.normal(size=(1000, 2))
    q = np.repeat(data[10:11], 300, axis=0)          # identical queries -> identical ucb
    vals = np.zeros(300, dtype=np.float32)
    best, _, _, ucb = engine.select_start(data, q, vals, 999, 1.0, 1.0, 2.0, want_ucb=True)
    assert np.all(ucb == ucb[0]) and best == 0
    vals[137] = np.nan
    vals[250] = np.nan
    best, best_ucb, _, _ = engine.select_start(data, q, vals, 999, 1.0, 1.0, 2.0)
    assert best == 137 and np.isnan(best_ucb)


def test_errors(engine):
    rng = np.random.default_rng(7)
    data = rng.normal(size=(100, 3))
    with pytest.raises(ValueError):
        engine.select_start(data[:3], data[:2], np.zeros(2, np.float32), 2)          # n <= d
    with pytest.raises(np.linalg.LinAlgError):
        flat = data.copy(); flat[:, 2] = flat[:, 0]                                   # singular covariance
        engine.select_start(flat, flat[:5], np.zeros(5, np.float32), 99)
    with pytest.raises(ValueError):
        engine.select_start(data, data[:5, :2], np.zeros(5, np.float32), 99)          # d mismatch


def test_full_size_c2_properties(engine):
    """BASELINE config 2 size (100 001 x 16 384, d=3): checked against the oracle on a query
    subset plus size-independent properties (permutation invariance of the data set)."""
    all_states, s2, _ = syn.pendulum_buffer(100_000, seed=0)
    rng = np.random.default_rng(0)
    idx = rng.choice(100_000, 16_384, replace=False)
    q = s2[idx]
    vals = syn.critic_like_values(q)
    best, best_ucb, dens, ucb = engine.select_start(all_states, q, vals, 100_000, 1e-3, 1.0, 2.0,
                                                    want_density=True, want_ucb=True)
    # the float64 oracle on ALL 16 384 queries (~15 s of numpy): densities, UCB and the choice
    obest, odens, oucb = kde_oracle.select_start(all_states, q, vals, 100_000, 1e-3, 1.0, 2.0)
    np.testing.assert_allclose(dens, odens, rtol=RTOL)
    _ucb_close(ucb, oucb, vals)
    _check_choice(best, oucb)
    assert best == int(np.argmax(ucb)) and best_ucb == ucb[best]
    perm = rng.permutation(len(all_states))
    best2, _, dens2, _ = engine.select_start(all_states[perm], q, vals, 100_000, 1e-3, 1.0, 2.0,
                                             want_density=True)
    np.testing.assert_allclose(dens2, dens, rtol=2e-5)
    _check_choice(best2, ucb)


def test_full_size_c5_properties(engine):
    """BASELINE config 5 KDE size (1 000 001 x 16 384, d=3): oracle on a 64-query subset, argmax
    consistency with the returned UCB vector, and linearity -- the density over the whole buffer is
    the weighted mean of the densities over its two halves evaluated with the SAME bandwidth, which
    scipy's estimator does not offer directly, so the property used is query-subset consistency: a
    slice of the queries evaluated alone reproduces its densities (different tiling, same values)."""
    all_states, s2, _ = syn.pendulum_buffer(1_000_000, seed=1)
    rng = np.random.default_rng(1)
    idx = rng.choice(1_000_000, 16_384, replace=False)
    q = s2[idx]
    vals = syn.critic_like_values(q)
    best, best_ucb, dens, ucb = engine.select_start(all_states, q, vals, 1_000_000, 1e-3, 1.0, 2.0,
                                                    want_density=True, want_ucb=True)
    # the float64 oracle on a 1 536-query subset (~15 s of numpy at 1 M points); the selection among
    # exactly those queries must be the oracle's
    sub = np.sort(rng.choice(16_384, 1536, replace=False))
    obest, odens, oucb = kde_oracle.select_start(all_states, q[sub], vals[sub], 1_000_000, 1e-3, 1.0, 2.0)
    np.testing.assert_allclose(dens[sub], odens, rtol=RTOL)
    _ucb_close(ucb[sub], oucb, vals[sub])
    best_sub, _, _, _ = engine.select_start(all_states, q[sub], vals[sub], 1_000_000, 1e-3, 1.0, 2.0)
    _check_choice(best_sub, oucb)
    assert best == int(np.argmax(ucb)) and best_ucb == ucb[best]
    part = slice(5000, 5000 + 777)
    _, _, dens_part, _ = engine.select_start(all_states, q[part], vals[part], 1_000_000, 1e-3, 1.0, 2.0,
                                             want_density=True)
    np.testing.assert_allclose(dens_part, dens[part], rtol=2e-5)


def test_device_mirror_selection_matches_host_path(engine):
    """Row f2: selection from the incrementally synced device mirror of the replay buffer's state
    ring equals selection from host arrays -- while the buffer grows, after it wraps (eviction),
    and after it is rebuilt."""
    from smartstartcontinuous_b200.replay_buffer import ReplayBuffer
    rng = np.random.default_rng(31)
    agent = object()
    rb = ReplayBuffer(agent, 3000)
    obs, _ = syn.pendulum_rollouts(rng, 30, 200)
    vals_all = rng.normal(size=4096).astype(np.float32)
    added = 0
    for ep, traj in enumerate(obs):
        rb.start_new_episode(agent)
        for t in range(len(traj) - 1):
            rb.add(agent, traj[t], np.zeros(1), 0.0, False, traj[t + 1])
            added += 1
        if ep in (3, 9, 14, 22, 29):                      # grows, fills (ep 14), wraps, wraps again
            ring = rb.state_ring()
            assert ring is not None
            idx = rb.get_possible_smart_start_indices(500)
            vals = vals_all[:len(idx)]
            want = engine.select_start(rb.get_all_states(), rb.states_s2(idx), vals, len(rb), 0.5, 1.0, 2.0,
                                       want_density=True, want_ucb=True)
            got = engine.select_start_mirror(ring, idx, vals, len(rb), 0.5, 1.0, 2.0, want_density=True, want_ucb=True)
            np.testing.assert_allclose(got[2], want[2], rtol=1e-5)      # same points, other summation order
            _ucb_close(got[3], want[3], vals)
            _check_choice(got[0], want[3])
            # and directly against the float64 oracle on the buffer's contents
            obest, odens, oucb = kde_oracle.select_start(rb.get_all_states(), rb.states_s2(idx), vals, len(rb), 0.5,
                                                         1.0, 2.0)
            np.testing.assert_allclose(got[2], odens, rtol=RTOL)
            _ucb_close(got[3], oucb, vals)
            _check_choice(got[0], oucb)
    assert added > 3000                                    # the ring did wrap
    rb._rebuild_mirror()                                   # new ring object: uploaded whole again
    ring = rb.state_ring()
    idx = rb.get_possible_smart_start_indices(200)
    want = engine.select_start(rb.get_all_states(), rb.states_s2(idx), vals_all[:len(idx)], len(rb), 0.5, 1.0, 2.0,
                               want_density=True)
    got = engine.select_start_mirror(ring, idx, vals_all[:len(idx)], len(rb), 0.5, 1.0, 2.0, want_density=True)
    np.testing.assert_allclose(got[2], want[2], rtol=1e-5)


def _random_value_net(rng, d, da, h1, h2, layer_norm, last_tanh, norms):
    def lin(i, o, scale=None):
        s = scale if scale is not None else 1.0 / np.sqrt(i)
        return rng.uniform(-s, s, (i, o)).astype(np.float32), rng.uniform(-s, s, o).astype(np.float32)
    net = dict(actor=[lin(d, h1), lin(h1, h2), lin(h2, da, 3e-3)],
               critic=[lin(d, h1), lin(h1 + da, h2), lin(h2, 1, 3e-3)], last_layer_tanh=last_tanh)
    if layer_norm:
        gb = lambda n: (rng.uniform(0.5, 1.5, n).astype(np.float32), rng.normal(0, 0.1, n).astype(np.float32))
        net["actor_ln"], net["critic_ln"] = [gb(h1), gb(h2)], [gb(h1), gb(h2)]
    if norms == "clip_only":
        # normalize_returns=False with a finite return_range: the clip still applies (ddpg_editted.py:130-131)
        net.update(ret_clip=(-0.002, 0.002), obs_clip=(-2.0, 2.0))
    elif norms:
        net.update(obs_mean=rng.normal(size=d), obs_std=rng.uniform(0.5, 2.0, d), obs_clip=(-5.0, 5.0),
                   ret_mean=-3.0, ret_std=2.5, ret_clip=(-4.0, 4.0))
    return net


@pytest.mark.parametrize("cfg", [(3, 1, 64, 32, False, True, False),      # the example's DDPG nets
                                 (2, 1, 64, 64, True, False, True), (4, 2, 200, 100, True, True, True),
                                 (3, 1, 64, 32, False, True, "clip_only"),
                                 # register-tiled kernel with widths that are not multiples of 16 and da = 2 / 5
                                 (4, 2, 48, 40, True, True, True), (5, 5, 20, 7, False, False, False)])
def test_value_net_matches_oracle_and_feeds_the_ucb(engine, cfg):
    """Row f4: critic(q, actor(q)) on the device vs the float32 numpy restatement, and a selection
    with values=None equals the selection fed with the host-evaluated values."""
    from oracle import value_oracle
    d, da, h1, h2, ln, last_tanh, norms = cfg
    rng = np.random.default_rng(41 + d)
    net = _random_value_net(rng, d, da, h1, h2, ln, last_tanh, norms)
    data = rng.normal(size=(6000, d)) * rng.uniform(0.5, 3.0, d)
    q = data[rng.choice(6000, 700, replace=False)]
    engine.set_value_net(net)
    try:
        got = engine.state_values(q)
        want = value_oracle.state_values(q, net)
        np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-6)
        a = engine.select_start(data, q, None, 5999, 0.3, 1.0, 2.0, want_ucb=True)
        b = engine.select_start(data, q, got, 5999, 0.3, 1.0, 2.0, want_ucb=True)
        assert a[0] == b[0]
        np.testing.assert_array_equal(a[3], b[3])
    finally:
        engine.set_value_net(None)
    with pytest.raises(ValueError):
        engine.select_start(data, q, None, 5999, 0.3, 1.0, 2.0)


def test_smart_start_with_device_value_net_equals_host_values(engine):
    """SmartStartContinuous(value_net=...): the candidates' values come from the device critic in front
    of the UCB (row f4) and agent.get_state_value is not called; the choice equals the one made with the
    same values evaluated on the host by the numpy restatement of the nets."""
    import random

    from oracle import value_oracle
    from smartstartcontinuous_b200.smart_start import SmartStartContinuous

    rng = np.random.default_rng(77)
    net = _random_value_net(rng, 3, 1, 64, 32, False, True, True)

    class Box:
        low, high, shape = np.array([-2.0]), np.array([2.0]), (1,)

    class Env:
        action_space = Box()

    class Base:
        calls = 0

        def get_action(self, s): return np.zeros(1)
        def observe(self, *a): pass
        def start_new_episode(self, s): pass
        def end_episode(self): pass
        def get_param_dict(self): return {}

        def get_state_value(self, states):
            Base.calls += 1
            return value_oracle.state_values(np.asarray(states), net).reshape(-1, 1)

    td = dict(dataX=rng.normal(size=(64, 3)), dataY=rng.normal(size=(64, 1)), dataZ=rng.normal(size=(64, 3)))
    s_all, _, _ = syn.pendulum_buffer(6000, seed=5)
    picks = []
    try:
        for value_net in (None, net, lambda: net):
            ss = SmartStartContinuous(Base(), Env(), None, buffer_size=6000, n_ss=900, print_ss_stuff=False,
                                      nnd_mb_num_fc_layers=1, nnd_mb_depth_fc_layers=32, nnd_mb_verbose=False,
                                      engine=engine, nnd_mb_extra=dict(training_data=td), value_net=value_net)
            rb = ss.replay_buffer
            for i in range(6000):
                if i % 200 == 0:
                    rb.start_new_episode(ss)
                rb.add(ss, s_all[i], np.zeros(1), 0.0, (i % 200) == 199, s_all[i + 1])
            random.seed(11)
            before = Base.calls
            path = ss.get_smart_start_path()
            assert (Base.calls == before) == (value_net is not None)
            picks.append((ss.last_selection, np.asarray(path[-1])))
    finally:
        engine.set_value_net(None)
    (i0, u0), p0 = picks[0]
    for (i1, u1), p1 in picks[1:]:
        assert i1 == i0 and abs(u1 - u0) <= 1e-4 * abs(u0) and np.array_equal(p0, p1)


@pytest.mark.parametrize("name,env,n,n_ss,seed", [("kde_pendulum.npz", "pendulum", 3000, 300, 0),
                                                  ("kde_mountaincar.npz", "mountaincar", 2000, 5000, 1)])
def test_smart_start_path_equals_the_references_from_the_seed(engine, name, env, n, n_ss, seed):
    """The selection goldens come from the reference's own get_smart_start_path under random.seed(seed): the
    drop-in class over the same (rebuilt) buffer, from the same seed, chooses the same buffer index and returns
    the same path -- candidate draw, device-mirror selection and episodic path extraction end to end."""
    import random

    from smartstartcontinuous_b200.smart_start import SmartStartContinuous
    from test_replay_buffer import _golden_buffer

    g = load_golden(name)
    rb, episodes = _golden_buffer(name, env, n, seed)
    d = episodes[0][0].shape[1]

    class Box:
        low, high, shape = np.array([-2.0]), np.array([2.0]), (1,)

    class Env:
        action_space = Box()

    class Base:
        replay_buffer = rb

        def set_replay_buffer_main_agent(self, main): pass
        def get_action(self, s): return np.zeros(1)
        def observe(self, *a): pass
        def start_new_episode(self, s): pass
        def end_episode(self): pass
        def get_param_dict(self): return {}
        def get_state_value(self, states): return syn.critic_like_values(np.asarray(states), seed + 7).reshape(-1, 1)

    rng = np.random.default_rng(0)
    td = dict(dataX=rng.normal(size=(64, d)), dataY=rng.normal(size=(64, 1)), dataZ=rng.normal(size=(64, d)))
    ss = SmartStartContinuous(Base(), Env(), None, n_ss=n_ss, print_ss_stuff=False, nnd_mb_num_fc_layers=1,
                              nnd_mb_depth_fc_layers=8, nnd_mb_verbose=False, engine=engine,
                              nnd_mb_extra=dict(training_data=td))
    assert ss.replay_buffer is rb
    ss.nnd_mb_agent.radii = g["in_radii"] if len(g["in_radii"]) else None
    random.seed(seed)
    path = ss.get_smart_start_path()
    chosen, ucb = ss.last_selection
    assert chosen == int(g["out_chosen_buffer_index"])
    assert ucb == pytest.approx(float(g["out_ucb"][int(g["out_best_j"])]), rel=1e-4)
    np.testing.assert_array_equal(np.asarray(path), g["out_path"])


def test_mirror_running_moments_track_the_buffer(engine):
    """The estimator of a mirror selection is fitted from moments maintained incrementally by the mirror writes
    (new rows added, overwritten rows subtracted, exact recompute after bulk uploads and once per buffer
    turnover): through growth and many ring wraps the densities equal those of the full re-reduction
    (SS_MIRROR_NO_RUNNING_MOMENTS) to 1e-9."""
    import os

    from smartstartcontinuous_b200.replay_buffer import ReplayBuffer
    rng = np.random.default_rng(77)
    agent = object()
    rb = ReplayBuffer(agent, 1000)
    obs, _ = syn.pendulum_rollouts(rng, 40, 150)
    vals = rng.normal(size=300).astype(np.float32)
    for ep, traj in enumerate(obs):
        rb.start_new_episode(agent)
        for t in range(len(traj) - 1):
            rb.add(agent, traj[t], np.zeros(1), 0.0, False, traj[t + 1])
        if ep < 1:
            continue
        ring = rb.state_ring()
        idx = rb.get_possible_smart_start_indices(300)
        got = engine.select_start_mirror(ring, idx, vals[:len(idx)], len(rb), 0.5, 1.0, 2.0, want_density=True)
        os.environ["SS_MIRROR_NO_RUNNING_MOMENTS"] = "1"
        try:
            want = engine.select_start_mirror(ring, idx, vals[:len(idx)], len(rb), 0.5, 1.0, 2.0, want_density=True)
        finally:
            del os.environ["SS_MIRROR_NO_RUNNING_MOMENTS"]
        np.testing.assert_allclose(got[2], want[2], rtol=1e-9, err_msg="episode %d" % ep)
        assert got[0] == want[0]
