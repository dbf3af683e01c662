"""Replay buffer semantics: the reference's tests (tests/RLAgents/test_replayBuffer.py:5-100)
restated, plus the contiguous mirror used to feed the KDE kernel."""
import random

import numpy as np
import pytest

from smartstartcontinuous_b200.replay_buffer import ReplayBuffer


def test_add_and_sample():
    rb = ReplayBuffer(0, 1)
    rb.add(0, 1, 1, 1, True, 1)
    s, a, r, t, s2 = rb.sample_batch(1)
    assert s[0] == 1 and a[0] == 1 and r[0] == 1 and t[0] == True and s2[0] == 1  # noqa: E712
    assert rb.size() == 1


def test_other_agents_cannot_add():
    rb = ReplayBuffer("main", 4)
    rb.add("other", 1, 1, 1, False, 2)
    rb.start_new_episode("other")
    assert len(rb) == 0 and len(rb.episode_starting_indices) == 0


def test_fifo_replacement():
    rb = ReplayBuffer(0, 1)
    rb.add(0, 2, 2, 2, False, 2)
    rb.add(0, 1, 1, 1, True, 1)
    assert rb.size() == 1
    s, a, r, t, s2 = rb.sample_batch(1)
    assert s[0] == 1 and t[0] == True  # noqa: E712


def test_episode_markers():
    rb = ReplayBuffer(0, 1)
    rb.start_new_episode(0)
    rb.add(0, 2, 2, 2, False, 2)
    assert rb.next_episode_number == 1
    assert rb.episode_starting_indices[0] == 0 and len(rb.episode_starting_indices) == 1
    for i in range(1, 21):
        rb = ReplayBuffer(0, i)
        for _ in range(i):
            rb.start_new_episode(0)
            rb.add(0, 2, 2, 2, False, 2)
        assert rb.next_episode_number == i
        assert list(rb.episode_starting_indices) == list(range(i))


def test_episode_removal_and_reindexing():
    rb = ReplayBuffer(0, 1)
    rb.start_new_episode(0)
    rb.add(0, 2, 2, 2, False, 2)
    rb.add(0, 2, 2, 2, False, 2)
    assert rb.next_episode_number == 1 and len(rb.episode_starting_indices) == 0
    rb = ReplayBuffer(0, 2)
    rb.start_new_episode(0)
    rb.add(0, 2, 2, 2, False, 2)
    rb.start_new_episode(0)
    rb.add(0, 2, 2, 2, False, 2)
    rb.add(0, 2, 2, 2, False, 2)
    assert rb.next_episode_number == 2
    assert list(rb.episode_starting_indices) == [0]


def test_double_start_is_ignored(capsys):
    rb = ReplayBuffer(0, 1)
    rb.start_new_episode(0)
    rb.start_new_episode(0)
    assert list(rb.episode_starting_indices) == [0]
    rb = ReplayBuffer(0, 2)
    rb.start_new_episode(0)
    rb.add(0, 2, 2, 2, False, 2)
    rb.start_new_episode(0)
    assert list(rb.episode_starting_indices) == [0, 1]


def test_episode_number_to_buffer_index():
    rb = ReplayBuffer(0, 10)
    for i in range(20):
        rb.start_new_episode(0)
        rb.add(0, i, i, i, False, i)
        for j in range(len(rb.episode_starting_indices)):
            e = rb.episode_starting_indices[-(j + 1)]
            assert (i - j, i - j, i - j, False, i - j) == rb.buffer[rb.episode_number_to_buffer_index(e)]


def _fill(rb, main, episodes, steps, d, rng):
    for _ in range(episodes):
        rb.start_new_episode(main)
        s = rng.normal(size=d)
        for t in range(steps):
            s2 = s + rng.normal(size=d) * 0.1
            rb.add(main, s.copy(), rng.normal(size=1), 0.0, t == steps - 1, s2.copy())
            s = s2


@pytest.mark.parametrize("cap", [1000, 137])
def test_mirror_matches_deque(cap):
    """get_all_states / states_s2 / episodic paths from the contiguous ring equal the
    reference's list-comprehension definitions (replay_buffer.py:102,154-176,205)."""
    rng = np.random.default_rng(0)
    main = object()
    rb = ReplayBuffer(main, cap)
    _fill(rb, main, 9, 40, 3, rng)
    ref_all = np.array([st[0] for st in rb.buffer] + [rb.buffer[-1][4]])
    np.testing.assert_array_equal(rb.get_all_states(), ref_all)
    random.seed(3)
    idx = rb.get_possible_smart_start_indices(50)
    first = rb.episode_number_to_buffer_index(rb.episode_starting_indices[0])
    assert idx.min() >= first and len(set(idx.tolist())) == len(idx)
    np.testing.assert_array_equal(rb.states_s2(idx), np.array([rb.buffer[int(i)][4] for i in idx]))
    for i in idx[:10]:
        path = rb.get_episodic_path_to_buffer_index(int(i))
        assert np.array_equal(path[-1], rb.buffer[int(i)][4])
        # consecutive states of one episode: each s2 is the next s
        k = int(i) - (len(path) - 2)
        assert np.array_equal(path[0], rb.buffer[k][0])
        assert k == first or rb.buffer[k - 1][3] or rb.buffer_index_to_episode_number(k) in rb.episode_starting_indices


def test_no_episode_errors():
    rb = ReplayBuffer(0, 5)
    assert rb.get_possible_smart_start_indices(3) is None
    rb.add(0, 1.0, 0.0, 0.0, False, 2.0)
    with pytest.raises(ValueError):
        rb.get_episodic_path_to_buffer_index(0)


def test_candidate_draw_in_cxx_equals_random_sample():
    """csrc/py_random.cu restates random.sample(range(first, stop), k): same indices in the same order
    and the same global generator state afterwards, on both of Random.sample's branches."""
    import random

    from smartstartcontinuous_b200 import replay_buffer as rbm
    rbm._probe_fast_sample()
    assert rbm._FAST_SAMPLE, "the library's restatement disagrees with this interpreter's random.sample"
    for seed, (first, stop, k) in enumerate([(0, 100, 100), (5, 70, 65), (200, 100001, 16384), (0, 2001, 2000),
                                             (17, 5000, 4000), (3, 1000003, 65536), (0, 300, 65), (9, 4200, 1400)]):
        random.seed(seed)
        random.random()
        want = np.array(random.sample(range(first, stop), k))
        tail_want = [random.random(), random.getrandbits(40), random.gauss(0, 1)]
        random.seed(seed)
        random.random()
        got = rbm.sample_range(first, stop, k)
        tail_got = [random.random(), random.getrandbits(40), random.gauss(0, 1)]
        assert got.dtype == want.dtype and np.array_equal(got, want), (first, stop, k)
        assert tail_got == tail_want
    # the buffer method goes through it
    random.seed(3)
    a = random.sample(range(0, 500), 120)
    random.seed(3)
    assert rbm.sample_range(0, 500, 120).tolist() == a


def test_candidate_draw_random_shapes_both_routes():
    """Random (first, n, k, generator position): the block-wise set branch, the pool branch and the two ways
    of reaching the interpreter's generator (in place behind the object / getstate + setstate) against
    random.sample itself, including where the generator stops."""
    import random

    from smartstartcontinuous_b200 import replay_buffer as rbm
    rbm._probe_fast_sample()
    assert rbm._FAST_SAMPLE
    assert rbm._IN_PLACE, "CPython's RandomObject layout was not recognised"
    pick = np.random.default_rng(11)
    for trial in range(60):
        n = int(pick.choice([1, 2, 70, 625, 1300, 5000, 70000, 250000]))
        k = int(pick.integers(0, min(n, 20000) + 1))
        first = int(pick.integers(0, 1000))
        burn = int(pick.integers(0, 1300))
        for in_place in (True, False):
            a, b = random.Random(trial), random.Random(trial)
            for _ in range(burn):
                a.getrandbits(32), b.getrandbits(32)
            want = a.sample(range(first, first + n), k)
            got = rbm._sample_with(rbm._FAST_SAMPLE, b, first, n, k, in_place=in_place)
            assert got.tolist() == want, (n, k, first, burn, in_place)
            assert a.getstate() == b.getstate(), (n, k, first, burn, in_place)


def test_episodic_path_equals_the_per_step_construction():
    """get_episodic_path_to_buffer_index (replay_buffer.py:154-176) from the contiguous mirror = the
    reference's list built step by step from the deque, for every buffer index, before and after eviction."""
    from smartstartcontinuous_b200.replay_buffer import ReplayBuffer
    main = object()
    rb = ReplayBuffer(main, 230)
    rng = np.random.default_rng(5)
    added = 0
    for ep in range(9):
        rb.start_new_episode(main)
        for t in range(int(rng.integers(3, 60))):
            rb.add(main, rng.normal(size=3), rng.normal(size=1), 0.0, False, rng.normal(size=3))
            added += 1
        if ep in (3, 8):
            assert (added > 230) == (ep == 8)
            first_ok = rb.episode_number_to_buffer_index(rb.episode_starting_indices[0])
            for idx in range(first_ok, len(rb)):
                got = rb.get_episodic_path_to_buffer_index(idx)
                count = rb.buffer_index_to_episode_number(idx)
                start = max(e for e in rb.episode_starting_indices if e <= count)
                lo = rb.episode_number_to_buffer_index(start)
                want = [rb.buffer[i][0] for i in range(lo, idx + 1)] + [rb.buffer[idx][4]]
                assert len(got) == len(want)
                assert all(isinstance(g, np.ndarray) and np.array_equal(g, w) for g, w in zip(got, want))
            got[0][0] = 123.0                                   # a private copy, not a view of the mirror
            assert rb.get_episodic_path_to_buffer_index(len(rb) - 1)[0][0] != 123.0


def _golden_buffer(name, env, n_transitions, seed):
    """The replay buffer of oracle/make_golden.golden_kde rebuilt through OUR ReplayBuffer API (same seeded
    synthetic episodes, same capacity, so the same FIFO evictions)."""
    from smartstartcontinuous_b200 import synthetic as syn
    from smartstartcontinuous_b200.replay_buffer import ReplayBuffer
    rng = np.random.default_rng(seed)
    steps = 100
    if env == "pendulum":
        obs, act = syn.pendulum_rollouts(rng, n_transitions // steps, steps)
        episodes = [(obs[e], act[e]) for e in range(len(obs))]
    else:
        episodes = [syn.mountaincar_rollout(rng, steps) for _ in range(n_transitions // steps)]
    main = object()
    rb = ReplayBuffer(main, n_transitions - 37)
    for states, actions in episodes:
        rb.start_new_episode(main)
        T = len(actions)
        for t in range(T):
            rb.add(main, np.array(states[t]), np.array(actions[t]), 0.0, t == T - 1, np.array(states[t + 1]))
    return rb, episodes


@pytest.mark.parametrize("name,env,n,n_ss,seed", [("kde_pendulum.npz", "pendulum", 3000, 300, 0),
                                                  ("kde_mountaincar.npz", "mountaincar", 2000, 5000, 1)])
def test_candidate_indices_equal_the_references_from_the_seed(name, env, n, n_ss, seed):
    """From random.seed(seed) our buffer (ring + C++ restatement of random.sample) hands out the candidate
    indices the REFERENCE's buffer handed out when the golden was made, and the same data set."""
    import random

    from conftest import load_golden
    g = load_golden(name)
    rb, _ = _golden_buffer(name, env, n, seed)
    assert len(rb) == int(g["in_n_transitions"])
    np.testing.assert_array_equal(np.asarray(rb.get_all_states()), g["in_all_states"])
    random.seed(seed)
    idx = rb.get_possible_smart_start_indices(n_ss)
    np.testing.assert_array_equal(idx, g["in_indices"])
    np.testing.assert_array_equal(np.asarray(rb.states_s2(idx)), g["in_queries"])
