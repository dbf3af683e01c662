"""Host-side logic of the drop-in agents that needs no GPU."""
import numpy as np
import numpy.random as npr
import pytest

from smartstartcontinuous_b200.distributed import ShardedPlanner


@pytest.mark.parametrize("low,high,shape", [([-2.0], [2.0], (300, 50, 1)), ([-1.0, 0.3], [1.0, 2.7], (77, 7, 2)),
                                            ([0.0, -5.0, 1e-3], [1e-9, 5.0, 1e3], (64, 3, 3))])
def test_host_sampling_is_bit_identical_to_npr_uniform(low, high, shape):
    """NND_MB_agent.get_best_sim_actions draws low + (high - low) * random_sample(shape) in place of
    npr.uniform(low, high, shape) (NND_MB_agent.py:500-501): same MT19937 stream, same bits."""
    low, high = np.asarray(low), np.asarray(high)
    npr.seed(5)
    want = npr.uniform(low, high, shape)
    npr.seed(5)
    got = npr.random_sample(shape)
    got *= high - low
    got += low
    np.testing.assert_array_equal(got, want)
    assert npr.random_sample() == npr.RandomState(5).random_sample(int(np.prod(shape)) + 1)[-1]


def test_sharded_planner_rejects_empty_shards():
    class Dummy:
        _model_shape = (2, 1, 2, 8)

        def rollout(self, *a, **k):
            raise AssertionError("no work may be queued for an unshardable batch")

    p = ShardedPlanner(Dummy(), device="cpu", tensors=object())
    p.world, p.rank = 4, 1
    with pytest.raises(ValueError):
        p.plan(np.zeros(2), 0, K=3, H=5)


def test_reference_harness_sources():
    """The CPU arm of bench.py runs the reference's own code: either /root/reference is mounted or
    oracle/stage_ref.py has packed it; when neither is there the arm falls back to the oracle port."""
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("no reference tree and no staged archive here")
    ref = rh.load_reference()
    assert hasattr(ref.nnd.NND_MB_agent, "get_best_sim_actions")
    assert hasattr(ref.ssc.SmartStartContinuous, "get_smart_start_path")


def test_numpy_global_generator_state_is_reachable_in_place():
    """The default agent path hands the library the address of numpy's global MT19937 state struct
    ({uint32 key[624]; int pos}, BitGenerator.ctypes.state_address) instead of paying get_state() +
    set_state() per decision: what is read there is what get_state() reports, through draws, re-seeding and
    set_state, and a state written there is what numpy continues from."""
    import ctypes as C

    from smartstartcontinuous_b200.engine import _NumpyGlobalMT

    addr = _NumpyGlobalMT.address()
    assert addr is not None

    def view():
        return np.ctypeslib.as_array((C.c_uint32 * 624).from_address(addr)), C.c_int.from_address(addr + 2496)

    saved = np.random.get_state()
    try:
        for prepare in (lambda: np.random.seed(123), lambda: np.random.random_sample(777),
                        lambda: np.random.set_state(saved), lambda: np.random.randint(0, 10, 3)):
            prepare()
            assert _NumpyGlobalMT.address() == addr
            key, pos = view()
            st = np.random.get_state()
            assert np.array_equal(key, st[1]) and pos.value == st[2]
        # advance a private generator, write its state into the struct: the global stream continues from there
        rs = np.random.RandomState(5)
        rs.random_sample(1000)
        st = rs.get_state()
        key, pos = view()
        key[:] = st[1]
        pos.value = st[2]
        assert np.array_equal(np.random.random_sample(10), rs.random_sample(10))
    finally:
        np.random.set_state(saved)
