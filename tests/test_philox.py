import numpy as np

from oracle.philox import kat_vectors, philox4x32_10, sample_actions


def test_random123_known_answers():
    for ctr, key, exp in kat_vectors:
        out = philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert out.tolist() == list(exp)


def test_sampler_is_shard_invariant_and_in_range():
    full = sample_actions(64, 7, 2, 99, [-1.0, 0.0], [1.0, 3.0])
    part = sample_actions(16, 7, 2, 99, [-1.0, 0.0], [1.0, 3.0], k_offset=32)
    np.testing.assert_array_equal(full[32:48], part)
    assert full[..., 0].min() >= -1.0 and full[..., 0].max() < 1.0
    assert full[..., 1].min() >= 0.0 and full[..., 1].max() < 3.0
    assert abs(full[..., 0].mean()) < 0.15
