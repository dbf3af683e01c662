"""Host geometry helpers: the reference's own known-answer tests
(tests/utilities/test_numerical.py:8-103 in the reference) restated against
smartstartcontinuous_b200.numerical, plus the plan set-up pinned by golden vectors."""
import numpy as np
import pytest

from conftest import load_golden
from smartstartcontinuous_b200 import numerical as num


def test_std_zero_for_constant_and_alternating_paths():
    assert np.equal([0], num.path_deltas_stds_and_means_per_dim([[1], [1], [1], [1]])[0])
    assert np.equal([0, 0], num.path_deltas_stds_and_means_per_dim([[1, 1]] * 4)[0]).all()
    assert np.equal([0, 0], num.path_deltas_stds_and_means_per_dim([[1, 10], [0, 9], [1, 8], [0, 9]])[0]).all()
    assert num.path_deltas_stds_and_means_per_dim([[1], [2], [4]])[0][0] == .50


def test_means():
    for i in range(1, 50):
        expected = i + 50
        path = [[0]]
        for j in range(expected - i, expected + i + 1):
            path.append([path[-1][0] + j])
        assert num.path_deltas_stds_and_means_per_dim(path)[1][0] == expected


def test_short_path_raises():
    with pytest.raises(ValueError):
        num.path_deltas_stds_and_means_per_dim([[1.0, 2.0]])


def test_projection_single_vectors():
    assert np.equal(num.projection_of_a_onto_b(np.array([1, 1]), np.array([0, 1])), np.array([0, 1])).all()
    assert np.allclose(num.projection_of_a_onto_b(np.array([1, 1, 1]), np.array([0, 1, 1])), [0, 1, 1])


def test_projection_batched_coefficient_is_global():
    """Q1: for [K, d] batches the reference's coefficient is one scalar for the whole batch."""
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=(5, 2)), rng.normal(size=(5, 2))
    lam = (a * b).sum() / (b * b).sum()
    np.testing.assert_allclose(num.projection_of_a_onto_b(a, b), lam * b)


@pytest.mark.parametrize("stretch", [1, 10])
def test_dist_line_seg_to_point(stretch):
    radii = [1, stretch]
    f = num.elliptical_euclidean_distance_function_generator(radii)
    d = num.dist_line_seg_to_point(np.array([0, 1 * stretch]), np.array([1, 2 * stretch]),
                                   np.array([2, 1 * stretch]), f, radii)
    assert np.isclose(d, 2 ** .5)


def test_radii_must_be_positive():
    with pytest.raises(AssertionError):
        num.elliptical_euclidean_distance_function_generator([1.0, 0.0])


def test_volumes():
    for r in range(1, 20):
        assert num.volume_of_n_dimensional_hyperellipsoid([r, r]) == pytest.approx(np.pi * r ** 2, rel=1e-15)
        assert abs(num.volume_of_n_dimensional_hyperellipsoid([r, r, r]) - 4 / 3 * np.pi * r ** 3) < 1e-10
    for r1 in range(1, 20, 3):
        for r2 in range(1, 20, 3):
            assert num.volume_of_n_dimensional_hyperellipsoid([r1, r2]) == pytest.approx(np.pi * r1 * r2, rel=1e-15)
            for r3 in range(1, 20, 3):
                assert abs(num.volume_of_n_dimensional_hyperellipsoid([r1, r2, r3]) - 4 / 3 * np.pi * r1 * r2 * r3) < 1e-10


def test_binary_search_index_lower():
    arr = list(range(100))
    for i in range(100):
        assert i == num.binary_search_index_lower(arr, i)
        assert i == num.binary_search_index_lower(arr, i + .5)
        assert i == num.binary_search_index_lower(arr, i + .75)
    assert 99 == num.binary_search_index_lower(arr, 1000)
    assert num.binary_search_index_lower(arr, -1) is None


def test_length_weighted_activities_solver():
    assert num.length_weighted_activities_solver([[1, 4], [2, 8], [3, 11], [5, 7], [8, 15], [13, 18]])[0] == 13
    assert num.length_weighted_activities_solver([]) == (0, [])


def test_path_shortcutter():
    f = num.elliptical_euclidean_distance_function_generator([1, 1])
    assert np.equal(num.path_shortcutter([[0, 0], [1, 1], [2, 2], [3, 3], [1, 1]], f, 1),
                    [[0, 0], [1, 1], [1, 1]]).all()
    path = [[0, 0], [1, 1], [2, 2], [3, 3], [4, 4]]
    assert np.equal(num.path_shortcutter(path, f, 1), path).all()


def test_waypoints():
    path = [[i, 0] for i in range(10)]
    assert num.get_start_waypoints_final_states_steps(path, 3) == [[0, 0], [3, 0], [6, 0], [9, 0]]
    assert num.get_start_waypoints_final_states_steps(np.array(path), 1) == path


@pytest.mark.parametrize("name", ["mpc_mountaincar_L2.npz", "mpc_pendulum_L1.npz",
                                  "mpc_mountaincar_L3_xavier.npz", "mpc_pendulum_2x500.npz"])
def test_plan_setup_matches_reference(name):
    """radii / shortcut path / waypoints / distances_left as the reference's
    start_new_episode_plan produced them (NND_MB_agent.py:375-418)."""
    from smartstartcontinuous_b200.nnd_mb_agent import plan_from_path
    g = load_golden(name)
    plan = plan_from_path(list(g["in_path"]), mean_per_stepsize=1, std_per_stepsize=1,
                          stepsizes_in_waypoint_radii=1, path_shortcutting=True, theta=1,
                          steps_per_waypoint=1)
    np.testing.assert_allclose(plan["radii"], g["out_radii"], rtol=1e-13)
    np.testing.assert_array_equal(plan["path_to_follow"], g["out_path_to_follow"])
    np.testing.assert_array_equal(plan["desired_states"], g["out_desired_states"])
    np.testing.assert_allclose(plan["distances_left"], g["out_distances_left"], rtol=1e-12)
