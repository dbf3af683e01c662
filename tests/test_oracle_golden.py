"""The oracle (numpy restatement) against golden vectors produced by the reference's own code
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden_model, load_golden
from oracle import kde_oracle, mpc_oracle

KDE_CASES = ["kde_pendulum.npz", "kde_mountaincar.npz"]
MPC_CASES = ["mpc_mountaincar_L2.npz", "mpc_pendulum_L1.npz", "mpc_mountaincar_L3_xavier.npz",
             "mpc_pendulum_2x500.npz"]


@pytest.mark.parametrize("name", KDE_CASES)
def test_kde_density_matches_reference(name):
    g = load_golden(name)
    dens = kde_oracle.kde_density(g["in_all_states"], g["in_queries"])
    np.testing.assert_allclose(dens, g["out_density"], rtol=1e-10)
    dens2 = kde_oracle.kde_density_direct(g["in_all_states"], g["in_queries"][:40])
    np.testing.assert_allclose(dens2, g["out_density"][:40], rtol=1e-11)
    dens3 = kde_oracle.scipy_density(g["in_all_states"], g["in_queries"])
    np.testing.assert_allclose(dens3, g["out_density"], rtol=1e-12)


@pytest.mark.parametrize("name", KDE_CASES)
def test_ucb_and_argmax_match_reference(name):
    g = load_golden(name)
    best, dens, ucb = kde_oracle.select_start(
        g["in_all_states"], g["in_queries"], g["in_values"], int(g["in_n_transitions"]),
        float(g["in_volume"]), float(g["in_alpha"]), float(g["in_beta"]))
    np.testing.assert_allclose(ucb, g["out_ucb"], rtol=1e-10)
    assert best == int(g["out_best_j"])
    assert int(g["in_indices"][best]) == int(g["out_chosen_buffer_index"])


def test_volume_matches_reference():
    g = load_golden("kde_pendulum.npz")
    assert kde_oracle.hyperellipsoid_volume(g["in_radii"]) == pytest.approx(float(g["in_volume"]), rel=1e-14)


@pytest.mark.parametrize("name", MPC_CASES)
def test_forward_sim_matches_reference(name):
    g = load_golden(name)
    w, b, norm = golden_model(g)
    states = mpc_oracle.forward_sim(g["in_start_state"], g["in_actions"], w, b, norm)
    np.testing.assert_allclose(states, g["out_states"], rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("name", MPC_CASES)
def test_scores_and_choice_match_reference(name):
    g = load_golden(name)
    w, b, norm = golden_model(g)
    scores = mpc_oracle.score_add_delta(g["out_states"], g["out_desired_states"], g["out_distances_left"],
                                        g["out_radii"], int(g["in_wp_index"]), float(g["in_gamma"]),
                                        float(g["in_hpf"]))
    np.testing.assert_allclose(scores, g["out_scores"], rtol=1e-10, atol=1e-11)
    res = mpc_oracle.plan(g["in_start_state"], g["in_actions"], w, b, norm, g["out_desired_states"],
                          g["out_distances_left"], g["out_radii"], int(g["in_wp_index"]),
                          float(g["in_gamma"]), float(g["in_hpf"]))
    assert res["best_k"] == int(g["out_best_k"])
    np.testing.assert_allclose(res["best_path"], g["out_best_path"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(res["best_action"], g["out_best_action"])


def test_per_sample_penalty_differs_from_reference_quirk():
    """Q1: the reference's projection coefficient is global over the batch; the per-sample
    variant is a different (documented) scorer."""
    g = load_golden("mpc_mountaincar_L2.npz")
    args = (g["out_states"], g["out_desired_states"], g["out_distances_left"], g["out_radii"],
            int(g["in_wp_index"]), float(g["in_gamma"]), float(g["in_hpf"]))
    ref = mpc_oracle.score_add_delta(*args, penalty_mode=mpc_oracle.PENALTY_REFERENCE)
    per = mpc_oracle.score_add_delta(*args, penalty_mode=mpc_oracle.PENALTY_PER_SAMPLE)
    assert np.abs(ref - per).max() > 1e-3
    # a single-sample batch makes both coincide
    one = tuple([g["out_states"][:, 5:6]]) + args[1:]
    np.testing.assert_allclose(mpc_oracle.score_add_delta(*one, penalty_mode=0),
                               mpc_oracle.score_add_delta(*one, penalty_mode=1), rtol=1e-12)
