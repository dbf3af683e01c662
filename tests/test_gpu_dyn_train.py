"""Row f1 on the B200 through the C ABI: the device trainer (csrc/dyn_train.cu) against the float64
numpy oracle of Dyn_Model.train's optimisation step -- identical initial parameters, identical
batches, weights compared after N Adam steps."""
import numpy as np
import numpy.random as npr
import pytest

from oracle import dyn_train_oracle as dto
from oracle import mpc_oracle
from smartstartcontinuous_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _assert_params_close(got_list, want_list, init_list):
    """FP32 device arithmetic against the float64 oracle, measured on how far Adam moved each tensor.
    The same algorithm run in float32 numpy shows: relative Frobenius error of the movement <= 7e-4,
    while single elements whose gradient hovers around the epsilon of Adam's denominator (weights
    of rarely active ReLU units) drift by up to 4 % of the largest update -- hence a tight norm bound
    and a loose element bound."""
    for got, want, init in zip(got_list, want_list, init_list):
        moved = np.abs(want - init).max()
        assert moved > 0
        assert np.linalg.norm(got - want) <= 3e-3 * np.linalg.norm(want - init), \
            (np.linalg.norm(got - want), np.linalg.norm(want - init))
        assert np.abs(got - want).max() <= 0.1 * moved + 1e-6, (np.abs(got - want).max(), moved)
        assert np.median(np.abs(got - want)) <= 5e-4 * moved + 1e-7


def _setup(engine, seed, d, da, L, h, n_old, n_new):
    rng = np.random.default_rng(seed)
    w, b = syn.xavier_mlp(rng, d, da, L, h)
    norm = dict(mean_x=np.zeros(d), std_x=np.ones(d), mean_y=np.zeros(da), std_y=np.ones(da), mean_z=np.zeros(d),
                std_z=np.ones(d))
    X = rng.normal(size=(n_old + n_new, d + da))
    Z = np.tanh(X[:, :d] * 0.7 + 0.3 * X[:, d:d + 1]) + 0.05 * rng.normal(size=(n_old + n_new, d))
    engine.set_model(w, b, norm)
    engine.dyn_reset_optimizer()
    engine.dyn_set_data(0, X[:n_old], Z[:n_old])
    engine.dyn_set_data(1, X[n_old:], Z[n_old:])
    return w, b, norm, X, Z


@pytest.mark.parametrize("cfg", [(3, 1, 2, 500, 512, 0.9, 40), (2, 1, 1, 32, 512, 0.9, 120), (3, 1, 3, 64, 200, 0.5, 60),
                                 (2, 1, 2, 500, 512, 1.0, 12)])
def test_device_training_matches_float64_oracle(engine, cfg):
    """The BASELINE network (2x500, batch 512: 52 old + 460 new rows per batch), the example's default
    (1x32), a 3-layer case with a ragged batch, and the new-data-only branch (fraction 1)."""
    d, da, L, h, batch, frac, steps = cfg
    n_old, n_new = 3000, 1500
    w, b, norm, X, Z = _setup(engine, 5, d, da, L, h, n_old, n_new)
    ow, ob = [a.copy() for a in w], [a.copy() for a in b]
    state = dto.AdamState(ow, ob)
    npr.seed(11)
    if frac < 1.0:
        io, inw = dto.epoch_batches(n_old, n_new, batch, frac)
    else:
        nb = n_new // batch
        io, inw = np.empty((nb, 0), dtype=np.int64), np.arange(nb * batch).reshape(nb, batch)
    io, inw = io[:steps], inw[:steps]
    want_losses = dto.train_batches(ow, ob, state, X[:n_old], Z[:n_old], X[n_old:], Z[n_old:], io, inw, 1e-3)
    got_losses = engine.dyn_train_batches(io, inw, 1e-3)
    np.testing.assert_allclose(got_losses, want_losses, rtol=2e-4)
    gw, gb = engine.dyn_get_params()
    _assert_params_close(gw + gb, ow + ob, w + b)
    # validation loss of the trained model (dynamics_model.py:139-166)
    loss_dev, nb = engine.dyn_eval_loss(0, batch)
    want = np.mean([dto.loss_and_grads(X[i * batch:(i + 1) * batch], Z[i * batch:(i + 1) * batch], ow, ob)[0]
                    for i in range(n_old // batch)])
    assert nb == n_old // batch
    assert loss_dev == pytest.approx(want, rel=2e-4)


def test_adam_state_persists_and_commit_feeds_the_planner(engine):
    """Two training calls continue one optimisation (the moments and the step count live on the device,
    like the optimizer slots of the reference's graph), and after ss_dyn_commit the rollout kernels --
    FP32 and tcgen05 -- use the trained parameters without a host round trip."""
    d, da, L, h = 3, 1, 2, 500
    w, b, norm, X, Z = _setup(engine, 7, d, da, L, h, 2000, 1000)
    ow, ob = [a.copy() for a in w], [a.copy() for a in b]
    state = dto.AdamState(ow, ob)
    npr.seed(2)
    io, inw = dto.epoch_batches(2000, 1000, 512, 0.9)
    dto.train_batches(ow, ob, state, X[:2000], Z[:2000], X[2000:], Z[2000:], io[:20], inw[:20], 1e-3)
    engine.dyn_train_batches(io[:8], inw[:8], 1e-3, want_losses=False)
    engine.dyn_train_batches(io[8:20], inw[8:20], 1e-3, want_losses=False)
    gw, gb = engine.dyn_get_params()
    _assert_params_close(gw + gb, ow + ob, w + b)
    engine.dyn_commit()
    from smartstartcontinuous_b200.nnd_mb_agent import plan_from_path
    obs, _ = syn.pendulum_rollouts(np.random.default_rng(0), 1, 80)
    plan = plan_from_path(list(obs[0][:60]), mean_per_stepsize=1, std_per_stepsize=1, stepsizes_in_waypoint_radii=1,
                          path_shortcutting=True, theta=1, steps_per_waypoint=1)
    engine.set_plan(plan["desired_states"], plan["distances_left"], plan["radii"])
    acts = np.random.RandomState(3).uniform(-2, 2, (700, 6, 1))
    o = mpc_oracle.plan(obs[0][0], acts, gw, gb, norm, plan["desired_states"], plan["distances_left"], plan["radii"],
                        0, .75, .5)
    o_init = mpc_oracle.plan(obs[0][0], acts, w, b, norm, plan["desired_states"], plan["distances_left"],
                             plan["radii"], 0, .75, .5)
    assert np.abs(o["scores"] - o_init["scores"]).max() > 1e-2          # training changed the predictions
    for prec, tol in (("fp32", 1e-4), ("bf16_tc", 5e-2)):
        res = engine.plan(obs[0][0], 0, actions=acts, precision=prec, want_scores=True)
        bad = np.abs(res["scores"] - o["scores"]) > tol * np.maximum(np.abs(o["scores"]), 1.0)
        assert bad.mean() <= 0.01, (prec, bad.mean())


def test_dyn_model_train_api(engine):
    """Dyn_Model.train (the reference's signature) end to end: the loss falls, the three returned
    numbers are the reference's, and do_forward_sim uses the trained model."""
    from smartstartcontinuous_b200.dynamics_model import Dyn_Model
    rng = np.random.default_rng(3)
    obs, act = syn.pendulum_rollouts(rng, 12, 200)
    xs = np.concatenate([o[:-1] for o in obs]); ys = np.concatenate(list(act)); zs = np.concatenate([o[1:] - o[:-1] for o in obs])
    st = {k: v for k, v in zip(("mean_x", "mean_y", "mean_z"), (xs.mean(0), ys.mean(0), zs.mean(0)))}
    sd = {k: v for k, v in zip(("std_x", "std_y", "std_z"), (xs.std(0), ys.std(0), zs.std(0)))}
    inputs = np.concatenate([(xs - st["mean_x"]) / sd["std_x"], (ys - st["mean_y"]) / sd["std_y"]], axis=1)
    outputs = (zs - st["mean_z"]) / sd["std_z"]
    m = Dyn_Model(4, 3, None, 1e-3, 512, 2, 64, st["mean_x"], st["mean_y"], st["mean_z"], sd["std_x"], sd["std_y"],
                  sd["std_z"], "float64", False, engine=engine, seed=0)
    engine.dyn_reset_optimizer()
    npr.seed(0)
    before = m.run_validation(inputs[:1024], outputs[:1024], quiet=True)
    tr, old_loss, new_loss = m.train(inputs[:1800], outputs[:1800], inputs[1800:], outputs[1800:], 6, None, 0.9,
                                     save_results=False)
    after = m.run_validation(inputs[:1024], outputs[:1024], quiet=True)
    assert after < 0.5 * before and tr > 0 and old_loss > 0 and new_loss > 0
    w, b = m.export()
    acts = rng.uniform(-2, 2, (5, 7, 1))
    states = np.stack(m.do_forward_sim([obs[0][0], 0], acts, True))
    want = mpc_oracle.forward_sim(obs[0][0], acts, w, b, m.norm())
    np.testing.assert_allclose(states, want, rtol=1e-4, atol=1e-5)
