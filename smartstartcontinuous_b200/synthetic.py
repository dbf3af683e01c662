"""Seeded synthetic inputs for tests and bench (gym is not installed anywhere we run).

Everything here is plain numpy and deterministic.  The simulators restate the
environment equations the reference's configs run on:

* MountainCar: smartstart/environments/continuous_mountain_car_editted.py:60-82
  (power = 0.0015 * power_scalar, velocity clip 0.07, position clip [-1.2, 0.6]).
* Pendulum-v0 (gym classic control): th'' = -3g/(2l) sin(th+pi) + 3u/(ml^2), g=10, m=l=1,
  dt=.05, |th'|<=8, |u|<=2, observation (cos th, sin th, th').

The dynamics MLPs use the reference's initialiser (feedforward_network.py:8,14-23:
xavier_initializer(uniform=False) for weights *and* biases).
"""
from __future__ import annotations

import math

import numpy as np


# --------------------------------------------------------------------------- simulators
def mountaincar_rollout(rng, steps, power_scalar=1.0, start=None):
    """Random-policy rollout.  Returns states [steps+1, 2], actions [steps, 1] (float64)."""
    power = 0.0015 * power_scalar
    pos = rng.uniform(-0.6, -0.4) if start is None else float(start[0])
    vel = 0.0 if start is None else float(start[1])
    states = np.empty((steps + 1, 2))
    actions = rng.uniform(-1.0, 1.0, size=(steps, 1))
    states[0] = (pos, vel)
    for t in range(steps):
        force = min(max(actions[t, 0], -1.0), 1.0)
        vel += force * power - 0.0025 * math.cos(3 * pos)
        vel = min(max(vel, -0.07), 0.07)
        pos += vel
        pos = min(max(pos, -1.2), 0.6)
        if pos == -1.2 and vel < 0:
            vel = 0.0
        states[t + 1] = (pos, vel)
    return states, actions


def pendulum_rollouts(rng, episodes, steps):
    """Vectorised random-torque Pendulum-v0 episodes.

    Returns obs [episodes, steps+1, 3] and actions [episodes, steps, 1] (float64).
    """
    th = rng.uniform(-math.pi, math.pi, size=episodes)
    thdot = rng.uniform(-1.0, 1.0, size=episodes)
    obs = np.empty((episodes, steps + 1, 3))
    act = rng.uniform(-2.0, 2.0, size=(episodes, steps, 1))
    obs[:, 0] = np.stack([np.cos(th), np.sin(th), thdot], axis=1)
    g, m, l, dt = 10.0, 1.0, 1.0, 0.05
    for t in range(steps):
        u = np.clip(act[:, t, 0], -2.0, 2.0)
        thdot = thdot + (-3 * g / (2 * l) * np.sin(th + math.pi) + 3.0 / (m * l ** 2) * u) * dt
        th = th + thdot * dt
        thdot = np.clip(thdot, -8.0, 8.0)
        obs[:, t + 1] = np.stack([np.cos(th), np.sin(th), thdot], axis=1)
    return obs, act


def pendulum_buffer(n_transitions, seed=0, steps=200):
    """Replay-buffer contents for the KDE configs: n transitions in 200-step episodes.

    Returns (all_states [n+1, 3]  -- every s plus the last s2, as ReplayBuffer.get_all_states,
             s2 [n, 3], episode_starts [E]).
    """
    rng = np.random.default_rng(seed)
    episodes = -(-n_transitions // steps)
    obs, _ = pendulum_rollouts(rng, episodes, steps)
    s = obs[:, :-1].reshape(-1, 3)[:n_transitions]
    s2 = obs[:, 1:].reshape(-1, 3)[:n_transitions]
    all_states = np.concatenate([s, s2[-1:]], axis=0)
    starts = np.arange(0, n_transitions, steps)
    return np.ascontiguousarray(all_states), np.ascontiguousarray(s2), starts


def critic_like_values(states, seed=0):
    """float32 stand-in for DDPG get_state_value (DDPG_Baselines_agent.py:197-204):
    a fixed seeded 2-layer (64/32) tanh MLP, output (m,) float32."""
    rng = np.random.default_rng(seed)
    d = states.shape[1]
    w1 = rng.normal(0, 1 / math.sqrt(d), (d, 64))
    w2 = rng.normal(0, 1 / math.sqrt(64), (64, 32))
    w3 = rng.normal(0, 1 / math.sqrt(32), (32, 1))
    h = np.tanh(np.tanh(states @ w1) @ w2) @ w3
    return h[:, 0].astype(np.float32)


# --------------------------------------------------------------------------- dynamics MLP
def xavier_mlp(rng, d, da, num_fc_layers, depth, scale=1.0):
    """Weights/biases as the reference initialises them (truncated-normal Xavier, biases too).

    Returns (weights, biases): weights[i] is [in, out] float64 (y = x @ W + b).
    """
    sizes = [d + da] + [depth] * num_fc_layers + [d]
    weights, biases = [], []
    for fan_in, fan_out in zip(sizes[:-1], sizes[1:]):
        std = math.sqrt(1.3 * 2.0 / (fan_in + fan_out)) * scale

        def trunc(shape, s):
            x = rng.normal(0.0, s, size=shape)
            bad = np.abs(x) > 2 * s
            while bad.any():
                x[bad] = rng.normal(0.0, s, size=int(bad.sum()))
                bad = np.abs(x) > 2 * s
            return x

        weights.append(trunc((fan_in, fan_out), std))
        # tf xavier on a 1-D bias of length fan_out: fan_in = fan_out = len
        biases.append(trunc((fan_out,), math.sqrt(1.3 * 2.0 / (2 * fan_out)) * scale))
    return weights, biases


def normalisation_stats(states, actions):
    """mean/std of (s, a, s'-s) exactly as NND_MB_agent.py:302-315 computes them."""
    x = states[:-1]
    z = states[1:] - states[:-1]
    return dict(mean_x=x.mean(0), std_x=(x - x.mean(0)).std(0),
                mean_y=actions.mean(0), std_y=(actions - actions.mean(0)).std(0),
                mean_z=z.mean(0), std_z=(z - z.mean(0)).std(0))


def fit_dynamics_mlp(states_list, actions_list, num_fc_layers, depth, seed=0, epochs=30,
                     batch=512, lr=1e-3):
    """Small torch-CPU Adam/MSE fit (mirrors Dyn_Model.train hyper-parameters,
    dynamics_model.py:52-171) so bench/test fixtures have a *realistic* dynamics model.
    Returns (weights, biases, norm) in float64 numpy.
    """
    import torch

    xs = np.concatenate([s[:-1] for s in states_list])
    ys = np.concatenate(list(actions_list))
    zs = np.concatenate([s[1:] - s[:-1] for s in states_list])
    norm = dict(mean_x=xs.mean(0), std_x=(xs - xs.mean(0)).std(0),
                mean_y=ys.mean(0), std_y=(ys - ys.mean(0)).std(0),
                mean_z=zs.mean(0), std_z=(zs - zs.mean(0)).std(0))
    inp = np.concatenate([(xs - norm["mean_x"]) / norm["std_x"],
                          (ys - norm["mean_y"]) / norm["std_y"]], axis=1)
    out = (zs - norm["mean_z"]) / norm["std_z"]
    rng = np.random.default_rng(seed)
    w0, b0 = xavier_mlp(rng, xs.shape[1], ys.shape[1], num_fc_layers, depth)
    tw = [torch.tensor(w, dtype=torch.float64, requires_grad=True) for w in w0]
    tb = [torch.tensor(b, dtype=torch.float64, requires_grad=True) for b in b0]
    opt = torch.optim.Adam(tw + tb, lr=lr)
    tin, tout = torch.tensor(inp), torch.tensor(out)
    g = torch.Generator().manual_seed(seed)
    for _ in range(epochs):
        perm = torch.randperm(len(tin), generator=g)
        for i in range(0, len(tin) - batch + 1, batch):
            idx = perm[i:i + batch]
            h = tin[idx]
            for li, (w, b) in enumerate(zip(tw, tb)):
                h = h @ w + b
                if li != len(tw) - 1:
                    h = torch.relu(h)
            loss = ((h - tout[idx]) ** 2).mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
    return ([w.detach().numpy().copy() for w in tw], [b.detach().numpy().copy() for b in tb], norm)
