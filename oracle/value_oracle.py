"""TEST INFRASTRUCTURE (not product code): float32 numpy restatement of the DDPG value batch
agent.get_state_value feeds into the UCB (smartexplorationcontinuous.py:274).

Follows the reference graph: observation normalisation + clip (ddpg_editted.py:106-107), Actor_Editted
(models_editted.py:38-58), Critic_Editted (:81-99, action concatenated after the first hidden layer),
clip + denormalise of the critic output (ddpg_editted.py:130-131), tf.contrib.layers.layer_norm
(biased variance over the units, epsilon 1e-12, centre and scale).

Parity unpinned: TensorFlow 1.5 / baselines are not installable here, so no reference output exists
for this function; the restatement is checked only against the reference source by reading.
"""
import numpy as np

F = np.float32


def _ln(x, g, b):
    m = x.mean(axis=-1, keepdims=True, dtype=F)
    v = ((x - m) ** 2).mean(axis=-1, keepdims=True, dtype=F)
    return (x - m) / np.sqrt(v + F(1e-12)) * g + b


def _mlp_head(x, layers, ln, last_tanh, action=None):
    (W1, b1), (W2, b2), (W3, b3) = [(np.asarray(W, F), np.asarray(b, F)) for W, b in layers]
    h = x @ W1 + b1
    if ln is not None:
        h = _ln(h, np.asarray(ln[0][0], F), np.asarray(ln[0][1], F))
    h = np.maximum(h, F(0))
    if action is not None:
        h = np.concatenate([h, action], axis=-1)
    h = h @ W2 + b2
    if ln is not None:
        h = _ln(h, np.asarray(ln[1][0], F), np.asarray(ln[1][1], F))
    h = np.tanh(h) if last_tanh else np.maximum(h, F(0))
    return h @ W3 + b3


def state_values(queries, net):
    x = np.asarray(queries, dtype=np.float64)
    lo, hi = net.get("obs_clip", (-5.0, 5.0))          # the clip is unconditional (ddpg_editted.py:106-109)
    if net.get("obs_mean") is not None:
        x = (x.astype(F) - np.asarray(net["obs_mean"], F)) * (F(1) / np.asarray(net["obs_std"], np.float64)).astype(F)
    x = np.clip(x.astype(F), F(lo), F(hi))
    last_tanh = bool(net.get("last_layer_tanh", False))
    action = np.tanh(_mlp_head(x, net["actor"], net.get("actor_ln"), last_tanh))
    v = _mlp_head(x, net["critic"], net.get("critic_ln"), last_tanh, action=action)[:, 0]
    lo, hi = net.get("ret_clip", (-np.inf, np.inf))    # unconditional too (:130-131); denormalize only with ret_rms
    v = np.clip(v, F(lo), F(hi))
    if net.get("ret_mean") is not None:
        v = v * F(net["ret_std"]) + F(net["ret_mean"])
    return v.astype(F)
