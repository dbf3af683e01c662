"""CPU oracle for the dynamics-model training step (float64 numpy).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Restates what one ``sess.run([train_step, mse_])`` of the reference does
(dynamics_model.py:36-50 builds it, :52-131 feeds it):

  * forward pass of feedforward_network.py:3-23 (Linear + ReLU hidden layers, linear output,
    y = x W + b with W [in, out]);
  * loss  mse = reduce_mean(square(z - f(x)))  over batch x outputs        (dynamics_model.py:42);
  * gradients of that loss with respect to every kernel and bias (what
    AdamOptimizer.compute_gradients returns, :45-49);
  * tf.train.AdamOptimizer(learning_rate) with its defaults beta1 = 0.9, beta2 = 0.999,
    epsilon = 1e-8 (:45).  TensorFlow 1.5 (Pipfile.lock:850) is a third-party dependency that cannot
    be installed here; its published update rule is
        t <- t + 1;  lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
        m <- beta1 m + (1 - beta1) g;  v <- beta2 v + (1 - beta2) g^2
        theta <- theta - lr_t * m / (sqrt(v) + epsilon)
    ("epsilon hat" form: epsilon is NOT scaled by the bias correction, unlike torch.optim.Adam);
  * the batching rule of Dyn_Model.train (:60-131): see ``epoch_batches``.

Pinning: the reference holds no test or fixture for training and TensorFlow cannot run, so the
optimizer arithmetic is **parity unpinned** by reference outputs; the forward / backward pass is
cross-checked against torch autograd in float64 (tests/test_dyn_train_oracle.py), and the Adam rule
against a torch.optim.Adam run with the equivalent epsilon.
"""
from __future__ import annotations

import math

import numpy as np
import numpy.random as npr

BETA1, BETA2, EPSILON = 0.9, 0.999, 1e-8


def forward(x, weights, biases):
    """Returns the list of layer inputs [x, h1, ..., hL] and the network output."""
    acts = [np.asarray(x, dtype=np.float64)]
    last = len(weights) - 1
    for i, (w, b) in enumerate(zip(weights, biases)):
        y = acts[-1] @ w + b
        if i != last:
            acts.append(np.maximum(y, 0.0))
        else:
            return acts, y


def loss_and_grads(x, z, weights, biases):
    """mse and d mse / d (W_l, b_l) for one batch."""
    acts, out = forward(x, weights, biases)
    z = np.asarray(z, dtype=np.float64)
    diff = out - z
    loss = float(np.mean(diff * diff))
    dy = 2.0 * diff / diff.size
    gw, gb = [None] * len(weights), [None] * len(weights)
    for l in range(len(weights) - 1, -1, -1):
        gw[l] = acts[l].T @ dy
        gb[l] = dy.sum(axis=0)
        if l > 0:
            dy = (dy @ weights[l].T) * (acts[l] > 0.0)
    return loss, gw, gb


class AdamState:
    def __init__(self, weights, biases):
        self.t = 0
        self.mw = [np.zeros_like(w, dtype=np.float64) for w in weights]
        self.vw = [np.zeros_like(w, dtype=np.float64) for w in weights]
        self.mb = [np.zeros_like(b, dtype=np.float64) for b in biases]
        self.vb = [np.zeros_like(b, dtype=np.float64) for b in biases]


def adam_step(weights, biases, gw, gb, state, lr):
    """In-place tf.train.AdamOptimizer update of float64 parameter lists."""
    state.t += 1
    lr_t = lr * math.sqrt(1.0 - BETA2 ** state.t) / (1.0 - BETA1 ** state.t)
    for params, grads, ms, vs in ((weights, gw, state.mw, state.vw), (biases, gb, state.mb, state.vb)):
        for p, g, m, v in zip(params, grads, ms, vs):
            m *= BETA1
            m += (1.0 - BETA1) * g
            v *= BETA2
            v += (1.0 - BETA2) * g * g
            p -= lr_t * m / (np.sqrt(v) + EPSILON)


def train_batches(weights, biases, state, X_old, Z_old, X_new, Z_new, idx_old, idx_new, lr):
    """Adam steps over explicit batches: batch i = old rows idx_old[i] then new rows idx_new[i].
    Parameters are updated in place; returns the per-batch losses (before each update)."""
    losses = []
    for io, inw in zip(idx_old, idx_new):
        xb = np.concatenate([X_old[io], X_new[inw]]) if len(inw) else X_old[io]
        zb = np.concatenate([Z_old[io], Z_new[inw]]) if len(inw) else Z_old[io]
        if len(io) == 0:
            xb, zb = X_new[inw], Z_new[inw]
        loss, gw, gb = loss_and_grads(xb, zb, weights, biases)
        adam_step(weights, biases, gw, gb, state, lr)
        losses.append(loss)
    return losses


def epoch_batches(n_old, n_new, batchsize, fraction_use_new):
    """Row indices of one epoch of Dyn_Model.train (dynamics_model.py:60-106), drawn from the global
    numpy stream with the reference's calls in the reference's order.  Returns (idx_old [nb, n_o],
    idx_new [nb, n_n]) for the mixed branch (old rows per batch > 0), the only one this helper
    covers."""
    n_n = n_new if n_new < batchsize * fraction_use_new else int(batchsize * fraction_use_new)
    n_o = int(batchsize - n_n)
    assert n_o > 0
    order = npr.choice(np.arange(n_old), size=(n_old,), replace=False)          # :76
    nb = int(math.floor(n_old / n_o))
    idx_old = np.empty((nb, n_o), dtype=np.int64)
    idx_new = np.empty((nb, n_n), dtype=np.int64)
    for b in range(nb):
        if n_new:
            idx_new[b] = npr.randint(0, n_new, (n_n,))                          # :88
        idx_old[b] = order[b * n_o:(b + 1) * n_o]                               # :93
    return idx_old, idx_new
