"""The training oracle (oracle/dyn_train_oracle.py) against torch autograd / torch.optim.Adam in
float64 (CPU).  TensorFlow itself cannot run here, so this is what pins the restatement of the loss,
its gradients and the Adam rule (with TF's epsilon placement mapped onto torch's)."""
import math

import numpy as np
import numpy.random as npr
import pytest
import torch

from oracle import dyn_train_oracle as dto
from smartstartcontinuous_b200 import synthetic as syn


def _problem(seed, din, dout, L, h, n):
    rng = np.random.default_rng(seed)
    w, b = syn.xavier_mlp(rng, dout, din - dout, L, h)
    X = rng.normal(size=(n, din))
    Z = rng.normal(size=(n, dout))
    return w, b, X, Z


@pytest.mark.parametrize("shape", [(4, 3, 2, 50), (3, 2, 1, 32), (5, 3, 3, 20)])
def test_loss_and_gradients_match_autograd(shape):
    din, dout, L, h = shape
    w, b, X, Z = _problem(1, din, dout, L, h, 64)
    loss, gw, gb = dto.loss_and_grads(X, Z, w, b)
    tw = [torch.tensor(a, requires_grad=True) for a in w]
    tb = [torch.tensor(a, requires_grad=True) for a in b]
    hcur = torch.tensor(X)
    for i, (a, c) in enumerate(zip(tw, tb)):
        hcur = hcur @ a + c
        if i != L:
            hcur = torch.relu(hcur)
    tl = ((torch.tensor(Z) - hcur) ** 2).mean()
    tl.backward()
    assert loss == pytest.approx(float(tl), rel=1e-13)
    for g, t in zip(gw + gb, tw + tb):
        np.testing.assert_allclose(g, t.grad.numpy(), rtol=1e-10, atol=1e-14)


def test_adam_rule_matches_torch_with_equivalent_epsilon():
    """TF: theta -= lr_t m / (sqrt(v) + eps), lr_t = lr sqrt(1-b2^t)/(1-b1^t).
    torch: theta -= lr/(1-b1^t) m / (sqrt(v)/sqrt(1-b2^t) + eps').  Equal when
    eps' = eps / sqrt(1-b2^t); torch's eps is fixed, so compare step by step with a fresh
    single-step torch optimizer state transplanted from the oracle."""
    w, b, X, Z = _problem(2, 4, 3, 2, 30, 128)
    state = dto.AdamState(w, b)
    lr = 1e-3
    for step in range(1, 6):
        loss, gw, gb = dto.loss_and_grads(X, Z, w, b)
        before = [a.copy() for a in w + b]
        m_before = [a.copy() for a in state.mw + state.mb]
        v_before = [a.copy() for a in state.vw + state.vb]
        dto.adam_step(w, b, gw, gb, state, lr)
        eps_t = dto.EPSILON / math.sqrt(1.0 - dto.BETA2 ** step)
        for p0, g, m0, v0, p1 in zip(before, gw + gb, m_before, v_before, w + b):
            tp = torch.tensor(p0.copy(), requires_grad=True)
            opt = torch.optim.Adam([tp], lr=lr, betas=(dto.BETA1, dto.BETA2), eps=eps_t)
            tp.grad = torch.tensor(g)
            if step > 1:
                opt.state[tp] = {"step": torch.tensor(float(step - 1)), "exp_avg": torch.tensor(m0),
                                 "exp_avg_sq": torch.tensor(v0)}
            opt.step()
            np.testing.assert_allclose(p1, tp.detach().numpy(), rtol=1e-9, atol=1e-15)


def test_epoch_batches_follow_the_reference_rule():
    npr.seed(3)
    io, inw = dto.epoch_batches(8325, 900, 512, 0.9)
    assert io.shape == (8325 // 52, 52) and inw.shape == (8325 // 52, 460)
    assert len(np.unique(io)) == io.size and io.max() < 8325 and inw.max() < 900
    npr.seed(3)
    io2, inw2 = dto.epoch_batches(8325, 100, 512, 0.9)        # fewer new rows than 0.9 * batch: all of them per batch
    assert io2.shape[1] == 412 and inw2.shape[1] == 100


def test_training_reduces_the_loss():
    w, b, X, Z = _problem(4, 4, 3, 2, 40, 600)
    Zfit = np.tanh(X[:, :3]) * 0.5
    state = dto.AdamState(w, b)
    npr.seed(0)
    first = last = None
    for _ in range(6):
        io, inw = dto.epoch_batches(500, 100, 64, 0.5)
        losses = dto.train_batches(w, b, state, X[:500], Zfit[:500], X[500:], Zfit[500:], io, inw, 1e-2)
        first = losses[0] if first is None else first
        last = losses[-1]
    assert last < 0.5 * first
