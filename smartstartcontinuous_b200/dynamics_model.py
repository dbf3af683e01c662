"""Learned dynamics model s' - s = f(s, a): API of the reference's ``Dyn_Model``
(smartstart/RLContinuousAlgorithms/NN_Dynamics_Model/dynamics_model.py).

The reference builds a float64 TF1 graph (feedforward_network.py:3-23) and trains it with
Adam / MSE on a mix of the initial random-policy data and the replay-buffer data
(dynamics_model.py:52-171).  TensorFlow is not part of this framework: the parameters
live in torch float64 tensors (on the GPU when there is one), training is a plain torch
Adam loop with the reference's batching rule, and *inference* -- the MPC hot path -- is
not done here at all: the weights are pushed into the CUDA engine with ``export()`` /
``Engine.set_model`` and rolled out by the hand-written kernels.
``do_forward_sim`` keeps the reference's signature and runs on the engine.
"""
from __future__ import annotations

import math
import time

import numpy as np
import numpy.random as npr


def _xavier_normal(rng, shape, fan_in, fan_out):
    """tf.contrib.layers.xavier_initializer(uniform=False): truncated normal,
    stddev sqrt(1.3 * 2 / (fan_in + fan_out)) (feedforward_network.py:8)."""
    std = math.sqrt(1.3 * 2.0 / (fan_in + fan_out))
    x = rng.normal(0.0, std, size=shape)
    bad = np.abs(x) > 2 * std
    while bad.any():
        x[bad] = rng.normal(0.0, std, size=int(bad.sum()))
        bad = np.abs(x) > 2 * std
    return x


class Dyn_Model:
    def __init__(self, inputSize, outputSize, sess, learning_rate, batchsize, num_fc_layers,
                 depth_fc_layers, mean_x, mean_y, mean_z, std_x, std_y, std_z, tf_datatype, verbose,
                 engine=None, seed=None):
        import torch

        self.sess = sess                      # accepted for signature compatibility, unused
        self.batchsize = batchsize
        self.inputSize = inputSize
        self.outputSize = outputSize
        self.mean_x, self.mean_y, self.mean_z = mean_x, mean_y, mean_z
        self.std_x, self.std_y, self.std_z = std_x, std_y, std_z
        self.verbose = verbose
        self.engine = engine
        self.num_fc_layers = num_fc_layers
        self.depth_fc_layers = depth_fc_layers
        self._torch = torch
        self._dev = torch.device("cuda", engine.device) if (engine is not None and torch.cuda.is_available()) \
            else torch.device("cpu")
        rng = np.random.default_rng(seed)
        sizes = [inputSize] + [depth_fc_layers] * num_fc_layers + [outputSize]
        self.weights, self.biases = [], []
        for fi, fo in zip(sizes[:-1], sizes[1:]):
            # weights AND biases use the Xavier initialiser in the reference (:14-23)
            w = _xavier_normal(rng, (fi, fo), fi, fo)
            b = _xavier_normal(rng, (fo,), fo, fo)
            self.weights.append(torch.tensor(w, dtype=torch.float64, device=self._dev, requires_grad=True))
            self.biases.append(torch.tensor(b, dtype=torch.float64, device=self._dev, requires_grad=True))
        self.opt = torch.optim.Adam(self.weights + self.biases, lr=learning_rate)

    # ------------------------------------------------------------------ parameters
    def norm(self):
        return dict(mean_x=self.mean_x, std_x=self.std_x, mean_y=self.mean_y, std_y=self.std_y,
                    mean_z=self.mean_z, std_z=self.std_z)

    def export(self):
        """(weights, biases) as float64 numpy, [in, out] / [out]."""
        return ([w.detach().cpu().numpy().copy() for w in self.weights],
                [b.detach().cpu().numpy().copy() for b in self.biases])

    def set_weights(self, weights, biases):
        torch = self._torch
        with torch.no_grad():
            for dst, src in zip(self.weights + self.biases, list(weights) + list(biases)):
                dst.copy_(torch.as_tensor(np.asarray(src, dtype=np.float64)).reshape(dst.shape))
        self.push_to_engine()

    def push_to_engine(self):
        if self.engine is not None:
            w, b = self.export()
            self.engine.set_model(w, b, self.norm())

    # ------------------------------------------------------------------ training (off the hot path)
    def _forward(self, x):
        torch = self._torch
        h = x
        last = len(self.weights) - 1
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            h = h @ w + b
            if i != last:
                h = torch.relu(h)
        return h

    def _mse(self, x, z, train):
        torch = self._torch
        xt = torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=self._dev)
        zt = torch.as_tensor(np.ascontiguousarray(z), dtype=torch.float64, device=self._dev)
        if train:
            loss = ((zt - self._forward(xt)) ** 2).mean()
            self.opt.zero_grad()
            loss.backward()
            self.opt.step()
        else:
            with torch.no_grad():
                loss = ((zt - self._forward(xt)) ** 2).mean()
        return float(loss)

    def train(self, dataX, dataZ, dataX_new, dataZ_new, nEpoch, save_dir, fraction_use_new,
              save_results=True):
        """Batching rule of dynamics_model.py:52-171: each batch = batchsize*(1-fraction) rows of
        the shuffled old data + batchsize*fraction rows drawn from the new data."""
        start = time.time()
        losses = []
        n_old, n_new = dataX.shape[0], dataX_new.shape[0]
        new_per_batch = n_new if n_new < self.batchsize * fraction_use_new else int(self.batchsize * fraction_use_new)
        old_per_batch = int(self.batchsize - new_per_batch)
        avg, nb = 0.0, 0
        for epoch in range(nEpoch):
            avg, nb = 0.0, 0
            if old_per_batch > 0:
                order = npr.permutation(n_old)
                for bi in range(n_old // old_per_batch):
                    sel = order[bi * old_per_batch:(bi + 1) * old_per_batch]
                    xb, zb = dataX[sel], dataZ[sel]
                    if n_new:
                        pick = npr.randint(0, n_new, (new_per_batch,))
                        xb = np.concatenate((xb, dataX_new[pick]))
                        zb = np.concatenate((zb, dataZ_new[pick]))
                    loss = self._mse(xb, zb, True)
                    losses.append(loss); avg += loss; nb += 1
            else:
                for bi in range(n_new // new_per_batch):
                    sl = slice(bi * new_per_batch, (bi + 1) * new_per_batch)
                    loss = self._mse(dataX_new[sl], dataZ_new[sl], True)
                    losses.append(loss); avg += loss; nb += 1
                p = npr.permutation(n_new)
                dataX_new, dataZ_new = dataX_new[p], dataZ_new[p]
            if save_results and save_dir:
                np.save(save_dir + '/training_losses.npy', losses)
            if self.verbose and epoch % 10 == 0:
                print("\n=== Epoch {} ===".format(epoch))
                print("loss: ", avg / max(nb, 1))
        if self.verbose:
            print("Training set size: ", n_old + n_new)
            print("Training duration: {:0.2f} s".format(time.time() - start))
        old_loss = self.run_validation(dataX, dataZ, quiet=True) if n_old >= self.batchsize else 0
        new_loss = self.run_validation(dataX_new, dataZ_new, quiet=True) if n_new >= self.batchsize else 0
        self.push_to_engine()
        return avg / max(nb, 1), old_loss, new_loss

    def run_validation(self, inputs, outputs, quiet=False):
        n = inputs.shape[0]
        tot, it = 0.0, 0
        for bi in range(n // self.batchsize):
            sl = slice(bi * self.batchsize, (bi + 1) * self.batchsize)
            tot += self._mse(inputs[sl], outputs[sl], False)
            it += 1
        if self.verbose and not quiet:
            print("Validation set size: ", n)
            print("Validation set's total loss: ", tot / max(it, 1))
        return tot / max(it, 1)

    # ------------------------------------------------------------------ inference (engine)
    def do_forward_sim(self, forwardsim_x_true, forwardsim_y, many_in_parallel):
        """Multi-step open-loop prediction (dynamics_model.py:199-270) on the GPU engine.
        Returns a list of H+1 arrays [N, d] (parallel) or [d] (single sequence)."""
        if self.engine is None:
            raise RuntimeError("Dyn_Model.do_forward_sim needs a CUDA engine (no CPU fallback)")
        y = np.asarray(forwardsim_y, dtype=np.float64)
        if many_in_parallel:
            if len(forwardsim_x_true) != 2:
                raise ValueError("per-sequence start states are not supported by the engine rollout")
            states = self.engine.forward_sim(forwardsim_x_true[0], y)
            return [states[t] for t in range(states.shape[0])]
        states = self.engine.forward_sim(forwardsim_x_true[0], y[None, :, :])
        return [states[t, 0] for t in range(states.shape[0])]
