"""Run the *unmodified reference code* on CPU (build container only).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

The reference (darren-huang/SmartStartContinuous) is pure Python but imports
TensorFlow 1.x, gym, matplotlib, baselines, mpi4py, google.cloud and seaborn at
module level, none of which exist in this image.  Everything on the hot path
except ``sess.run`` is numpy, so the harness

  1. copies ``/root/reference/smartstart`` into a temp dir (utilities.py:14-17
     creates a data dir at import time and the mount is read-only),
  2. serves empty stub modules for the missing third-party packages,
  3. restores ``np.product`` (removed in numpy 2; numerical.py:164 uses it),
  4. trims ``sys.argv`` (smartexplorationcontinuous.py:383-389 parses argv at import),
  5. builds agents with ``object.__new__`` + attributes (constructors need TF/gym),
  6. replaces the TF session by ``NumpySession`` which evaluates the float64 MLP of
     feedforward_network.py:12-23 (hidden: Linear+ReLU, output: Linear, y = xW + b).

It is used by oracle/make_golden.py (to produce tests/golden/*.npz) and by the CPU arm
of bench.py (`--impl reference` / cpu_baseline), which times the reference's own
get_best_sim_actions from the staged copy oracle/_ref (oracle/stage_ref.py).
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import shutil
import sys
import tempfile
import types
from collections import deque

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("SS_REFERENCE_ROOT", "/root/reference")
# on the GPU box: the unmodified python files of the reference packed by oracle/stage_ref.py
STAGED_ARCHIVE = os.path.join(_HERE, "_ref", "smartstart_ref.zip")

_STUB_ROOTS = ("tensorflow", "gym", "matplotlib", "mpl_toolkits", "google",
               "baselines", "mpi4py", "seaborn")


class _StubModule(types.ModuleType):
    """A module whose every attribute is another stub (callable, subclassable)."""

    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        if name == "float64":
            return np.float64
        if name == "float32":
            return np.float32
        cls = type(name, (object,), {"__init__": lambda self, *a, **k: None,
                                     "__call__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        # `from baselines.ddpg.noise import *` (DDPG_Baselines_agent.py:5) needs real names
        if module.__name__ == "baselines.ddpg.noise":
            for name in ("ActionNoise", "NormalActionNoise", "OrnsteinUhlenbeckActionNoise",
                         "AdaptiveParamNoiseSpec"):
                setattr(module, name, type(name, (object,), {"__init__": lambda self, *a, **k: None}))


_loaded = None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "smartstart")) or os.path.isfile(STAGED_ARCHIVE)


def load_reference():
    """Import the reference package from a writable temp copy; returns a namespace."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    tmp = tempfile.mkdtemp(prefix="ss_ref_")
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "smartstart")):
        shutil.copytree(os.path.join(REFERENCE_ROOT, "smartstart"), os.path.join(tmp, "smartstart"))
    else:
        import zipfile
        with zipfile.ZipFile(STAGED_ARCHIVE) as z:
            z.extractall(tmp)
    if not hasattr(np, "product"):
        np.product = np.prod  # numerical.py:164
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, os.path.join(tmp, "smartstart"))   # bare `utilities...` imports
    sys.path.insert(0, tmp)
    argv, sys.argv = sys.argv, [sys.argv[0]]
    try:
        import smartstart.utilities.numerical as numerical
        import smartstart.utilities.utilities as utilities
        import smartstart.RLAgents.replay_buffer as replay_buffer
        import smartstart.RLAgents.NND_MB_agent as nnd
        import smartstart.RLContinuousAlgorithms.NN_Dynamics_Model.dynamics_model as dyn
        import smartstart.smartexploration.smartexplorationcontinuous as ssc
    finally:
        sys.argv = argv
    _loaded = types.SimpleNamespace(numerical=numerical, utilities=utilities,
                                    replay_buffer=replay_buffer, nnd=nnd, dyn=dyn, ssc=ssc,
                                    tmpdir=tmp)
    return _loaded


class NumpySession:
    """Stands in for tf.Session: evaluates the dynamics MLP in float64 numpy.

    feedforward_network.py:12-23: ``num_fc_layers`` x (fully_connected, relu) then a
    linear output layer; tf.contrib fully_connected computes x @ W + b, W [in, out].
    """

    def __init__(self, weights, biases):
        self.weights = [np.asarray(w, dtype=np.float64) for w in weights]
        self.biases = [np.asarray(b, dtype=np.float64) for b in biases]

    def run(self, fetches, feed_dict=None):
        (x,) = feed_dict.values()
        h = np.asarray(x, dtype=np.float64)
        last = len(self.weights) - 1
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            h = h @ w + b
            if i != last:
                h = np.maximum(h, 0.0)
        return [h]


class _ArrayableDeque(deque):
    """numpy>=2 refuses the ragged ``np.array(buffer)`` at smartexplorationcontinuous.py:272."""

    def __array__(self, dtype=None, copy=None):
        out = np.empty((len(self), 5), dtype=object)
        for i, step in enumerate(self):
            for j in range(5):
                out[i, j] = step[j]
        return out


def make_dyn_model(ref, weights, biases, norm):
    """Reference Dyn_Model without TF: attributes as set in dynamics_model.py:14-33."""
    m = object.__new__(ref.dyn.Dyn_Model)
    m.sess = NumpySession(weights, biases)
    m.x_ = "x_placeholder"
    m.curr_nn_output = "nn_output"
    for k in ("mean_x", "std_x", "mean_y", "std_y", "mean_z", "std_z"):
        setattr(m, k, np.asarray(norm[k], dtype=np.float64))
    return m


class _Box:
    def __init__(self, low, high):
        self.low = np.asarray(low, dtype=np.float64)
        self.high = np.asarray(high, dtype=np.float64)
        self.shape = self.low.shape


class _Env:
    def __init__(self, low, high):
        self.action_space = _Box(low, high)


def make_nnd_agent(ref, weights, biases, norm, act_low, act_high, *, horizon, num_control_samples,
                   gamma=.75, horizontal_penalty_factor=.5, steps_per_waypoint=1,
                   mean_per_stepsize=1, std_per_stepsize=1, stepsizes_in_waypoint_radii=1,
                   path_shortcutting=True, final_steps=10, steps_before_giving_up_on_waypoint=5,
                   replay_buffer=None):
    """Reference NND_MB_agent with the attributes NND_MB_agent.py:142-206 would set."""
    a = object.__new__(ref.nnd.NND_MB_agent)
    a.env = _Env(act_low, act_high)
    a.N = num_control_samples
    a.horizon = horizon
    a.gamma = gamma
    a.horizontal_penalty_factor = horizontal_penalty_factor
    a.steps_per_waypoint = steps_per_waypoint
    a.mean_per_stepsize = mean_per_stepsize
    a.std_per_stepsize = std_per_stepsize
    a.stepsizes_in_waypoint_radii = stepsizes_in_waypoint_radii
    a.theta = 1
    a.path_shortcutting = path_shortcutting
    a.final_steps = final_steps
    a.steps_before_giving_up_on_waypoint = steps_before_giving_up_on_waypoint
    a.num_episodes_finished = 1            # != 0 mod aggregation -> no TF training call
    a.num_episodes_for_aggregation = 10 ** 9
    a.actions_done_for_current_waypoint = None
    a.radii = None
    a.distance_function = None
    a.stds = None
    a.path_to_follow = None
    a.desired_states = None
    a.current_desired_state_index = None
    a.distances_left = None
    a.noise_amount = 0.0
    a.replay_buffer = replay_buffer
    a.dyn_model = make_dyn_model(ref, weights, biases, norm)
    return a


class _ValueAgent:
    """Base agent exposing get_state_value like DDPG_Baselines_agent.py:197-204 ((m,1) float32)."""

    def __init__(self, fn):
        self.fn = fn

    def get_state_value(self, states):
        return np.asarray(self.fn(np.asarray(states)), dtype=np.float32).reshape(-1, 1)


def make_replay_buffer(ref, episodes, max_size):
    """Fill a reference ReplayBuffer through its own add/start_new_episode API.

    ``episodes``: list of (states [T+1,d], actions [T,da]) arrays.
    """
    main = object()
    rb = ref.replay_buffer.ReplayBuffer(main, max_size)
    rb.buffer = _ArrayableDeque()
    for states, actions in episodes:
        rb.start_new_episode(main)
        T = len(actions)
        for t in range(T):
            rb.add(main, np.array(states[t]), np.array(actions[t]), 0.0, t == T - 1,
                   np.array(states[t + 1]))
    return rb


def make_smartstart(ref, replay_buffer, nnd_agent, value_fn, *, n_ss, exploitation_param=1.,
                    exploration_param=2.):
    s = object.__new__(ref.ssc.SmartStartContinuous)
    s.replay_buffer = replay_buffer
    s.nnd_mb_agent = nnd_agent
    s.agent = _ValueAgent(value_fn)
    s.n_ss = n_ss
    s.exploitation_param = exploitation_param
    s.exploration_param = exploration_param
    s.print_ss_stuff = False
    return s
