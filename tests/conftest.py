import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_addoption(parser):
    parser.addoption("--require-gpu", action="store_true", default=False,
                     help="fail (instead of skipping) the gpu-marked tests when no CUDA device is visible")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _cuda_device_visible():
    """True when the machine shows a CUDA device at all.  The library itself never falls back to the
    CPU: with a device present every failure of ss_create / a missing libss_b200.so stays a hard
    error, so a GPU box can never pass the gpu tests vacuously."""
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return os.path.exists("/dev/nvidia0")


def pytest_collection_modifyitems(config, items):
    if config.getoption("--require-gpu") or _cuda_device_visible():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible: gpu-marked tests need a B200 "
                                   "(pass --require-gpu to turn this skip into an error)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def golden_model(g):
    L = int(g["in_num_layers"])
    if "in_weight_seed" in g:        # big networks: weights regenerated from the seed (oracle/make_golden.py)
        from smartstartcontinuous_b200 import synthetic as syn
        weights, biases = syn.xavier_mlp(np.random.default_rng(int(g["in_weight_seed"])), int(g["in_d"]), 1, L,
                                         int(g["in_hidden"]), scale=float(g["in_weight_scale"]))
    else:
        weights = [g["in_w%d" % i] for i in range(L + 1)]
        biases = [g["in_b%d" % i] for i in range(L + 1)]
    norm = {k: g["in_" + k] for k in ("mean_x", "std_x", "mean_y", "std_y", "mean_z", "std_z")}
    return weights, biases, norm


@pytest.fixture(scope="session")
def engine():
    """One libss_b200 context on cuda:0 (GPU tests only).  No fallback: a missing library or
    GPU is an error, not a skip."""
    from smartstartcontinuous_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()
