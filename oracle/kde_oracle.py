"""CPU oracle for SmartStart stage 1: Gaussian KDE + UCB + argmax (float64 numpy).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

The algorithm lives in a third-party dependency of the reference that is not
vendored under /root/reference: ``scipy.stats.gaussian_kde`` (pinned scipy==1.1.0 in
Pipfile.lock:802; this image has scipy 1.18.1 -- same estimator: Scott factor,
ddof=1 covariance, equal weights).  Call sites: smartexplorationcontinuous.py:260
(fit) and :275 (evaluate).  ``kde_density`` restates the published formula;
``scipy_density`` calls the library itself, and tests pin one against the other and
both against tests/golden/kde_*.npz (outputs of the reference's own
get_smart_start_path, produced by oracle/make_golden.py).
"""
from __future__ import annotations

import math

import numpy as np


def kde_fit(data):
    """scipy gaussian_kde.__init__/set_bandwidth('scott')/_compute_covariance.

    data: [N, d] (rows are points; the reference passes all_states.T, i.e. [d, N]).
    Returns (chol [d,d] lower Cholesky factor of cov*factor^2, log-free norm constant).
    """
    data = np.asarray(data, dtype=np.float64)
    N, d = data.shape
    factor = N ** (-1.0 / (d + 4))                         # scotts_factor
    cov = np.atleast_2d(np.cov(data.T, bias=False))        # ddof = 1
    chol = np.linalg.cholesky(cov * factor ** 2)           # raises LinAlgError when singular
    norm = (2 * math.pi) ** (d / 2.0) * np.prod(np.diag(chol)) * N
    return chol, norm


def kde_density(data, queries, chunk=256):
    """p(q_j) = sum_i exp(-|L^-1 (q_j - x_i)|^2 / 2) / (N (2 pi)^(d/2) det L)."""
    data = np.asarray(data, dtype=np.float64)
    queries = np.asarray(queries, dtype=np.float64)
    chol, norm = kde_fit(data)
    inv = np.linalg.inv(chol)
    wd = data @ inv.T
    wq = queries @ inv.T
    d2 = (wd * wd).sum(1)
    out = np.empty(len(queries))
    for s in range(0, len(queries), chunk):
        q = wq[s:s + chunk]
        e = (q * q).sum(1)[:, None] + d2[None, :] - 2.0 * (q @ wd.T)
        # the expansion loses ~1e-13 abs; clamp tiny negatives
        out[s:s + chunk] = np.exp(-0.5 * np.maximum(e, 0.0)).sum(1)
    return out / norm


def kde_density_direct(data, queries):
    """Same as kde_density but with explicit differences (slow, small cases only)."""
    data = np.asarray(data, dtype=np.float64)
    queries = np.asarray(queries, dtype=np.float64)
    chol, norm = kde_fit(data)
    inv = np.linalg.inv(chol)
    wd = data @ inv.T
    wq = queries @ inv.T
    out = np.empty(len(queries))
    for j, q in enumerate(wq):
        diff = wd - q
        out[j] = np.exp(-0.5 * (diff * diff).sum(1)).sum()
    return out / norm


def scipy_density(data, queries):
    """The reference's literal calls: gaussian_kde(all_states.T, 'scott')(queries.T)."""
    import scipy.stats

    kernel = scipy.stats.gaussian_kde(np.asarray(data).T, bw_method="scott")
    return kernel(np.asarray(queries).T)


def ucb_scores(densities, values, n_transitions, volume, alpha, beta):
    """smartexplorationcontinuous.py:275-279.

    probability = density * volume ; C_hat = n * probability ;
    ucb = alpha * V + sqrt(beta * ln(n) / C_hat)      (n = len(replay_buffer)).
    """
    with np.errstate(divide="ignore", invalid="ignore"):
        prob = np.asarray(densities, dtype=np.float64) * volume
        c_hat = n_transitions * prob
        return alpha * np.asarray(values, dtype=np.float64) + \
            np.sqrt((beta * np.log(n_transitions)) / c_hat)


def select_start(data, queries, values, n_transitions, volume, alpha, beta, density_fn=None):
    """Stage-1 numeric core: returns (best_j, densities, ucb); best_j = first argmax (np.argmax)."""
    dens = (density_fn or kde_density)(data, queries)
    ucb = ucb_scores(dens, values, n_transitions, volume, alpha, beta)
    return int(np.argmax(ucb)), dens, ucb


def hyperellipsoid_volume(radii):
    """numerical.py:157-164: pi^(d/2) / Gamma(d/2 + 1) * prod(radii)."""
    d = len(radii)
    return (math.pi ** (d / 2.0)) / math.gamma(d / 2.0 + 1) * float(np.prod(radii))
