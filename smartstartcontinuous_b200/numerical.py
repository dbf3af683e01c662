"""Host-side geometry for plan set-up (per SmartStart episode, cheap, stays on the CPU).

Same function names and argument meaning as the reference's
``smartstart/utilities/numerical.py`` so callers (and the reference's own unit
tests, restated in tests/test_numerical.py) work unchanged:

  path_deltas_stds_and_means_per_dim  numerical.py:30-59
  radii_calc                           numerical.py:61-62
  euclidean_distance                   numerical.py:64-72
  dist_line_seg_to_point               numerical.py:74-81
  projection_of_a_onto_b               numerical.py:84-98
  elliptical_euclidean_distance_function_generator   numerical.py:101-126
  volume_of_n_dimensional_hyperellipsoid             numerical.py:157-164
  binary_search_index_lower            numerical.py:166-187
  length_weighted_activities_solver    numerical.py:189-222
  path_shortcutter                     numerical.py:226-246

The per-(sample, step) versions of the distance / projection maths run inside the
CUDA kernels (csrc/score.cuh); these host versions are what the agents use for the
O(P) / O(P^2) per-episode preparation and for single-state checks in observe().
"""
from __future__ import annotations

import bisect
import math

import numpy as np


def moving_average(values, window=10):
    if window == 1:
        return values
    return np.convolve(values, np.full(window, 1.0 / window), "valid")


def path_deltas_stds_and_means_per_dim(path):
    """Per-dimension std and mean of |s[t+1] - s[t]| along a path -> (stds, means)."""
    if len(path) <= 1:
        # the reference returns a str here and its caller then fails to unpack it
        raise ValueError("path needs at least two states to have step sizes")
    p = np.asarray(path, dtype=np.float64)
    steps = np.abs(np.diff(p, axis=0))
    return steps.std(axis=0), steps.mean(axis=0)


def radii_calc(means, stds, num_means, num_stds, num_steps):
    return (num_means * means + num_stds * stds) * num_steps


def euclidean_distance(state, other_state):
    state, other_state = np.asarray(state), np.asarray(other_state)
    return np.sqrt(((state - other_state) ** 2).sum(axis=max(state.ndim, other_state.ndim) - 1))


def projection_of_a_onto_b(a, b, radii=None):
    """Projection of a onto b (optionally in the radii-scaled basis).

    NOTE (reference quirk, kept on purpose): the dot products are taken over *all*
    elements -- for [K, d] batches the coefficient is one scalar shared by the whole
    batch (numerical.py:89-93 call np.sum without an axis).  The MPC kernels implement
    exactly this in penalty_mode="reference".
    """
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if radii is not None:
        radii = np.asarray(radii, dtype=np.float64)
        a, b = a / radii, b / radii
    b_len = np.sqrt((b * b).sum())
    proj = ((a * b).sum() / b_len) * (b / b_len)
    if radii is not None:
        proj = proj * radii
    return proj


def dist_line_seg_to_point(line_seg_begin, line_seg_end, pt, dist_func, radii):
    rel_pt = np.asarray(pt) - np.asarray(line_seg_begin)
    rel_line = np.asarray(line_seg_end) - np.asarray(line_seg_begin)
    return dist_func(projection_of_a_onto_b(rel_pt, rel_line, radii=radii), rel_pt)


def elliptical_euclidean_distance_function_generator(radii):
    """Distance in which every point of the ellipse with semi-axes ``radii`` is at 1."""
    for r in radii:
        assert r > 0
    radii = np.asarray(radii, dtype=np.float64)

    def distance_func(state, other_state):
        state, other_state = np.asarray(state), np.asarray(other_state)
        axis = max(state.ndim, other_state.ndim) - 1
        return np.sqrt((((state - other_state) / radii) ** 2).sum(axis=axis))

    distance_func.radii = radii
    return distance_func


def volume_of_n_dimensional_hyperellipsoid(radii):
    d = len(radii)
    return (math.pi ** (d / 2.0)) / math.gamma(d / 2.0 + 1) * np.prod(np.asarray(radii, dtype=np.float64))


def binary_search_index_lower(sorted_array, target, key=lambda x: x):
    """Index of ``target`` if present, else of the largest smaller element (None if none)."""
    t = key(target)
    if key(sorted_array[0]) > t:
        return None
    lo, hi = 0, len(sorted_array) - 1
    while lo <= hi:
        mid = (lo + hi) // 2
        v = key(sorted_array[mid])
        if v == t:
            return mid
        if v < t:
            lo = mid + 1
        else:
            hi = mid - 1
    return hi


def length_weighted_activities_solver(activities, sub_extra=0):
    """Weighted interval scheduling with weight = end - start - sub_extra.

    Returns (optimal weight, chosen [start, end] intervals in order).  Tie-breaking and
    the reference's first-interval quirk (its weight ignores ``sub_extra``,
    numerical.py:202) are preserved so path_shortcutter picks identical shortcuts.
    """
    if len(activities) == 0:
        return 0, []
    acts = sorted(activities, key=lambda iv: iv[1])          # stable, by end time
    # column form of the DP table: ends strictly increase
    ends = [0, acts[0][1]]
    best = [0, acts[0][1] - acts[0][0]]
    taken = [None, acts[0]]
    back = [0, 0]
    for iv in acts[1:]:
        j = bisect.bisect_right(ends, iv[0]) - 1             # last column with end <= start
        with_iv = best[j] + (iv[1] - iv[0] - sub_extra)
        if iv[1] == ends[-1]:
            if with_iv >= best[-1]:                          # taking wins ties
                best[-1], taken[-1], back[-1] = with_iv, iv, j
        else:
            ends.append(iv[1])
            if with_iv >= best[-1]:
                best.append(with_iv); taken.append(iv); back.append(j)
            else:
                best.append(best[-1]); taken.append(None); back.append(len(ends) - 2)
    chosen = []
    col = len(ends) - 1
    while True:
        if taken[col] is not None:
            chosen.append(taken[col])
        if back[col] == col:
            break
        col = back[col]
    chosen.reverse()
    return best[-1], chosen


def path_shortcutter(path, distance_func, theta, engine=None):
    """Drop interior states between any two states (>= 2 apart) that are within theta.

    With ``engine`` (a CUDA Engine) and an elliptical ``distance_func`` the whole function runs on
    the device (ss_path_shortcut: same float64 pair decisions, same DP tie-breaking)."""
    p = np.asarray(path)
    if engine is not None and getattr(distance_func, "radii", None) is not None and len(p) <= 16384:
        return p[engine.path_shortcut(p, distance_func.radii, theta)]
    if engine is not None and getattr(distance_func, "radii", None) is not None:
        pairs = engine.path_close_pairs(p, distance_func.radii, theta)
    else:
        pair_dist = distance_func(p[:, None, :], p[None, :, :])
        pairs = np.argwhere(np.triu(pair_dist <= theta, k=2))
    _, chosen = length_weighted_activities_solver(pairs.tolist(), sub_extra=1)
    drop = [i for s, e in chosen for i in range(s + 1, e)]
    return np.delete(p, drop, axis=0)


def get_start_waypoints_final_states_steps(path, steps_per_waypoint):
    """utilities.py:63-75: every steps_per_waypoint-th state of path[:-1] plus the last state."""
    if isinstance(path, np.ndarray):
        path = path.tolist()
    out = list(path[:-1:steps_per_waypoint])
    out.append(path[-1])
    return out
