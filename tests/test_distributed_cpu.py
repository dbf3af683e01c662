"""World-size-2 gloo test of the sharding / merge logic (no GPU).  The engine is replaced by a
numpy test double built on the oracle; what is under test is distributed.py: shard bounds,
global-index Philox sharding, the projection-sum all-reduce, the (score, k) merge and the
winner broadcast."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden_model, load_golden
from oracle import kde_oracle, mpc_oracle, philox
from smartstartcontinuous_b200.distributed import ShardedPlanner, ShardedSelector, argmax_pick, shard_bounds


def test_shard_bounds_cover_range():
    for total in (0, 1, 7, 8, 1000, 131072):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (o1, c1), (o2, _) in zip(spans, spans[1:]):
                assert o1 + c1 == o2
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_argmax_pick_semantics():
    assert argmax_pick([1.0, 3.0, 3.0], [10, 30, 20]) == 2          # tie -> lower global index
    assert argmax_pick([1.0, float("nan"), 5.0], [0, 7, 3]) == 1    # NaN wins like np.argmax
    assert argmax_pick([float("nan"), float("nan")], [9, 4]) == 1
    assert argmax_pick([2.0, 9.0], [5, -1]) == 0                    # empty shard ignored


class OracleEngine:
    """CPU test double with the Engine methods ShardedPlanner / ShardedSelector call."""

    def __init__(self, g):
        self.w, self.b, self.norm = golden_model(g)
        self.g = g
        self._model_shape = (self.w[-1].shape[1], self.w[0].shape[0] - self.w[-1].shape[1], len(self.w) - 1,
                             self.w[0].shape[1])

    def rollout(self, state, wp_index, *, actions, K, H, seed, act_low, act_high, gamma,
                horizontal_penalty_factor, penalty_mode, precision, k_offset, K_global, rng_state=None):
        self.sampler = None
        if rng_state is not None:
            # what Engine.rollout(rng_state=...) does on the GPU: this rank's slice of the ONE global
            # npr.uniform(low, high, (K_global, H, da)) draw, and the generator state after the whole draw
            rs = np.random.RandomState()
            rs.set_state(rng_state)
            full = rs.uniform(np.asarray(act_low, dtype=np.float64), np.asarray(act_high, dtype=np.float64),
                              (K_global, H, 1))
            actions = full[k_offset:k_offset + K]
            self._rng_after = rs.get_state()
        if actions is None:
            self.sampler = (H, seed, act_low, act_high)
            actions = philox.sample_actions(K, H, 1, seed, act_low, act_high, k_offset=k_offset)
        g = self.g
        self.args = (g["out_desired_states"], g["out_distances_left"], g["out_radii"], wp_index, gamma,
                     horizontal_penalty_factor)
        self.actions, self.k_offset, self.mode = actions, k_offset, penalty_mode
        self.states = mpc_oracle.forward_sim(state, actions, self.w, self.b, self.norm)
        lam = []
        mpc_oracle.score_add_delta(self.states, *self.args, penalty_mode=0, lambdas_out=lam)
        self.sums = torch.tensor(np.array(lam).reshape(-1), dtype=torch.float64)
        self.state = state

    def mt19937_state(self):
        return self._rng_after

    def projection_sums_tensor(self):
        return self.sums

    def finish(self):
        if self.mode == "reference":
            scores = _score_with_sums(self.states, self.sums.numpy().reshape(-1, 2), *self.args)
        else:
            scores = mpc_oracle.score_add_delta(self.states, *self.args, penalty_mode=1)
        k = int(np.argmax(scores))
        return k + self.k_offset, float(scores[k]), scores

    def finish_package_tensor(self, want_path):
        k, score, _ = self.finish()
        seq, path = self.replay(k)
        body = np.concatenate([seq.reshape(-1), path.reshape(-1)]) if want_path else np.zeros(seq.size + path.size)
        t = torch.tensor(np.concatenate([[score, float(k)], body]), dtype=torch.float64)
        return t, t.numel()

    def replay(self, k_global):
        if self.sampler is not None:      # device sampling: any rank can regenerate any sequence
            H, seed, lo, hi = self.sampler
            seq = philox.sample_actions(1, H, 1, seed, lo, hi, k_offset=k_global)
            return seq[0], mpc_oracle.forward_sim(self.state, seq, self.w, self.b, self.norm)[:, 0]
        k = k_global - self.k_offset
        return self.actions[k].copy(), self.states[:, k].copy()

    def select_start(self, all_states, queries, values, n, volume, alpha, beta):
        j, dens, ucb = kde_oracle.select_start(all_states, queries, values, n, volume, alpha, beta)
        return j, float(ucb[j]), None, None


def _score_with_sums(states, sums, DS, DL, r, wp, gamma, hpf):
    """generate_scores_add_delta with the projection coefficient taken from (all-reduced) sums."""
    T, K, _ = states.shape
    W = len(DS)
    dist_ = lambda a, b: np.sqrt((((a - b) / r) ** 2).sum(-1))
    idx = np.full(K, wp)
    prev = DL[idx] + dist_(states[0], DS[idx])
    scores = np.zeros(K)
    for t in range(T):
        x = states[t]
        dc, dn = dist_(DS[idx], x), dist_(DS[np.minimum(idx + 1, W - 1)], x)
        mv = np.logical_and(np.logical_or(dc <= 1.0, dn <= dc), idx != W - 1)
        idx = idx + mv
        dc = np.where(mv, dn, dc)
        to_end = DL[idx] + dc
        scores += (prev - to_end) * gamma ** t
        prev = to_end
        b0 = np.maximum(idx - 1, 0)
        a_, b_ = (x - DS[b0]) / r, (DS[b0 + 1] - DS[b0]) / r
        lam = sums[t, 0] / sums[t, 1]
        scores -= np.sqrt(((lam * b_ - a_) ** 2).sum(-1)) * hpf * gamma
    return scores


def _worker(rank, world, port, mode, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden("mpc_mountaincar_L2.npz")
        eng = OracleEngine(g)
        planner = ShardedPlanner(eng, device="cpu", tensors=eng)   # the oracle double is its own tensor strategy
        res = planner.plan(g["in_start_state"], 0, K=101, H=6, seed=11, act_low=[-1.0], act_high=[1.0],
                           penalty_mode=mode)
        res_host = planner.plan(g["in_start_state"], 0, K=len(g["in_actions"]), H=g["in_actions"].shape[1],
                                actions=g["in_actions"], penalty_mode=mode)
        # the default agent route: every rank takes its slice of one draw from numpy's legacy generator
        rs = np.random.RandomState(2024)
        rs.random_sample(77)
        res_mt = planner.plan(g["in_start_state"], 0, K=97, H=5, act_low=[-1.0], act_high=[1.0], penalty_mode=mode,
                              rng_state=rs.get_state())
        k = load_golden("kde_pendulum.npz")
        sel = ShardedSelector(eng, device="cpu").select_start(
            k["in_all_states"], k["in_queries"], k["in_values"], int(k["in_n_transitions"]),
            float(k["in_volume"]), 1.0, 2.0)
        out[rank] = (res["best_k"], res["best_score"], res["best_path"], res_host["best_k"],
                     res_host["best_sequence"], res_host["best_path"], sel,
                     (res_mt["best_k"], res_mt["best_sequence"], res_mt["rng_state"][1], res_mt["rng_state"][2]))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode", ["reference", "per_sample"])
def test_two_rank_plan_matches_single_process(mode):
    g = load_golden("mpc_mountaincar_L2.npz")
    w, b, norm = golden_model(g)
    args = (g["out_desired_states"], g["out_distances_left"], g["out_radii"], 0, .75, .5)
    acts = philox.sample_actions(101, 6, 1, 11, [-1.0], [1.0])
    single = mpc_oracle.plan(g["in_start_state"], acts, w, b, norm, *args, penalty_mode=0 if mode == "reference" else 1)
    single_host = mpc_oracle.plan(g["in_start_state"], g["in_actions"], w, b, norm, *args,
                                  penalty_mode=0 if mode == "reference" else 1)
    rs = np.random.RandomState(2024)
    rs.random_sample(77)
    acts_mt = rs.uniform(np.array([-1.0]), np.array([1.0]), (97, 5, 1))
    single_mt = mpc_oracle.plan(g["in_start_state"], acts_mt, w, b, norm, *args, penalty_mode=0 if mode == "reference" else 1)
    k = load_golden("kde_pendulum.npz")
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), mode, out), nprocs=2, join=True)
        for rank in (0, 1):
            best_k, best_score, path, hk, hseq, hpath, sel, mt = out[rank]
            assert mt[0] == single_mt["best_k"]
            np.testing.assert_array_equal(mt[1], acts_mt[single_mt["best_k"]])
            assert mt[3] == rs.get_state()[2] and np.array_equal(mt[2], rs.get_state()[1])
            assert best_k == single["best_k"]
            assert best_score == pytest.approx(single["best_score"], rel=1e-12)
            np.testing.assert_allclose(path, single["best_path"], rtol=1e-12)
            assert hk == single_host["best_k"]
            np.testing.assert_allclose(hseq, single_host["best_sequence"])
            np.testing.assert_allclose(hpath, single_host["best_path"], rtol=1e-12)
            assert sel[0] == int(k["out_best_j"])
            assert sel[1] == pytest.approx(float(k["out_ucb"][sel[0]]), rel=1e-10)
