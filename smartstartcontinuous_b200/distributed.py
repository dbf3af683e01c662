"""One process per GPU: shard the K action sequences / the KDE queries over the ranks.

The reference has no multi-device path; the north star adds exactly one
strategy: data-parallel sharding with replicated weights / plan / buffer and tiny merges:

  MPC   rank g rolls out sequences [g*K/G, (g+1)*K/G) (device Philox is indexed by the GLOBAL
        sequence number, so the samples do not depend on G); reference-exact penalty needs the
        per-time-step projection sums of ALL sequences (numerical.py:89-93) -> one all-reduce of
        2*(H+1) float64 (device memory, on the compute stream); then every rank packs its local
        winner (score, k, best_sequence, best_path) on the device, ONE all-gather of those
        packages (2 + H*da + (H+1)*d float64 per rank) and an np.argmax-ordered pick on every
        rank: two small collectives and a single device->host copy per decision.
  KDE   rank g scores queries [g*m/G, (g+1)*m/G) against the full (replicated) buffer; one
        all-gather of (ucb, j).

Collectives go through torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests).
There is no data-path collective beyond those few bytes.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(total, world, rank):
    """Contiguous, balanced split of range(total): (offset, count) of `rank`."""
    base, rem = divmod(int(total), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def argmax_pick(values, indices):
    """np.argmax ordering over (value, global index) pairs: NaN beats everything, then the
    larger value, ties -> the lower global index (first occurrence).  Entries with index < 0
    are ignored."""
    best = -1
    for i, (v, k) in enumerate(zip(values, indices)):
        if k < 0:
            continue
        if best < 0:
            best = i
            continue
        bv, bk = values[best], indices[best]
        v_nan, b_nan = v != v, bv != bv
        if v_nan or b_nan:
            better = (v_nan and b_nan and k < bk) or (v_nan and not b_nan)
        else:
            better = v > bv or (v == bv and k < bk)
        if better:
            best = i
    return best


class _DevView:
    """Zero-copy torch view of library-owned device memory (via __cuda_array_interface__)."""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _dist():
    import torch.distributed as dist
    return dist


class EngineTensors:
    """How ShardedPlanner turns library-owned device memory into tensors for the collectives:
    zero-copy torch views of the Engine's projection sums and winner package.  (The CPU tests inject
    an object with the same two methods backed by the oracle.)"""

    def __init__(self, engine, device):
        self.engine = engine
        self.device = device

    def projection_sums_tensor(self):
        import torch
        ptr, n = self.engine.projection_sums_ptr()
        if n == 0:
            return None
        return torch.as_tensor(_DevView(ptr, n), device=self.device)

    def finish_package_tensor(self, want_path):
        """(tensor view of this rank's package, element count)."""
        import torch
        ptr, n = self.engine.finish_package(want_path)
        return torch.as_tensor(_DevView(ptr, n), device=self.device), n


class ShardedPlanner:
    """MPC over all ranks of a torch.distributed group; every rank returns the same result.

    The Engine queues its kernels on ONE stream and returns from rollout / finish_package without a
    host synchronisation; the collectives of the NCCL route read and write library memory in place.
    Both therefore have to run on the same stream: the constructor binds the Engine to torch's
    current stream of `device` (Engine.set_stream), and plan() asserts the binding still holds."""

    def __init__(self, engine, device=None, group=None, tensors=None):
        dist = _dist()
        self.engine = engine
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device          # torch device for the merge tensors ("cuda:N" or "cpu")
        self.tensors = tensors if tensors is not None else EngineTensors(engine, device)
        self._stream = None
        if tensors is None and self._on_cuda():
            import torch
            self._stream = torch.cuda.current_stream(device).cuda_stream
            engine.set_stream(self._stream)

    def _on_cuda(self):
        return self.device is not None and str(self.device).startswith("cuda")

    def plan(self, state, wp_index, *, K, H, seed=0, act_low=None, act_high=None, actions=None,
             gamma=.75, horizontal_penalty_factor=.5, penalty_mode="reference", precision="auto",
             want_path=True, local_actions=None, rng_state=None):
        """actions: the GLOBAL [K, H, da] host samples (every rank passes the same array), or
        local_actions: this rank's own [K/world, H, da] slice, or rng_state: numpy's legacy generator
        state (identical on every rank) -- each rank then generates its slice of the reference's
        npr.uniform(low, high, (K, H, da)) draw on its GPU (MT19937 jump-ahead) and the result carries
        the advanced state -- or none of them (device Philox)."""
        import torch
        dist = _dist()
        if K < self.world:
            # every rank sees the same K: all of them raise before any exchange kernel is queued
            raise ValueError("K = %d sequences cannot be sharded over %d ranks (empty shard)" % (K, self.world))
        if self._stream is not None and self.world > 1:
            cur = torch.cuda.current_stream(self.device).cuda_stream
            if cur != self._stream:
                raise RuntimeError("ShardedPlanner: torch's current stream changed since construction; the Engine "
                                   "and the collectives must share one stream (call Engine.set_stream)")
        k_offset, k_local = shard_bounds(K, self.world, self.rank)
        if local_actions is None and actions is not None:
            local_actions = actions[k_offset:k_offset + k_local]
        if self.world == 1 and isinstance(self.tensors, EngineTensors):
            # nothing to merge: the single-call form (one trip through the C ABI, ss_mpc_plan)
            res = self.engine.plan(state, wp_index, actions=local_actions, K=K, H=H, seed=seed, act_low=act_low,
                                   act_high=act_high, gamma=gamma,
                                   horizontal_penalty_factor=horizontal_penalty_factor, penalty_mode=penalty_mode,
                                   precision=precision, want_path=want_path, rng_state=rng_state)
            res.update(owner=0, k_offset=0, k_local=k_local)
            return res
        self.engine.rollout(state, wp_index, actions=local_actions, K=k_local, H=H, seed=seed,
                            act_low=act_low, act_high=act_high, gamma=gamma,
                            horizontal_penalty_factor=horizontal_penalty_factor,
                            penalty_mode=penalty_mode, precision=precision, k_offset=k_offset,
                            K_global=K, **({} if rng_state is None else dict(rng_state=rng_state)))
        # with an open peer exchange (Engine.peer_setup) both merges already happened inside the
        # kernels over NVLink peer memory; otherwise they are two small collectives
        peer = self.world > 1 and bool(getattr(self.engine, "peer_ready", False))
        if self.world > 1 and not peer and penalty_mode in ("reference", 0):
            sums = self.tensors.projection_sums_tensor()
            if sums is not None:
                dist.all_reduce(sums, group=self.group)      # 2*(H+1) float64
        d, da = self.engine._model_shape[0], self.engine._model_shape[1]
        if isinstance(self.tensors, EngineTensors) and (peer or self.world == 1):
            _, n = self.engine.finish_package(want_path)
            pk = self.engine.read_package(n).reshape(1, -1)            # peer route: already the global winner
        else:
            mine, n = self.tensors.finish_package_tensor(want_path)
            if self.world > 1:
                gathered = torch.empty(self.world * mine.numel(), dtype=torch.float64, device=mine.device)
                dist.all_gather_into_tensor(gathered, mine, group=self.group)
                pk = gathered.cpu().numpy().reshape(self.world, -1)    # the one host sync
            else:
                pk = mine.cpu().numpy().reshape(1, -1)
        row = argmax_pick(pk[:, 0].tolist(), [int(v) for v in pk[:, 1]])
        best_score, best_k = float(pk[row, 0]), int(pk[row, 1])
        w = row
        if peer:                                                       # owner = the shard that holds best_k
            w = next((r for r in range(self.world)
                      if shard_bounds(K, self.world, r)[0] <= best_k < sum(shard_bounds(K, self.world, r))), 0)
        seq = path = None
        if want_path:
            seq = pk[row, 2:2 + H * da].reshape(H, da).copy()
            path = pk[row, 2 + H * da:2 + H * da + (H + 1) * d].reshape(H + 1, d).copy()
        res = dict(best_k=best_k, best_score=best_score, best_sequence=seq, best_path=path,
                   owner=w, k_offset=k_offset, k_local=k_local)
        if isinstance(rng_state, int):
            self.engine.mt19937_state_into(rng_state)   # numpy's generator advanced in place
        elif rng_state is not None:
            res["rng_state"] = self.engine.mt19937_state()
        return res


class ShardedSelector:
    """KDE + UCB + argmax with the queries sharded over the ranks (buffer replicated)."""

    def __init__(self, engine, device=None, group=None):
        dist = _dist()
        self.engine = engine
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device

    def _merge(self, j, ucb):
        import torch
        dist = _dist()
        if self.world == 1:
            return j, ucb
        if bool(getattr(self.engine, "peer_ready", False)):
            # one small kernel over NVLink peer memory, result through mapped host memory: no NCCL
            # call and no stream synchronisation
            v, k = self.engine.peer_argmax_merge(ucb, j)
            return k, v
        mine = torch.tensor([ucb, float(j)], dtype=torch.float64, device=self.device)
        gathered = torch.empty(2 * self.world, dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(gathered, mine, group=self.group)
        pairs = gathered.cpu().numpy().reshape(self.world, 2)
        w = argmax_pick(pairs[:, 0].tolist(), [int(v) for v in pairs[:, 1]])
        return int(pairs[w, 1]), float(pairs[w, 0])

    def select_start_dev(self, data_ptr, n_pts, d, queries_ptr, m, values_ptr, n_transitions, volume=1.0,
                         alpha=1.0, beta=2.0):
        """Same with everything already resident on this rank's GPU (device pointers to the full
        [n_pts, d] float64 buffer, [m, d] float64 queries, [m] float32 values): each rank scores its
        contiguous slice of the queries."""
        off, cnt = shard_bounds(m, self.world, self.rank)
        if cnt > 0:
            j, ucb = self.engine.select_start_dev(data_ptr, n_pts, d, queries_ptr + off * d * 8, cnt,
                                                  values_ptr + off * 4, n_transitions, volume, alpha, beta)
            j += off
        else:
            j, ucb = -1, 0.0
        return self._merge(j, ucb)

    def select_start(self, all_states, queries, values, n_transitions, volume=1.0, alpha=1.0, beta=2.0):
        import torch
        dist = _dist()
        off, cnt = shard_bounds(len(queries), self.world, self.rank)
        if cnt > 0:
            j, ucb, _, _ = self.engine.select_start(all_states, queries[off:off + cnt],
                                                    np.asarray(values).reshape(-1)[off:off + cnt],
                                                    n_transitions, volume, alpha, beta)
            j += off
        else:
            j, ucb = -1, 0.0
        return self._merge(j, ucb)
