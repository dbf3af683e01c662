#!/usr/bin/env python
"""bench.py -- headline benchmark of the SmartStart hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  Headline metric: NND-MPC rollout-steps/s; the second half of
BASELINE.json's metric, KDE kernel-evals/s, is reported in the same line under "kde".

  step      one MPC decision (NND_MB_agent.get_best_sim_actions): K action sequences rolled H steps
            through the 2x500 dynamics MLP, scored against the waypoint plan, arg-best
            (reference-exact penalty mode).  Workload per GPU: Pendulum d=3, da=1, K=131072
            (= BASELINE config 4's K=1M / 8), H=50 -- at N=8 this is exactly config 4.
  value     K_total * H / device time, inputs resident (actions sampled on the device, Philox)
  e2e       same decision through the C ABI with HOST buffers: action samples [K,H,da] float64 in
            pinned host memory -> H2D inside the call -> rollout -> score -> D2H of the result
  kde       BASELINE config 2: 100 001 Pendulum states x 16 384 candidate queries per GPU
  roofline  tensor pipe for the rollout kernel (achieved = 507 000 FLOP x K*H / kernel time, against
            the measured sustained bf16 peak; frac_of_burst_peak = against the burst figure), SFU pipe
            (16 MUFU.EX2 / clk / SM) for the KDE pair kernel
  --impl reference: the oracle port of the reference's CPU path (numpy float64, BLAS on all host
            cores; scipy gaussian_kde with the queries chunked over a multiprocessing.Pool for the
            KDE), timed on bounded samples of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_STEP = {2: 505_000, 3: 507_000}      # 2*[(d+da)h + h^2 + h d], h=500
K_PER_GPU = 131_072
HORIZON = 50
KDE_N = 100_000
KDE_M = 16_384
SFU_PER_CLK_PER_SM = 16                        # MUFU.EX2 lanes per SM per clock (sm_100)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
# captures of exactly these workloads (profiles/r01e_ncu_summary.md); not measurable live
NCU_TRAFFIC = {"mpc_rollout_tc_kernel": 675_072 + 52_459_008, "kde_pairs_tc_kernel": 7_491_072}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_sustained=p["bf16_tflops_sustained"], bf16_burst=p["bf16_tflops"],
                    sm_max_mhz=p.get("sm_max_mhz", 1965.0), source="MEASURED_PEAKS.json")
    return dict(hbm_gbs=6650.0, bf16_sustained=1400.0, bf16_burst=1590.0, sm_max_mhz=1965.0,
                source="fallback (B200_PROFILING.md)")


def make_workload(seed=0):
    """Seeded synthetic inputs of the benchmark shapes."""
    from smartstartcontinuous_b200 import synthetic as syn
    from smartstartcontinuous_b200.nnd_mb_agent import plan_from_path
    rng = np.random.default_rng(seed)
    obs, act = syn.pendulum_rollouts(rng, 25, 333)          # 25 x 333 random-policy transitions
    norm = syn.normalisation_stats(np.concatenate(list(obs)),
                                   np.concatenate([np.concatenate([a, a[-1:]]) for a in act]))
    w, b = syn.xavier_mlp(rng, 3, 1, 2, 500, scale=0.5)
    plan = plan_from_path(list(obs[0][:80]), mean_per_stepsize=1, std_per_stepsize=1,
                          stepsizes_in_waypoint_radii=1, path_shortcutting=True, theta=1, steps_per_waypoint=1)
    return dict(w=w, b=b, norm=norm, plan=plan, state=obs[0][0].copy(), low=[-2.0], high=[2.0])


def kde_workload(seed=0, n=KDE_N, m=KDE_M):
    from smartstartcontinuous_b200 import synthetic as syn
    all_states, s2, _ = syn.pendulum_buffer(n, seed=seed)
    rng = np.random.default_rng(seed)
    q = np.ascontiguousarray(s2[rng.choice(n, m, replace=False)])
    return dict(all_states=all_states, queries=q, values=syn.critic_like_values(q), n=n, volume=1e-3)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except OSError:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        busy = [v for v in sm if v > 0]
        return dict(sm_mhz=float(np.median(busy)) if busy else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def workload_name():
    return ("NND_MB random-shooting MPC, Pendulum-v0 d=3 da=1, K=%d per GPU (BASELINE config 4 = K=1M over 8 GPUs), "
            "H=%d, MLP 2x500, reference-exact penalty" % (K_PER_GPU, HORIZON))


def cpu_reference_mpc(wl, K, H, seed):
    """The reference's CPU planner (oracle port): npr.uniform sampling + float64 numpy rollout +
    scoring.  Returns seconds."""
    from oracle import mpc_oracle
    rs = np.random.RandomState(seed)
    t0 = time.perf_counter()
    acts = rs.uniform(wl["low"], wl["high"], (K, H, 1))
    mpc_oracle.plan(wl["state"], acts, wl["w"], wl["b"], wl["norm"], wl["plan"]["desired_states"],
                    wl["plan"]["distances_left"], wl["plan"]["radii"], 0, .75, .5)
    return time.perf_counter() - t0


def cpu_reference_kde(kw, m):
    """scipy.stats.gaussian_kde fit + evaluate + UCB + argmax on one core (the reference as written,
    smartexplorationcontinuous.py:260-280).  Returns seconds."""
    from oracle import kde_oracle
    t0 = time.perf_counter()
    kde_oracle.select_start(kw["all_states"], kw["queries"][:m], kw["values"][:m], kw["n"], kw["volume"], 1.0, 2.0,
                            density_fn=kde_oracle.scipy_density)
    return time.perf_counter() - t0


_POOL_KDE = {}


def _pool_kde_chunk(span):
    lo, hi = span
    return _POOL_KDE["kernel"](_POOL_KDE["queries"][lo:hi].T)


def cpu_reference_kde_pool(kw, m, procs, reps):
    """The same numeric core with the queries chunked over a multiprocessing.Pool (the reference's
    own parallel idiom, smartstart/utilities/experimenter.py:84-92): fit once in the parent
    (timed), evaluate in `procs` forked workers, UCB + argmax in the parent.  Returns the mean
    seconds of `reps` repetitions (pool start-up excluded)."""
    import multiprocessing as mp
    from scipy.stats import gaussian_kde
    q = np.ascontiguousarray(kw["queries"][:m])
    _POOL_KDE["queries"] = q
    _POOL_KDE["kernel"] = gaussian_kde(kw["all_states"].T, bw_method="scott")
    spans = [(int(a[0]), int(a[-1]) + 1) for a in np.array_split(np.arange(m), procs) if len(a)]
    times = []
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_pool_kde_chunk, [(0, 8)] * procs)                      # warm the workers
        for _ in range(reps):
            t0 = time.perf_counter()
            kernel = gaussian_kde(kw["all_states"].T, bw_method="scott")   # the fit is part of the path
            _ = kernel.factor
            dens = np.concatenate(pool.map(_pool_kde_chunk, spans))
            c_hat = kw["n"] * dens * kw["volume"]
            ucb = 1.0 * np.asarray(kw["values"][:m], dtype=np.float64) + np.sqrt(2.0 * np.log(kw["n"]) / c_hat)
            int(np.argmax(ucb))
            times.append(time.perf_counter() - t0)
    return float(np.mean(times))


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


REF_K_SAMPLE = 4096          # sequences per timed MPC step of the CPU arm (~0.5 s on a server CPU)


def run_reference(args, guard):
    """--impl reference: the reference's CPU path (oracle port; scipy for the KDE) on rank 0 only,
    all host threads, bounded samples of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload()
    kw = kde_workload()
    K_s = REF_K_SAMPLE
    for _ in range(args.warmup):
        cpu_reference_mpc(wl, 512, HORIZON, 0)
    t_mpc = [cpu_reference_mpc(wl, K_s, HORIZON, i) for i in range(args.steps)]
    v = K_s * HORIZON / float(np.mean(t_mpc))
    cores = blas_threads()
    procs = os.cpu_count() or 1
    m_pool = min(KDE_M, 256 * procs)
    reps = max(1, min(args.steps, 5))
    t_pool = cpu_reference_kde_pool(kw, m_pool, procs, reps)
    kv = m_pool * (kw["n"] + 1) / t_pool
    t_one = [cpu_reference_kde(kw, 256) for _ in range(2)]
    kv_one = 256 * (kw["n"] + 1) / float(np.mean(t_one))
    sample = "%d steps of K=%d (of %d) sequences, H=%d (rate is K-independent: GEMM-bound)" % (
        args.steps, K_s, K_PER_GPU, HORIZON)
    guard.emit(json.dumps({
        "impl": "reference", "metric": "mpc_rollout_steps_per_s", "value": v, "unit": "rollout-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(t_mpc)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(),
                   "note": "oracle port of the reference's numpy/TF-CPU path; TensorFlow 1.5 is not installable, "
                           "its float64 GEMMs run in numpy/BLAS"},
        "cpu_baseline": {"value": v, "unit": "rollout-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "kde": {"metric": "kde_kernel_evals_per_s", "value": kv, "unit": "kernel-evals/s", "cores": procs,
                "kind": "reference-library (scipy.stats.gaussian_kde; queries chunked over a "
                        "multiprocessing.Pool, the reference's parallel idiom)",
                "sample": "%d x %d (of %d) queries x %d points" % (reps, m_pool, KDE_M, kw["n"] + 1),
                "single_core": {"value": kv_one, "cores": 1, "sample": "2 x 256 queries x %d points" % (kw["n"] + 1)}},
    }))


class _StdoutGuard:
    """Everything libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; only the
    final JSON line reaches the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())


def main():
    guard = _StdoutGuard()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args, guard)
        return

    import torch
    import torch.distributed as dist
    from smartstartcontinuous_b200.distributed import ShardedPlanner, ShardedSelector
    from smartstartcontinuous_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    eng = Engine(local_rank)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    wl = make_workload()
    eng.set_model(wl["w"], wl["b"], wl["norm"])
    eng.set_plan(wl["plan"]["desired_states"], wl["plan"]["distances_left"], wl["plan"]["radii"])
    if world > 1 and os.environ.get("SS_PEER", "1") != "0":
        eng.peer_setup()         # projection sums + winner packages over NVLink peer memory, inside the kernels
    planner = ShardedPlanner(eng, device=dev)
    selector = ShardedSelector(eng, device=dev)
    K_total = K_PER_GPU * world
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """Per-step CUDA-event times on the launching stream; L2 flushed between steps (outside
        the events).  Returns (sum of per-step ms, max over ranks)."""
        for _ in range(warmup):
            fn(0)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            flush.zero_()
            ev[i][0].record(stream)
            fn(i + 1)
            ev[i][1].record(stream)
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    precision = "bf16_tc" if eng.tc_supported() else "fp32"
    rollout_ms = []

    def mpc_step(i):
        planner.plan(wl["state"], 0, K=K_total, H=HORIZON, seed=1000 + i, act_low=wl["low"], act_high=wl["high"],
                     penalty_mode="reference", precision=precision, want_path=True)
        if i > 0:
            rollout_ms.append(dict(eng.last_timings()).get("mpc_rollout", 0.0))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = eng.launch_count()
    ms_total = timed(mpc_step, args.steps, args.warmup)
    launches = eng.launch_count() - launches0
    ms_per_step = ms_total / args.steps
    value = K_total * HORIZON / (ms_per_step * 1e-3)

    # ---- e2e: host action samples in pinned memory through the C ABI -------------------------
    n_act = K_PER_GPU * HORIZON
    pinned = torch.empty(n_act, dtype=torch.float64).pin_memory()
    host_actions = pinned.numpy().reshape(K_PER_GPU, HORIZON, 1)
    host_actions[:] = np.random.RandomState(rank).uniform(wl["low"], wl["high"], (K_PER_GPU, HORIZON, 1))

    def e2e_step(i):
        # every rank feeds its own K_PER_GPU host samples through the C ABI (H2D inside the call)
        planner.plan(wl["state"], 0, K=K_total, H=HORIZON, local_actions=host_actions, penalty_mode="reference",
                     precision=precision, want_path=True)

    e2e_ms = timed(e2e_step, max(3, args.steps // 2), 2) / max(3, args.steps // 2)
    e2e_value = K_total * HORIZON / (e2e_ms * 1e-3)

    # ---- the same decision with the per-sample projection (penalty_mode=1: the evidently intended
    # maths; fully fused, no trajectory spill, no second pass) -----------------------
    def mpc_per_sample_step(i):
        planner.plan(wl["state"], 0, K=K_total, H=HORIZON, seed=3000 + i, act_low=wl["low"], act_high=wl["high"],
                     penalty_mode="per_sample", precision=precision, want_path=True)

    ps_steps = max(3, args.steps // 2)
    per_sample_ms = timed(mpc_per_sample_step, ps_steps, 2) / ps_steps

    # ---- KDE (BASELINE config 2 per GPU) -----------------------------------------------------
    kw = kde_workload()
    d_data = torch.as_tensor(kw["all_states"], device=dev)
    d_q = torch.as_tensor(kw["queries"], device=dev)
    d_v = torch.as_tensor(kw["values"], device=dev)
    pairs_ms = []

    def kde_step(i):
        eng.select_start_dev(d_data.data_ptr(), d_data.shape[0], 3, d_q.data_ptr(), d_q.shape[0], d_v.data_ptr(),
                             kw["n"], kw["volume"], 1.0, 2.0)
        if i > 0:
            pairs_ms.append(dict(eng.last_timings()).get("kde_pairs", 0.0))

    kde_ms = timed(kde_step, args.steps, args.warmup) / args.steps
    evals = KDE_M * (KDE_N + 1) * world
    kde_value = evals / (kde_ms * 1e-3)

    def kde_e2e_step(i):
        eng.select_start(kw["all_states"], kw["queries"], kw["values"], kw["n"], kw["volume"], 1.0, 2.0)

    kde_e2e_ms = timed(kde_e2e_step, max(3, args.steps // 2), 2) / max(3, args.steps // 2)

    # the same selection from the device mirror of the replay buffer's state ring (row f2): the
    # buffer is resident, each step uploads the 64 newest states + the candidate indices and values
    from smartstartcontinuous_b200.replay_buffer import _StateRing
    ring = _StateRing(KDE_N, 3)
    st_all = kw["all_states"]
    ring.alloc = KDE_N
    ring.s = np.ascontiguousarray(st_all[:KDE_N])
    ring.s2 = np.ascontiguousarray(np.concatenate([st_all[1:KDE_N], st_all[KDE_N:KDE_N + 1]]))
    ring.count = ring.pushes = KDE_N
    cand = np.random.default_rng(0).choice(KDE_N, KDE_M, replace=False)
    eng.mirror_sync(ring)

    def kde_mirror_step(i):
        # 64 transitions arrived since the last selection (written like ReplayBuffer.add does, as one
        # block: the per-step Python cost of add() belongs to the environment loop, not to selection)
        r0 = ring.pushes % KDE_N                          # the ring is full: the oldest rows are overwritten
        if r0 + 64 > KDE_N:
            r0 = 0
            ring.pushes += KDE_N - ring.pushes % KDE_N
        ring.s[r0:r0 + 64] = st_all[r0:r0 + 64]
        ring.s2[r0:r0 + 64] = st_all[r0 + 1:r0 + 65]
        ring.pushes += 64
        ring.head = ring.pushes % KDE_N
        eng.select_start_mirror(ring, cand, kw["values"], kw["n"], kw["volume"], 1.0, 2.0)

    kde_mirror_ms = timed(kde_mirror_step, max(3, args.steps // 2), 2) / max(3, args.steps // 2)
    # ---- BASELINE config 5 (strong scaling): 1 000 001-state buffer KDE with the 16 384 candidates
    # sharded over the ranks + MPC with K = 262 144 sequences in total, both device-resident ----------
    C5_K = 262_144
    kw5 = kde_workload(seed=1, n=1_000_000)
    d5_data = torch.as_tensor(kw5["all_states"], device=dev)
    d5_q = torch.as_tensor(kw5["queries"], device=dev)
    d5_v = torch.as_tensor(kw5["values"], device=dev)

    def c5_kde_step(i):
        selector.select_start_dev(d5_data.data_ptr(), d5_data.shape[0], 3, d5_q.data_ptr(), d5_q.shape[0],
                                  d5_v.data_ptr(), kw5["n"], kw5["volume"], 1.0, 2.0)

    def c5_mpc_step(i):
        planner.plan(wl["state"], 0, K=C5_K, H=HORIZON, seed=2000 + i, act_low=wl["low"], act_high=wl["high"],
                     penalty_mode="reference", precision=precision, want_path=True)

    c5_kde_ms = timed(c5_kde_step, 5, 3) / 5
    c5_mpc_ms = timed(c5_mpc_step, 5, 3) / 5
    del d5_data, d5_q, d5_v
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    info = eng.device_info()
    k_ms = float(np.mean(rollout_ms)) if rollout_ms else ms_per_step
    achieved_tf = FLOP_PER_STEP[3] * K_PER_GPU * HORIZON / (k_ms * 1e-3) / 1e12
    p_ms = float(np.mean(pairs_ms)) if pairs_ms else kde_ms
    sfu_peak = SFU_PER_CLK_PER_SM * info["sm_count"] * peaks["sm_max_mhz"] * 1e6
    kde_achieved = KDE_M * (KDE_N + 1) / (p_ms * 1e-3)
    kde_bytes = 4 * 3 * (KDE_N + 1 + KDE_M) + 4 * KDE_M + 8
    out = {
        "metric": "mpc_rollout_steps_per_s", "value": value, "unit": "rollout-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 (tcgen05, fp32 accumulate; first/last layer, state, scoring fp32)" if precision == "bf16_tc" else "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(),
                   "K_total": K_total, "H": HORIZON, "mlp": "2x500", "actions": "device Philox4x32-10",
                   "l2": "flushed between timed steps (256 MiB memset, outside the event-timed region)",
                   "parallelism": "K sharded over %d GPU(s); all-reduce of %d float64 + all-gather of the winner packages, %s"
                                  % (world, 2 * (HORIZON + 1),
                                     "fused into the kernels over NVLink peer memory (csrc/peer.cu)" if eng.peer_ready
                                     else ("NCCL" if world > 1 else "none at N=1"))},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "rollout-steps/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(n_act * 8 + 3 * 8), "d2h_bytes_per_step": int(16 + HORIZON * 8 + (HORIZON + 1) * 3 * 8),
                "path": "ShardedPlanner.plan -> ss_mpc_rollout / ss_mpc_finish_package with HOST float64 action samples "
                        "(pinned, uploaded in chunks that overlap the rollout) + D2H of the winner package; the agent's "
                        "default call (device Philox sampling: 24 B of state in, the same package out) is what `value` times"},
        "gpu_launches": int(launches),
        "per_sample_penalty": {"value": K_total * HORIZON / (per_sample_ms * 1e-3), "unit": "rollout-steps/s",
                               "ms_per_step": per_sample_ms,
                               "note": "penalty_mode=per_sample (fully fused scoring, no penalty passes); measured after the headline and e2e legs"},
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": achieved_tf / peaks["bf16_sustained"],
                     "traffic": NCU_TRAFFIC["mpc_rollout_tc_kernel"] if precision == "bf16_tc" else None,
                     "traffic_source": "profiles/r01e_ncu_summary.md (ncu --set full, same workload)",
                     "frac_of_burst_peak": achieved_tf / peaks["bf16_burst"],
                     "kernel": "mpc_rollout_tc_kernel" if precision == "bf16_tc" else "mpc_rollout_simt_kernel",
                     "kernel_ms": k_ms, "peak_source": peaks["source"] + " (bf16_tflops_sustained: kernel timed inside a long step)",
                     "algorithmic_flop_per_rollout_step": FLOP_PER_STEP[3]},
        "kde": {"metric": "kde_kernel_evals_per_s", "value": kde_value, "unit": "kernel-evals/s", "ms_per_step": kde_ms,
                "config": {"workload": "KDE+UCB+argmax, %d Pendulum states (d=3) x %d queries per GPU (BASELINE config 2)"
                                       % (KDE_N + 1, KDE_M)},
                "e2e": {"value": evals / (kde_e2e_ms * 1e-3), "unit": "kernel-evals/s", "ms_per_step": kde_e2e_ms,
                        "h2d_bytes_per_step": int(8 * 3 * (KDE_N + 1 + KDE_M) + 4 * KDE_M), "d2h_bytes_per_step": 16},
                "e2e_mirror": {"value": evals / (kde_mirror_ms * 1e-3), "unit": "kernel-evals/s", "ms_per_step": kde_mirror_ms,
                               "h2d_bytes_per_step": int(2 * 64 * 3 * 8 + 12 * KDE_M), "d2h_bytes_per_step": 16,
                               "path": "Engine.select_start_mirror: replay-state ring mirrored on the device, 64 new "
                                       "transitions + candidate indices + values uploaded per selection"},
                "roofline": {"bound": "sfu", "achieved": kde_achieved, "peak": sfu_peak, "unit": "kernel-evals/s",
                             "frac": kde_achieved / sfu_peak, "kernel": "kde_pairs_tc_kernel", "kernel_ms": p_ms,
                             "peak_source": "16 MUFU.EX2/clk/SM x %d SMs x %.0f MHz (max SM clock); the kernel takes 3/8 of "
                                            "its exp2 evaluations on the FMA pipe (polynomial) and the exponents from the "
                                            "tensor pipe, so frac > 1 is possible" % (info["sm_count"], peaks["sm_max_mhz"]),
                             "hbm_achieved_gbs": kde_bytes / (p_ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks["hbm_gbs"],
                             "traffic": NCU_TRAFFIC["kde_pairs_tc_kernel"],
                             "traffic_source": "profiles/r01e_ncu_summary.md (ncu --set full, same workload)"}},
    }
    out["config5"] = {"workload": "BASELINE config 5, strong scaling over %d GPU(s): KDE 1 000 001 states x %d queries "
                                  "(queries sharded) + MPC K=%d total, H=%d (sequences sharded), device-resident"
                                  % (world, KDE_M, C5_K, HORIZON),
                      "kde_ms": c5_kde_ms, "kde_evals_per_s": KDE_M * 1_000_001 / (c5_kde_ms * 1e-3),
                      "mpc_ms": c5_mpc_ms, "mpc_rollout_steps_per_s": C5_K * HORIZON / (c5_mpc_ms * 1e-3),
                      "episode_ms": c5_kde_ms + c5_mpc_ms}
    # ---- row f3 (plan set-up): pair extraction of path_shortcutter, P = 1000 path states ----------
    try:
        from smartstartcontinuous_b200 import numerical as num
        from smartstartcontinuous_b200 import synthetic as syn
        obs, _ = syn.pendulum_rollouts(np.random.default_rng(7), 1, 1000)
        path = np.asarray(obs[0][:1000])
        stds, means = num.path_deltas_stds_and_means_per_dim(path)
        radii = num.radii_calc(means, stds, 1, 1, 1)
        dist_fn = num.elliptical_euclidean_distance_function_generator(radii)
        eng.path_close_pairs(path, radii, 1.0)
        t0 = time.perf_counter()
        for _ in range(10):
            pairs = eng.path_close_pairs(path, radii, 1.0)
        t_dev = (time.perf_counter() - t0) / 10
        t0 = time.perf_counter()
        for _ in range(3):
            ref_pairs = np.argwhere(np.triu(dist_fn(path[:, None, :], path[None, :, :]) <= 1.0, k=2))
        t_np = (time.perf_counter() - t0) / 3
        eng.path_shortcut(path, radii, 1.0)
        t0 = time.perf_counter()
        for _ in range(10):
            keep = eng.path_shortcut(path, radii, 1.0)
        t_full_dev = (time.perf_counter() - t0) / 10
        t0 = time.perf_counter()
        host_path = num.path_shortcutter(path, dist_fn, 1.0)
        t_full_host = time.perf_counter() - t0
        out["plan_setup"] = {"what": "path_shortcutter (numerical.py:226-246), P=1000, d=3, host buffers in and out",
                             "pairs_gpu_ms": 1e3 * t_dev, "pairs_numpy_ms": 1e3 * t_np, "pairs": int(len(pairs)),
                             "pairs_identical": bool(np.array_equal(pairs, ref_pairs)),
                             "shortcut_gpu_ms": 1e3 * t_full_dev, "shortcut_host_ms": 1e3 * t_full_host,
                             "shortcut_identical": bool(np.array_equal(path[keep], host_path)),
                             "kept_states": int(len(keep))}
    except Exception as exc:
        out["plan_setup"] = {"error": repr(exc)}
    if world == 1 and not args.no_cpu_baseline:
        # the CPU leg runs in a fresh process (fork-based pool, no CUDA context): the same code as
        # `--impl reference`, ~20 s of CPU work on bounded samples of the workload
        try:
            ref = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "16",
                                  "--warmup", "1"], capture_output=True, text=True, timeout=600)
            r = json.loads(ref.stdout.strip().splitlines()[-1])
            out["cpu_baseline"] = r["cpu_baseline"]
            out["kde"]["cpu_baseline"] = {k: r["kde"][k] for k in ("value", "unit", "cores", "kind", "sample", "single_core")}
        except Exception as exc:                                     # never lose the GPU line
            out["cpu_baseline"] = {"value": None, "unit": "rollout-steps/s", "cores": 0, "kind": "port",
                                   "sample": "failed: %r" % (exc,)}
    guard.emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
