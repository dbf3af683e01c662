"""torchrun --nproc-per-node N scripts/multigpu_check.py : the sharded planner / selector over N real
GPUs (NCCL) must reproduce the single-GPU decision (every rank also runs the full batch locally)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from smartstartcontinuous_b200.distributed import ShardedPlanner, ShardedSelector
from smartstartcontinuous_b200.engine import Engine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = Engine(local)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
wl = bench.make_workload()
eng.set_model(wl["w"], wl["b"], wl["norm"])
eng.set_plan(wl["plan"]["desired_states"], wl["plan"]["distances_left"], wl["plan"]["radii"])
planner = ShardedPlanner(eng, device=dev)
ok = True
if os.environ.get("SS_PEER", "1") != "0":
    eng.peer_setup()
if rank == 0:
    print("peer exchange:", eng.peer_ready)
for mode in ("reference", "per_sample"):
    for prec in ("fp32", "bf16_tc"):
        K, H = 20000, 12
        kw = dict(K=K, H=H, seed=7, act_low=wl["low"], act_high=wl["high"], penalty_mode=mode, precision=prec)
        sharded = planner.plan(wl["state"], 0, **kw)
        single = eng.plan(wl["state"], 0, **kw)
        same = sharded["best_k"] == single["best_k"] and abs(sharded["best_score"] - single["best_score"]) <= 1e-5 * max(1, abs(single["best_score"]))
        path_ok = np.allclose(sharded["best_path"], single["best_path"], rtol=1e-4, atol=1e-5)
        ok &= same and path_ok
        if rank == 0:
            print(mode, prec, "sharded", sharded["best_k"], round(sharded["best_score"], 6), "single", single["best_k"],
                  round(single["best_score"], 6), "owner", sharded["owner"], "OK" if same and path_ok else "MISMATCH")
kw = bench.kde_workload(n=20000, m=4096)
sel = ShardedSelector(eng, device=dev).select_start(kw["all_states"], kw["queries"], kw["values"], kw["n"], kw["volume"])
one = eng.select_start(kw["all_states"], kw["queries"], kw["values"], kw["n"], kw["volume"])
ok &= sel[0] == one[0]
if rank == 0:
    print("kde sharded", sel, "single", one[:2], "OK" if sel[0] == one[0] else "MISMATCH")
t = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU_CHECK", "PASS" if t.item() == 1.0 else "FAIL", "world", world)
dist.destroy_process_group()
