import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from smartstartcontinuous_b200.distributed import ShardedPlanner, _DevView, argmax_pick
from smartstartcontinuous_b200.engine import Engine
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = Engine(local); eng.set_stream(torch.cuda.current_stream().cuda_stream)
wl = bench.make_workload()
eng.set_model(wl["w"], wl["b"], wl["norm"]); eng.set_plan(wl["plan"]["desired_states"], wl["plan"]["distances_left"], wl["plan"]["radii"])
K = bench.K_PER_GPU * world; H = bench.HORIZON
def sync(): torch.cuda.synchronize()
for it in range(6):
    T = {}
    sync(); dist.barrier(); sync()
    t0 = time.perf_counter()
    eng.rollout(wl["state"], 0, K=bench.K_PER_GPU, H=H, seed=it, act_low=wl["low"], act_high=wl["high"], penalty_mode="reference", precision="bf16_tc", k_offset=rank * bench.K_PER_GPU, K_global=K)
    sync(); T["rollout"] = time.perf_counter() - t0; t0 = time.perf_counter()
    ptr, n = eng.projection_sums_ptr(); sums = torch.as_tensor(_DevView(ptr, n), device=dev)
    T["view"] = time.perf_counter() - t0; t0 = time.perf_counter()
    dist.all_reduce(sums); sync(); T["allreduce"] = time.perf_counter() - t0; t0 = time.perf_counter()
    bk, bs, _ = eng.finish(); T["finish"] = time.perf_counter() - t0; t0 = time.perf_counter()
    mine = torch.tensor([bs, float(bk)], dtype=torch.float64, device=dev); g = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(g, mine); pairs = torch.stack(g).cpu().numpy(); T["allgather"] = time.perf_counter() - t0; t0 = time.perf_counter()
    w = argmax_pick(pairs[:, 0].tolist(), [int(v) for v in pairs[:, 1]])
    seq, path = eng.replay(int(pairs[w, 1])); T["replay(all ranks)"] = time.perf_counter() - t0
    if rank == 0: print(it, {k: round(v * 1e3, 3) for k, v in T.items()})
dist.destroy_process_group()
