"""Development helper: wall / event time of the host-buffer KDE selection, with and without the bench's
stream binding and L2 flush."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
import bench
from smartstartcontinuous_b200.engine import Engine
eng = Engine(0)
kw = bench.kde_workload()
def run(tag, flush=None, stream=None):
    ts, evs = [], []
    for i in range(8):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        eng.select_start(kw["all_states"], kw["queries"], kw["values"], kw["n"], kw["volume"], 1.0, 2.0)
        ts.append(1e3 * (time.perf_counter() - t0))
        e1.record(stream)
        evs.append((e0, e1))
    torch.cuda.synchronize()
    print(tag, "wall", np.round(ts[2:], 3), "events", np.round([a.elapsed_time(b) for a, b in evs[2:]], 3), eng.last_timings())
run("own stream          ")
stream = torch.cuda.current_stream()
eng.set_stream(stream.cuda_stream)
run("torch stream        ", stream=stream)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
run("torch stream + flush", flush=flush, stream=stream)
