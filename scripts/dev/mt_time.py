"""Development helper: time of ss_mt19937_uniform (+ state read-back) at the benchmarked stream lengths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import time, numpy as np, ctypes as C
from smartstartcontinuous_b200.engine import Engine
import torch
eng = Engine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
for n in (5000*4, 4096*20, 20000*12, 131072*50, 262144*50):
    rs = np.random.RandomState(1); rs.random_sample(100)
    st = rs.get_state()
    t0=time.perf_counter(); ptr = eng.mt19937_uniform(st, n, [-2.0], [2.0]); eng.mt19937_state(); t1=time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10): ptr = eng.mt19937_uniform(st, n, [-2.0], [2.0])
    ev1.record(); torch.cuda.synchronize()
    t2=time.perf_counter()
    for _ in range(10):
        ptr = eng.mt19937_uniform(st, n, [-2.0], [2.0]); eng.mt19937_state()
    t3=time.perf_counter()
    print("n=%d first call %.1f ms, kernel %.1f us, call+state wall %.1f us" % (n, (t1-t0)*1e3, ev0.elapsed_time(ev1)*100, (t3-t2)*1e5))
