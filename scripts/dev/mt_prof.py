import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from smartstartcontinuous_b200.engine import Engine
eng = Engine(0)
for n in (81920, 6553600):
    rs = np.random.RandomState(1); rs.random_sample(100)
    st = rs.get_state()
    for _ in range(3):
        eng.mt19937_uniform(st, n, [-2.0], [2.0]); eng.mt19937_state()
