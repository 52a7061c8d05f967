"""One config-4 launch size (RRT*-Dubins, 500 iterations) for profiling:  python tools/probe_c4_one.py [Q] [reps]"""
import math, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "robotics-path-planning_b200"))
import numpy as np
from rrtk import dubins_planner as DP
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
iters = 500
rng = np.random.default_rng(7)
st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
st[rng.integers(0, 101, (Q, iters)) <= 10] = (10.0, 10.0, 0.0)
obs = [[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2)]] * Q
tm = {}
for rep in range(reps):
    res = DP.run_batch([[0.0, 0.0, 0.0]] * Q, [[10.0, 10.0, 0.0]] * Q, obs, 3.0, iters, st, timing=tm)
    print(Q, "kernel ms %.2f" % tm["kernel_ms"], flush=True)
