"""One RRT*-Reeds-Shepp launch (rrt_06 semantics) for profiling:  python tools/probe_rs_one.py [Q] [iters] [reps]"""
import math, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "robotics-path-planning_b200"))
import numpy as np
from rrtk import rs_planner as RS
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rng = np.random.default_rng(8)
st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
st[rng.integers(0, 101, (Q, iters)) <= 10] = (6.0, 7.0, math.pi / 2)
obs = [[(5, 5, 1), (4, 6, 1), (4, 8, 1), (4, 10, 1), (6, 5, 1), (7, 5, 1), (8, 6, 1), (8, 8, 1), (8, 10, 1)]] * Q
tm = {}
for rep in range(reps):
    res = RS.run_batch([[0.0, 0.0, 0.0]] * Q, [[6.0, 7.0, math.pi / 2]] * Q, obs, 3.0, iters, st, timing=tm)
    print(Q, iters, "kernel ms %.2f" % tm["kernel_ms"], "mean nodes %.0f" % np.mean([r["n"] for r in res]), flush=True)
