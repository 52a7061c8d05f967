import math, os, sys
sys.path.insert(0, "/root/repo/robotics-path-planning_b200")
import numpy as np
from rrtk import dubins_planner as DP
iters = 500
for Q in (64, 256, 1024, 4096):
    rng = np.random.default_rng(7)
    st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
    st[rng.integers(0, 101, (Q, iters)) <= 10] = (10.0, 10.0, 0.0)
    obs = [[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2)]] * Q
    tm = {}
    for rep in range(2):
        res = DP.run_batch([[0.0, 0.0, 0.0]] * Q, [[10.0, 10.0, 0.0]] * Q, obs, 3.0, iters, st, timing=tm)
    n = np.array([r["n"] for r in res])
    print(Q, "kernel ms %.1f" % tm["kernel_ms"], "nodes mean %.0f max %d p90 %d" % (n.mean(), n.max(), np.percentile(n, 90)), flush=True)
