"""Informed RRT* batch (rrt_07): kernel ms by batch size (bench's scenario)."""
import sys
sys.path.insert(0, "/root/repo/robotics-path-planning_b200")
import numpy as np
from rrtk import informed as INF
iters = 1000
for Q in (64, 296, 512, 592, 1024):
    rng = np.random.default_rng(8)
    free = rng.uniform(-2, 15, (Q, iters, 2)); coin = rng.integers(0, 101, (Q, iters)) <= 10
    free[coin] = (6.0, 10.0)
    ball = rng.random((Q, iters, 2))
    obs = [[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]] * Q
    outs = {}
    for mode in ("warp", "cta"):
        tm = {}
        for rep in range(2):
            res = INF.run_batch([[0.0, 0.0]] * Q, [[6.0, 10.0]] * Q, obs, 0.5, iters, free, ball, timing=tm, exec_mode=mode)
        outs[mode] = res
        print("informed Q=%d x %d %s: kernel %.1f ms  %.1f M it/s  mean nodes %.0f" % (Q, iters, mode, tm["kernel_ms"],
              Q * iters / tm["kernel_ms"] / 1e3, np.mean([r["n"] for r in res])), flush=True)
    same = all(a["n"] == b["n"] and np.array_equal(a["parent"], b["parent"]) and np.array_equal(a["cost"], b["cost"]) and
               np.array_equal(a["x"], b["x"]) and a["path"] == b["path"] for a, b in zip(outs["warp"], outs["cta"]))
    print("   identical:", same, flush=True)
