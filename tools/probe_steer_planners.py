"""Kernel times of the steering planners on the bench workloads (RRT*-Dubins c4, RRT*-Reeds-Shepp, closed-loop RRT* trees)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-path-planning_b200"))
import numpy as np
from rrtk import dubins_planner as DP, rs_planner as RP
Q, iters = 1024, 500
rng = np.random.default_rng(7)
st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
coin = rng.integers(0, 101, (Q, iters)) <= 10
st[coin] = (10.0, 10.0, 0.0)
obs = [[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2)]] * Q
tm = {}
for rep in range(3):
    res = DP.run_batch([[0.0, 0.0, 0.0]] * Q, [[10.0, 10.0, 0.0]] * Q, obs, 3.0, iters, st, timing=tm)
    print("c4 dubins kernel ms %.1f" % tm["kernel_ms"], "mean nodes", np.mean([r["n"] for r in res]), flush=True)
Q, iters = 512, 300
rng = np.random.default_rng(17)
st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
obs = [[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]] * Q
for rep in range(2):
    res = RP.run_batch([[0.0, 0.0, 0.0]] * Q, [[10.0, 9.0, 0.0]] * Q, obs, 3.0, iters, st, robot_radius=0.6, curvature=2.0, step_size=0.1, timing=tm)
    print("rs kernel ms %.1f" % tm["kernel_ms"], "mean nodes", np.mean([r["n"] for r in res]), flush=True)
Q, iters = 256, 100
rng = np.random.default_rng(19)
st = np.concatenate([rng.uniform(-2, 20, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
for rep in range(2):
    res = RP.run_batch([[0.0, 0.0, 0.0]] * Q, [[6.0, 9.0, math.radians(90.0)]] * Q, obs[:Q], float("inf"), iters, st, curvature=1.0, step_size=0.2, near_cap=224, timing=tm, rs_cost=True)
    print("closed-loop planner kernel ms %.1f" % tm["kernel_ms"], "mean nodes", np.mean([r["n"] for r in res]), flush=True)
