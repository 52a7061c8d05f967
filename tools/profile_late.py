"""One launch each of the closed-loop RRT* planner (rrtstar_dubins_kernel<2>), closed_loop_kernel and bitstar_kernel on
bench-sized inputs (for `ncu`; nothing printed here is a bench value)."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-path-planning_b200"))
import torch  # noqa: E402
from rrtk import bitstar as BS, closed_loop as CL, rs_planner as RP  # noqa: E402

Q, iters = 256, 100
rng = np.random.default_rng(19)
st = np.concatenate([rng.uniform(-2, 20, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
obs1 = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]
start, goal = [0.0, 0.0, 0.0], [6.0, 9.0, math.radians(90.0)]
trees = RP.run_batch([start] * Q, [goal] * Q, [obs1] * Q, float("inf"), iters, st, curvature=1.0, step_size=0.2, near_cap=224,
                     rs_cost=True)
courses = []
for tr in trees:
    gi = [i for i in range(tr["n"]) if math.hypot(tr["x"][i] - goal[0], tr["y"][i] - goal[1]) <= 0.5
          and abs(tr["yaw"][i] - goal[2]) <= math.radians(3.0)]
    courses += [np.asarray(c)[::-1] for c in CL.final_courses(tr, gi[:8], start, goal, 1.0, 0.2)]
res = CL.closed_loop_batch(courses, obs1)
Q, iters = 1024, 200
draws = np.random.default_rng(23).random((Q, 6000))
obs2 = [(5, 5, 0.5), (9, 6, 1), (7, 5, 1), (1, 5, 1), (3, 6, 1), (7, 9, 1)]
BS.run_batch([[-1.0, 0.0]] * Q, [[3.0, 8.0]] * Q, [obs2] * Q, [-2, 15], iters, draws)
# RRT*-Dubins (config 4) and RRT*-Reeds-Shepp on the bench workloads, astar_torus on 64 grids
from rrtk import arm as A, dubins_planner as DP  # noqa: E402
Q, iters = 1024, 500
rng = np.random.default_rng(7)
st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
st[rng.integers(0, 101, (Q, iters)) <= 10] = (10.0, 10.0, 0.0)
DP.run_batch([[0.0, 0.0, 0.0]] * Q, [[10.0, 10.0, 0.0]] * Q, [[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2)]] * Q,
             3.0, iters, st)
Q, iters = 512, 300
rng = np.random.default_rng(17)
st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
RP.run_batch([[0.0, 0.0, 0.0]] * Q, [[10.0, 9.0, 0.0]] * Q, [obs1] * Q, 3.0, iters, st, robot_radius=0.6, curvature=2.0, step_size=0.1)
rng = np.random.default_rng(3)
M, Q = 512, 64
ang, rad = rng.uniform(0, 2 * np.pi, (Q, 5)), rng.uniform(0.9, 2.0, (Q, 5))
sets = np.stack([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(0.15, 0.45, (Q, 5))], axis=2)
grids = A.occupancy_grids_device([1.0, 1.0], sets, M)
host = grids.cpu().numpy()
stt, gl = [], []
for k in range(Q):
    fc = np.argwhere(host[k] == 0); stt.append(fc[rng.integers(len(fc))]); gl.append(fc[rng.integers(len(fc))])
A.astar_torus_batch(grids, np.array(stt), np.array(gl))
torch.cuda.synchronize()
print("ok", len(courses))
