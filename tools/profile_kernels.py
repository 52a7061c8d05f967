"""Launch each secondary kernel a few times on realistic sizes (for `ncu`; timings printed here are not bench values).

    python tools/profile_kernels.py            # plain run
    ncu --set full -k regex:'nearest_kernel|arm_grid|dubins_steer|informed_kernel|rrtstar_dubins' ... python tools/profile_kernels.py
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-path-planning_b200"))
import torch  # noqa: E402
from rrtk import _lib, arm as A, dubins, dubins_planner as DP, informed as INF  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda")
s = torch.cuda.current_stream().cuda_stream
REPS = int(os.environ.get("RRTK_PROFILE_REPS", "3"))

# NN search, 2^26 float2 nodes (537 MB), B = 1 and 8
n = 1 << 26
xy = torch.rand((n, 2), dtype=torch.float32, device=dev) * 17 - 2
for B in (1, 8):
    smp = torch.rand((B, 2), dtype=torch.float32, device=dev) * 17 - 2
    scratch = torch.empty(B, dtype=torch.int64, device=dev)
    idx = torch.empty(B, dtype=torch.int32, device=dev)
    d2 = torch.empty(B, dtype=torch.float32, device=dev)
    for _ in range(REPS):
        _lib.check(L.rrtk_nearest_f32_dev(xy.data_ptr(), n, smp.data_ptr(), B, scratch.data_ptr(), idx.data_ptr(),
                                          d2.data_ptr(), s))
out = torch.empty(1 << 20, dtype=torch.int32, device=dev)
cnt = torch.zeros(1, dtype=torch.int32, device=dev)
for _ in range(REPS):
    _lib.check(L.rrtk_near_f32_dev(xy.data_ptr(), n, 6.5, 7.25, 0.01, out.data_ptr(), 1 << 20, cnt.data_ptr(), s))
torch.cuda.synchronize()
del xy

# arm grid, M = 4096, 64 sets
M, S = 4096, 64
rng = np.random.default_rng(5)
sets = np.concatenate([rng.uniform(-2, 2, (S, 5, 2)), rng.uniform(0.2, 0.7, (S, 5, 1))], axis=2)
for _ in range(REPS):
    g = A.occupancy_grids_device([0.5, 0.5, 0.3, 0.5, 0.1], sets, M)
torch.cuda.synchronize()
del g

# Dubins steering, 65536 edges among 16 circles
ne = 1 << 16
f = np.column_stack([rng.uniform(0, 12, (ne, 2)), rng.uniform(-math.pi, math.pi, ne)])
t = np.column_stack([f[:, 0:2] + rng.uniform(-4, 4, (ne, 2)), rng.uniform(-math.pi, math.pi, ne)])
obs = [[(float(x), float(y), float(r)) for (x, y), r in zip(rng.uniform(0, 12, (16, 2)), rng.uniform(0.2, 0.8, 16))]]
for _ in range(REPS):
    dubins.steer_batch(f, t, 1.0, 0.1, obstacle_sets=obs)

# RRT*-Dubins, 1024 queries x 300 iterations
Q, iters = 1024, 300
st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
st[rng.integers(0, 101, (Q, iters)) <= 10] = (10.0, 10.0, 0.0)
ob = [[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2)]] * Q
for _ in range(REPS):
    DP.run_batch([[0.0, 0.0, 0.0]] * Q, [[10.0, 10.0, 0.0]] * Q, ob, 3.0, iters, st)

# Informed RRT*, 512 queries x 600 iterations
Q, iters = 512, 600
free = rng.uniform(-2, 15, (Q, iters, 2))
free[rng.integers(0, 101, (Q, iters)) <= 10] = (6.0, 10.0)
ball = rng.random((Q, iters, 2))
ob = [[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]] * Q
for _ in range(REPS):
    INF.run_batch([[0.0, 0.0]] * Q, [[6.0, 10.0]] * Q, ob, 0.5, iters, free, ball)
torch.cuda.synchronize()

# round-1 additions: single-tree informed kernel (batched and unbatched), Reeds-Shepp steering and planner, smoothing, astar
from rrtk import reeds_shepp as RS, rs_planner as RP, smoothing, workloads as W  # noqa: E402
import rrtk  # noqa: E402
it = 30000
fr = rng.uniform(-2, 15, (it, 2)); fr[rng.integers(0, 101, it) <= 10] = (6.0, 10.0); bl = rng.random((it, 2))
OB = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]
for b in (8, 1):
    INF.run_tree([0.0, 0.0], [6.0, 10.0], OB, 0.5, it, fr, bl, batch=b)
torch.cuda.synchronize()
RS.steer_batch(f, t, 1.0, 0.1, obstacle_sets=obs)
Q, iters = 256, 200
st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
RP.run_batch([[0.0, 0.0, 0.0]] * Q, [[10.0, 9.0, 0.0]] * Q, [OB] * Q, 3.0, iters, st, robot_radius=0.6, curvature=2.0, step_size=0.1)
cfg = W.C2
Q = 1024
rows = W.c2_rows(list(range(Q)), 256)
bt = rrtk.RRTStarBatch(np.tile(cfg["start"], (Q, 1)), np.tile(cfg["goal"], (Q, 1)), rows, cfg["rand_area"], cfg["expand_dis"],
                       cfg["path_resolution"], cfg["goal_sample_rate"], 1000, None, 0.0, "sobol", cfg["connect_circle_dist"], True, seed=1)
res = bt.run()
sp, sl = res.paths_device(64 + 500)
smoothing.smooth_batch(sp, sl, 500, bt.obstacles[:, :, :3].contiguous(), bt.n_obs)
ang, rad = rng.uniform(0, 2 * np.pi, (32, 5)), rng.uniform(0.9, 2.0, (32, 5))
sets2 = np.stack([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(0.15, 0.45, (32, 5))], axis=2)
grids = A.occupancy_grids_device([1.0, 1.0], sets2, 256)
host = grids.cpu().numpy()
stt, gl = [], []
for k in range(32):
    fc = np.argwhere(host[k] == 0)
    stt.append(fc[rng.integers(len(fc))]); gl.append(fc[rng.integers(len(fc))])
A.astar_torus_batch(grids, np.array(stt), np.array(gl))
torch.cuda.synchronize()
print("ok")
