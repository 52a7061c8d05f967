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
print("ok")
