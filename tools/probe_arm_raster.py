"""Row rasteriser (rrtk_arm_grid_dev) against the per-cell kernel (rrtk_arm_grid_cells_dev): bitwise, over grid sizes, arms,
obstacle sets (random, tangent to the reach / to joint 1, radius 0, far away), row shards; then config-5 timing."""
import sys
sys.path.insert(0, "/root/repo/robotics-path-planning_b200")
import numpy as np, torch
from rrtk import _lib, arm as A
L = _lib.lib()
dev = torch.device("cuda")
st = torch.cuda.current_stream().cuda_stream

def both(M, link, sets, row0=0, n_rows=None, theta=None):
    n_rows = M - row0 if n_rows is None else n_rows
    th = torch.from_numpy(A.theta_list(M) if theta is None else theta).to(dev)
    d_obs = torch.from_numpy(np.ascontiguousarray(sets, dtype=np.float64)).to(dev)
    link = np.asarray(link, dtype=np.float64)
    S, O = sets.shape[0], sets.shape[1]
    out = []
    for fn in (L.rrtk_arm_grid_dev, L.rrtk_arm_grid_cells_dev):
        g = torch.full((S, n_rows, M), 7, dtype=torch.uint8, device=dev)
        rc = fn(M, th.data_ptr(), row0, n_rows, len(link), link.ctypes.data, d_obs.data_ptr(), S, O, g.data_ptr(), st)
        assert rc == 0, _lib.last_error()
        torch.cuda.synchronize()
        out.append(g)
    return out

bad = 0
rng = np.random.default_rng(11)
cases = []
for M in (64, 100, 101, 257, 1000, 1024, 2048):
    for link in ([0.5, 0.5, 0.3, 0.5, 0.1], [1.0, 1.0], [0.7], [0.3, 1.2, 0.2], [0.5, 0.5, 0.3, 0.5, 0.1, 0.2, 0.2, 0.4]):
        S, O = 16, 5
        sets = np.concatenate([rng.uniform(-2.5, 2.5, (S, O, 2)), rng.uniform(0.05, 0.9, (S, O, 1))], axis=2)
        Ls, l0 = float(np.sum(link[1:])), link[0]
        # tangent cases: circles just touching the reach circle around a joint-1 position, circles through the origin,
        # radius 0, far away, containing the whole arm
        ang = rng.uniform(-np.pi, np.pi)
        p1 = np.array([l0 * np.cos(ang), l0 * np.sin(ang)])
        u = rng.normal(size=2); u /= np.linalg.norm(u)
        sets[1, 0] = [*(p1 + u * (Ls + 0.4)), 0.4]                       # tangent to the reach from that joint-1 position
        sets[2, 0] = [*(p1 + u * (Ls + 0.4)), 0.4 * (1 + 1e-12)]
        sets[3, 0] = [*(p1 + u * 0.3), 0.3]                               # joint 1 on the circle
        sets[4, 0] = [0.4, 0.0, 0.4]                                      # the base on the circle
        sets[5, :, 2] = 0.0                                               # points
        sets[6, :, :2] += 50.0                                            # far away
        sets[7, 0] = [0.0, 0.0, 10.0]                                     # everything inside
        sets[8, 0] = [*(p1 + u * np.sqrt(max(Ls * Ls + 0.25, 0))), 0.5]   # the far end reaches exactly the tangent point
        cases.append((M, link, sets))
for M, link, sets in cases:
    a, b = both(M, link, sets)
    if not torch.equal(a, b):
        bad += 1
        diff = (a != b).nonzero()
        print("MISMATCH M=%d links=%d cells=%d first=%s" % (M, len(link), diff.shape[0], diff[0].tolist()), flush=True)
# row shards and a theta array that is not the reference's (falls back to cell-by-cell, still equal)
M, link = 1000, [0.5, 0.5, 0.3, 0.5, 0.1]
sets = np.concatenate([rng.uniform(-2, 2, (8, 5, 2)), rng.uniform(0.2, 0.7, (8, 5, 1))], axis=2)
a, b = both(M, link, sets, row0=333, n_rows=200)
bad += 0 if torch.equal(a, b) else 1
th = A.theta_list(M).copy(); th[500] += 1e-3
a, b = both(M, link, sets, theta=th)
bad += 0 if torch.equal(a, b) else 1
a, b = both(M, [0.5, 0.0, 0.3], sets)        # a zero-length link: the reference's 0 / 0
bad += 0 if torch.equal(a, b) else 1
print("cases", len(cases) + 3, "mismatching", bad, flush=True)

# config 5
M, S = 8192, 64
rng = np.random.default_rng(5)
sets = np.concatenate([rng.uniform(-2, 2, (S, 5, 2)), rng.uniform(0.2, 0.7, (S, 5, 1))], axis=2)
sets[0] = [[1.75, 0.75, 0.6], [0.55, 1.5, 0.5], [0, -1, 0.7], [0, -0.6, 0.4], [-1, 1., 0.3]]
link = np.array([0.5, 0.5, 0.3, 0.5, 0.1])
theta = torch.from_numpy(A.theta_list(M)).to(dev)
d_obs = torch.from_numpy(sets).to(dev)
grids = []
for fn, name in ((L.rrtk_arm_grid_dev, "rows"), (L.rrtk_arm_grid_cells_dev, "cells")):
    grid = torch.empty((S, M, M), dtype=torch.uint8, device=dev)
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(M, theta.data_ptr(), 0, M, 5, link.ctypes.data, d_obs.data_ptr(), S, 5, grid.data_ptr(), st)
        e1.record(); e1.synchronize()
        assert rc == 0
        ms = e0.elapsed_time(e1)
        print("%s M=%d S=%d  %.2f ms  %.0f Gcell/s  %.0f GB/s written" % (name, M, S, ms, M * M * S / ms / 1e6, M * M * S / ms / 1e6), flush=True)
    grids.append(grid)
print("config 5 identical:", torch.equal(grids[0], grids[1]), "occupied", int(grids[0].sum(dtype=torch.int64).item()))
# the store stream's ceiling on this box: a plain fill of the same 4.3 GB
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); grids[1].fill_(1); e1.record(); e1.synchronize()
    print("torch fill_ of the same tensor: %.2f ms  %.0f GB/s" % (e0.elapsed_time(e1), M * M * S / e0.elapsed_time(e1) / 1e6), flush=True)
