"""astar_torus over 64 occupancy grids (bench extras' case): ms, expansions, a hash of routes + marks (A/B between builds)."""
import math, sys, hashlib
sys.path.insert(0, "/root/repo/robotics-path-planning_b200")
import numpy as np, torch
from rrtk import arm as A
M, S = 512, 64
rng = np.random.default_rng(15)
ang, rad = rng.uniform(0, 2 * math.pi, (S, 5)), rng.uniform(0.9, 2.0, (S, 5))
sets = np.stack([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(0.15, 0.45, (S, 5))], axis=2)
base = A.occupancy_grids_device([1.0, 1.0], sets, M)
host = base.cpu().numpy()
st, gl = [], []
for k in range(S):
    free = np.argwhere(host[k] == 0)
    st.append(free[rng.integers(len(free))]); gl.append(free[rng.integers(len(free))])
for rep in range(3):
    grids = base.clone()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    routes, rlen, expanded = A.astar_torus_batch(grids, np.array(st), np.array(gl))
    b.record(); torch.cuda.synchronize()
    hsh = hashlib.sha1(grids.cpu().numpy().tobytes() + routes.cpu().numpy().tobytes() + rlen.cpu().numpy().tobytes()).hexdigest()[:16]
    print("astar 64 x M=512: %.1f ms  expansions %d  max %d  routes found %d  hash %s" % (
        a.elapsed_time(b), int(expanded.sum().item()), int(expanded.max().item()), int((rlen > 0).sum().item()), hsh), flush=True)
