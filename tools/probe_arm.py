"""Config 5 (arm C-space grid, M = 8192, 64 obstacle sets, the script's 5-link arm): kernel ms, occupied cells of set 0."""
import sys
sys.path.insert(0, "/root/repo/robotics-path-planning_b200")
import numpy as np, torch
from rrtk import _lib, arm as A
L = _lib.lib()
dev = torch.device("cuda")
M, S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 64
rng = np.random.default_rng(5)
sets = np.concatenate([rng.uniform(-2, 2, (S, 5, 2)), rng.uniform(0.2, 0.7, (S, 5, 1))], axis=2)
sets[0] = [[1.75, 0.75, 0.6], [0.55, 1.5, 0.5], [0, -1, 0.7], [0, -0.6, 0.4], [-1, 1., 0.3]]
link = np.array([0.5, 0.5, 0.3, 0.5, 0.1])
theta = torch.from_numpy(A.theta_list(M)).to(dev)
d_obs = torch.from_numpy(sets).to(dev)
grid = torch.empty((S, M, M), dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    rc = L.rrtk_arm_grid_dev(M, theta.data_ptr(), 0, M, 5, link.ctypes.data, d_obs.data_ptr(), S, 5, grid.data_ptr(), s)
    b.record(); b.synchronize()
    assert rc == 0
    print("arm grid M=%d S=%d  %.1f ms  %.1f Gcell/s" % (M, S, a.elapsed_time(b), M * M * S / a.elapsed_time(b) / 1e6), flush=True)
print("occupied per set (first 8):", [int(grid[k].sum().item()) for k in range(8)], "total", int(grid.sum(dtype=torch.int64).item()))
