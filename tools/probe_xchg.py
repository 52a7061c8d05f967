import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-path-planning_b200"))
import torch
from rrtk import _lib
L = _lib.lib()
for grid in (0, 74, 16, 2):
    nb = L.rrtk_informed_tree_workspace_bytes(512, grid)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    out = torch.zeros(148, dtype=torch.int64, device="cuda")
    for _ in range(2):
        _lib.check(L.rrtk_tree_exchange_probe_dev(grid, 20000, out.data_ptr(), ws.data_ptr(), nb, None))
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    o = o[o > 0]
    print("grid", grid, "cycles/exchange min %d max %d" % (o.min(), o.max()), flush=True)
