"""Per-phase SM cycles of the CTA-per-query RRT* kernel (thread 0's clock64 between the CTA barriers), averaged over
queries and iteration ranges.  Needs the profiling build:  make -C robotics-path-planning_b200/csrc profile
    RRTK_LIB=robotics-path-planning_b200/rrtk/librrtk_prof.so python tools/probe_cta_phases.py [Q]"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "robotics-path-planning_b200"))
os.environ["RRTK_EXEC"] = "cta"
import rrtk
from rrtk import workloads as W

cfg = W.C2
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters, n_obs = 2000, 256
rows = W.c2_rows(list(range(Q)), n_obs)
starts = np.tile(np.array(cfg["start"]), (Q, 1)); goals = np.tile(np.array(cfg["goal"]), (Q, 1))
b = rrtk.RRTStarBatch(starts, goals, rows, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"],
                      cfg["goal_sample_rate"], iters, None, cfg["robot_radius"], "sobol", cfg["connect_circle_dist"],
                      True, seed=0xC2, sobol_offset=np.arange(Q, dtype=np.int64) * iters)
b.run(want_trace=True)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); r = b.run(want_trace=True); e.record(); torch.cuda.synchronize()
tr = r.trace.cpu().numpy().astype(np.float64)          # [Q, iters, 8]
names = ["scan", "B1", "first edge", "B2 wait", "choose_parent", "B3 wait", "rewire edges+B4", "apply+append"]
print(f"Q={Q} launch {a.elapsed_time(e):.2f} ms; cycles per iteration by phase (mean over queries)")
for lo, hi in ((0, 200), (200, 1000), (1000, 2000), (0, 2000)):
    m = tr[:, lo:hi].mean(axis=(0, 1))
    print(f"  it {lo:4d}-{hi:4d}: " + "  ".join(f"{nm} {v:7.0f}" for nm, v in zip(names, m)) + f"   sum {m.sum():7.0f}")
