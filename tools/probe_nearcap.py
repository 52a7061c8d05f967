"""C2 kernel time for a few near_cap values (shared memory per warp vs L1 size)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-path-planning_b200"))
import numpy as np, torch, rrtk
from rrtk import workloads as W
cfg = W.C2
Q, iters, n_obs = 4096, 2000, 256
qids = list(range(Q)); rows = W.c2_rows(qids, n_obs)
starts = np.tile(np.array(cfg["start"]), (Q, 1)); goals = np.tile(np.array(cfg["goal"]), (Q, 1))
for nc in (256, 128, 96, 64):
    b = rrtk.RRTStarBatch(starts, goals, rows, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"], cfg["goal_sample_rate"],
                          iters, None, cfg["robot_radius"], "sobol", cfg["connect_circle_dist"], True, seed=0xC2, near_cap=nc,
                          sobol_offset=np.asarray(qids, dtype=np.int64) * iters)
    for _ in range(2): b.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.run(); b.run(); e1.record(); torch.cuda.synchronize()
    st = b.result.status.cpu().numpy()
    print(f"near_cap {nc}: {e0.elapsed_time(e1) / 2:.1f} ms, overflowed queries {(st != 0).sum()}", flush=True)
