"""Per-phase wall times of bench.py's e2e step (why is e2e slower than the kernel?)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-path-planning_b200"))
import numpy as np
import torch
import rrtk
from rrtk import workloads as W
cfg = W.C2
Q, iters, n_obs = 4096, 2000, 256
qids = list(range(Q))
rows = W.c2_rows(qids, n_obs)
starts = np.tile(np.array(cfg["start"]), (Q, 1)); goals = np.tile(np.array(cfg["goal"]), (Q, 1))
batch = rrtk.RRTStarBatch(starts, goals, rows, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"],
                          cfg["goal_sample_rate"], iters, None, cfg["robot_radius"], "sobol",
                          cfg["connect_circle_dist"], True, seed=0xC2,
                          sobol_offset=np.asarray(qids, dtype=np.int64) * iters)
for _ in range(2):
    batch.run()
torch.cuda.synchronize()
h_path = torch.empty((Q, 256, 2), dtype=torch.float64).pin_memory()
h_plen = torch.empty((Q,), dtype=torch.int32).pin_memory()
for k in range(14):
    t = [time.perf_counter()]
    batch.upload(); torch.cuda.synchronize(); t.append(time.perf_counter())
    r = batch.run(); torch.cuda.synchronize(); t.append(time.perf_counter())
    path, plen = r.paths_device(256); torch.cuda.synchronize(); t.append(time.perf_counter())
    h_path.copy_(path, non_blocking=True); h_plen.copy_(plen, non_blocking=True); torch.cuda.synchronize(); t.append(time.perf_counter())
    print("step", k, " ".join(f"{(b - a) * 1e3:.1f}" for a, b in zip(t, t[1:])), "ms (upload, run, paths, d2h)", flush=True)
