"""Launch time of the config-2 RRT* kernel under both executions (warp per query / CTA per query) for several batch
sizes, each launch timed with its own CUDA event pair; checks that both give the same trees.
    python tools/probe_exec_modes.py [Q ...]"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "robotics-path-planning_b200"))
import rrtk
from rrtk import workloads as W

cfg = W.C2
sizes = [int(a) for a in sys.argv[1:]] or [128, 512, 1024, 2368, 4096]
iters, n_obs = 2000, 256
for Q in sizes:
    rows = W.c2_rows(list(range(Q)), n_obs)
    starts = np.tile(np.array(cfg["start"]), (Q, 1)); goals = np.tile(np.array(cfg["goal"]), (Q, 1))
    sig = {}
    for mode in os.environ.get("RRTK_PROBE_MODES", "warp,cta").split(","):
        os.environ["RRTK_EXEC"] = mode
        b = rrtk.RRTStarBatch(starts, goals, rows, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"],
                              cfg["goal_sample_rate"], iters, None, cfg["robot_radius"], "sobol", cfg["connect_circle_dist"],
                              True, seed=0xC2, sobol_offset=np.arange(Q, dtype=np.int64) * iters)
        ms = []
        for k in range(5):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r = b.run(); e.record(); torch.cuda.synchronize()
            ms.append(a.elapsed_time(e))
        sig[mode] = (r.n_nodes.cpu().numpy(), r.goal_index.cpu().numpy(), r.status.cpu().numpy(),
                     float(r.cost.cpu().numpy()[np.arange(Q), np.maximum(r.goal_index.cpu().numpy(), 0)].sum()))
        ms = sorted(ms[1:])
        print(f"Q={Q:5d} {mode:4s} ms min/med/max = {ms[0]:.2f} / {ms[len(ms)//2]:.2f} / {ms[-1]:.2f}   "
              f"{Q * iters / ms[len(ms)//2] / 1e3:.1f} M it/s   mean nodes {sig[mode][0].mean():.0f}", flush=True)
        del b, r
    if len(sig) < 2:
        continue
    same = all(np.array_equal(x, y) for x, y in zip(sig["warp"][:3], sig["cta"][:3])) and sig["warp"][3] == sig["cta"][3]
    print(f"Q={Q:5d} identical results: {same}", flush=True)
