"""Per-role cycles of the CTA-per-query RRT* kernel's software pipeline (apply phase of iteration i beside stage A of i + 1).
Needs the statistics build:  make -C robotics-path-planning_b200/csrc stats
    RRTK_LIB=robotics-path-planning_b200/rrtk/librrtk_stats.so python tools/probe_cta_stats.py"""
import os, sys
sys.path.insert(0, "/root/repo/robotics-path-planning_b200")
os.environ["RRTK_EXEC"] = "cta"
import numpy as np, torch, rrtk
from rrtk import workloads as W
cfg = W.C2; Q = 128; iters = 2000; n_obs = 256
rows = W.c2_rows(list(range(Q)), n_obs)
starts = np.tile(np.array(cfg["start"]), (Q, 1)); goals = np.tile(np.array(cfg["goal"]), (Q, 1))
b = rrtk.RRTStarBatch(starts, goals, rows, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"], cfg["goal_sample_rate"], iters, None,
                      cfg["robot_radius"], "sobol", cfg["connect_circle_dist"], True, seed=0xC2, sobol_offset=np.arange(Q, dtype=np.int64) * iters)
r = b.run(want_trace=True)
torch.cuda.synchronize()
t = r.trace.cpu().numpy().reshape(-1)[:32 * Q].reshape(Q, 32).astype(np.float64)
print("pending, overlapped, big-list, moves (mean per query):", t[:, :4].mean(axis=0), "nodes", r.n_nodes.float().mean().item())
names = ["role work (ov)", "B2 wait (ov)", "scan (ov)", "choose_parent", "rewire edges", "ov rounds"]
for w, role in enumerate(["apply", "sample+cull", "scan+first edge", "scan+rank"]):
    c = t[:, 4 + 6 * w: 10 + 6 * w].mean(axis=0) * 16
    rounds = max(c[5], 1)
    print(f"warp {w} ({role:16s}): " + "  ".join(f"{nm} {v / rounds:8.0f}" for nm, v in zip(names[:5], c[:5])) + f"   cycles per overlapped round; rounds {rounds:.0f}")
print("mean round cycles [ov+accepted, ov+rejected, plain+accepted, plain+rejected]:", t[:, 28:32].mean(axis=0))
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); b.run(want_trace=True); e.record(); torch.cuda.synchronize()
print("launch ms", a.elapsed_time(e), " -> cycles per iteration", a.elapsed_time(e) * 1e-3 * 1.965e9 / iters)
