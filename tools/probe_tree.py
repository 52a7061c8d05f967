"""Time the single-tree Informed RRT* kernel (config 3) at a few sizes.  usage: probe_tree.py [iters ...]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-path-planning_b200"))
import numpy as np
import torch
from rrtk import informed

OBS = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]
sizes = [int(a) for a in sys.argv[1:]] or [20000, 200000]
for iters in sizes:
    rng = np.random.default_rng(9)
    free = rng.uniform(-2, 15, (iters, 2)); coin = rng.integers(0, 101, iters) <= 10; free[coin] = (6.0, 10.0)
    ball = rng.random((iters, 2))
    informed.near_table(iters + 1)
    d_free, d_ball = torch.from_numpy(free).cuda(), torch.from_numpy(ball).cuda()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run = informed.run_tree([0.0, 0.0], [6.0, 10.0], OBS, 0.5, iters, d_free, d_ball, batch=int(os.environ.get("RRTK_TREE_BATCH", "8")))
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    i = run.info
    print(f"iters {iters}: {ms:.1f} ms, {ms * 1e3 / iters:.2f} us/iter, nodes {i['n_nodes']}, c_best {i['c_best']:.6f}, "
          f"hits/iter {i['total_hits'] / iters:.1f}, goal_events {i['goal_events']}, resamples {i['resamples']}, "
          f"slow {i['slow_paths']}, reext/cuts {i['reextends']}, batches {i['cycles'][0] if int(os.environ.get('RRTK_TREE_BATCH', '8')) > 1 else '-'}, status {i['status']}, grid {i['grid']}, "
          f"cycles/batch {[round(c / max(1, i['cycles'][0])) for c in i['cycles'][1:]]} "
          f"sub/batch {[round(c / max(1, i['cycles'][0])) for c in i['cycles_max']]} "
          f"cycles/iter {[round(c / iters) for c in i['cycles']]} max {[round(c / iters) for c in i['cycles_max']]} "
          f"min {[round(-c / iters) for c in i['cycles_negmin']]}", flush=True)
