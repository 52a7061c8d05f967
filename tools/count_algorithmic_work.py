"""Counts the reference-equivalent FP64 work per tree-iteration of a config-2 style workload with the CPU oracle
(TEST INFRASTRUCTURE; the result is the constant ALGORITHMIC_FLOP_PER_ITER in bench.py).

    python tools/count_algorithmic_work.py [iters] [n_obs] [n_sample]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "robotics-path-planning_b200"))
import oracle as O  # noqa: E402
from rrtk import sampling, workloads as W  # noqa: E402


def algorithmic_work(iters, n_obs, n_sample=4):
    """Reference-equivalent FP64 work per tree-iteration, counted on sample queries by the C oracle:
    every node-distance evaluation of nearest/near (5 flop) and every point-circle test the
    reference's check_collision performs (5 flop)."""
    cfg = W.C2
    tot_pairs = tot_scan = 0
    for w in range(n_sample):
        coins = sampling.kernel_coins(0xC2, w, iters, 5)
        pts = O.sobol_fill(2, w * iters, int((~coins).sum()))
        stream = np.empty((iters, 2))
        stream[~coins] = -2.0 + pts * 17.0
        stream[coins] = (13.0, 13.0)
        p, ob = O.make_params(cfg["start"], cfg["goal"], W.c2_obstacles(w, n_obs).tolist(), cfg["expand_dis"],
                              cfg["path_resolution"], iters, None, 0.0, cfg["connect_circle_dist"], True)
        r = O.rrtstar_run(p, ob, stream, want_trace=True)
        tr = r["trace"]
        tot_pairs += r["work_pairs"]
        tot_scan += r["work_scan"]
    per_iter = 5.0 * (tot_pairs + tot_scan) / (n_sample * iters)
    return dict(flop_per_iter=per_iter,
                note=f"5 flop x (point-circle tests + node-distance evaluations) of the brute-force reference "
                     f"algorithm, counted by the oracle over {n_sample} sample queries: "
                     f"{tot_pairs / (n_sample * iters):.0f} tests + {tot_scan / (n_sample * iters):.0f} distances "
                     f"per iteration")



if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    print(algorithmic_work(*(a + [2000, 256, 4][len(a):])))
