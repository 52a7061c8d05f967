import sys, time; sys.path.insert(0, "/root/repo/robotics-path-planning_b200")
import numpy as np, torch
from rrtk import bitstar as BS
obs1 = [(5, 5, 0.5), (9, 6, 1), (7, 5, 1), (1, 5, 1), (3, 6, 1), (7, 9, 1)]
for Q, iters in ((64, 200), (1024, 200), (4096, 80)):
    draws = np.random.default_rng(23).random((Q, 6000))
    tm = {}
    for rep in range(2):
        res = BS.run_batch([[-1.0, 0.0]] * Q, [[3.0, 8.0]] * Q, [obs1] * Q, [-2, 15], iters, draws, timing=tm)
    ok = [r for r in res if r["status"] == 0]
    print(Q, iters, "kernel ms %.1f" % tm["kernel_ms"], "ok", len(ok), "solved", sum(r["path_len"] > 0 for r in ok),
          "statuses", sorted(set(r["status"] for r in res)), "mean batches", np.mean([r["batches"] for r in ok]),
          "max n_eq", max(r["n_eq"] for r in res), "max samples", max(r["n_sample_slots"] for r in res))
