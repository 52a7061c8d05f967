"""One BIT* launch (rrt_08 semantics, the bench's scene) for profiling:  python tools/probe_bitstar_one.py [Q] [iters] [reps]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "robotics-path-planning_b200"))
import numpy as np
from rrtk import bitstar as BS
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rng = np.random.default_rng(23)
draws = rng.random((Q, 6000))
obs1 = [(5, 5, 0.5), (9, 6, 1), (7, 5, 1), (1, 5, 1), (3, 6, 1), (7, 9, 1)]
tm = {}
for rep in range(reps):
    res = BS.run_batch([[-1.0, 0.0]] * Q, [[3.0, 8.0]] * Q, [obs1] * Q, [-2, 15], iters, draws, timing=tm)
    print(Q, iters, "kernel ms %.2f" % tm["kernel_ms"], flush=True)
