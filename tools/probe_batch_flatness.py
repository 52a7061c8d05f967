"""Kernel time against batch size for the secondary batch kernels: a time that does not move with the batch is one
item's serial chain (that is how the quadratic `.index()` look-up of informed_kernel was found)."""
import sys
sys.path.insert(0, "/root/repo/robotics-path-planning_b200")
import numpy as np, torch
from rrtk import bitstar as BS, smoothing as SM

rng = np.random.default_rng(23)
obs1 = [(5, 5, 0.5), (9, 6, 1), (7, 5, 1), (1, 5, 1), (3, 6, 1), (7, 9, 1)]
for Q in (32, 256, 1024):
    draws = rng.random((Q, 6000))
    tm = {}
    for rep in range(2):
        res = BS.run_batch([[-1.0, 0.0]] * Q, [[3.0, 8.0]] * Q, [obs1] * Q, [-2, 15], 200, draws, timing=tm)
    print("bitstar Q=%d x 200: %.1f ms" % (Q, tm["kernel_ms"]), flush=True)
