"""RRT*-Dubins / RRT*-Reeds-Shepp: a warp per query vs a CTA per query -- same trees (bitwise), kernel ms by batch size."""
import math, sys
sys.path.insert(0, "/root/repo/robotics-path-planning_b200")
import numpy as np
from rrtk import dubins_planner as DP
which = sys.argv[1] if len(sys.argv) > 1 else "dubins"
iters = 500 if which == "dubins" else 300
obs1 = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2)]
for Q in (16, 64, 128, 256, 444, 1024):
    rng = np.random.default_rng(7)
    st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
    st[rng.integers(0, 101, (Q, iters)) <= 10] = (10.0, 10.0, 0.0)
    out, ms = {}, {}
    for mode in ("warp", "cta"):
        tm = {}
        for rep in range(2):
            kw = dict(steer="dubins") if which == "dubins" else dict(steer="rs", step_size=0.2, rs_cost=(which == "rs2"))
            res = DP.run_batch([[0.0, 0.0, 0.0]] * Q, [[10.0, 10.0, 0.0]] * Q, [obs1] * Q, 3.0 if which != "rs2" else float("inf"),
                               iters, st, timing=tm, exec_mode=mode, near_cap=512, **kw)
        out[mode], ms[mode] = res, tm["kernel_ms"]
    same = all(a["n"] == b["n"] and all(np.array_equal(a[k], b[k], equal_nan=True) for k in a if isinstance(a[k], np.ndarray))
               and a.get("goal_index") == b.get("goal_index") for a, b in zip(out["warp"], out["cta"]))
    n = np.array([r["n"] for r in out["warp"]])
    print(which, Q, "warp %.1f ms  cta %.1f ms  identical %s  nodes mean %.0f" % (ms["warp"], ms["cta"], same, n.mean()), flush=True)
