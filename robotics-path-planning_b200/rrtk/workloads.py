"""Synthetic workloads of BASELINE.json (seeded, shared by the GPU path, the oracle and the bench)."""
from __future__ import annotations

import numpy as np

# BASELINE config 2: 4096 queries x 256 random circles x 2000 iterations
C2 = dict(n_queries=4096, n_obs=256, max_iter=2000, rand_area=(-2.0, 15.0), start=(0.0, 0.0),
          goal=(13.0, 13.0), expand_dis=1.0, path_resolution=0.1, goal_sample_rate=5,
          connect_circle_dist=50.0, robot_radius=0.0, play_area=None, rmin=0.1, rmax=0.4)


def c2_obstacles(query_id: int, n_obs: int = 256, cfg=C2) -> np.ndarray:
    """[n_obs, 3] circles (x, y, radius) of query `query_id`: centres ~ U[rand_area]^2, radius ~
    U[rmin, rmax], rejected if within radius + 0.5 of the start or the goal; numpy default_rng(1234 + q)."""
    rng = np.random.default_rng(1234 + int(query_id))
    lo, hi = cfg["rand_area"]
    out = np.empty((0, 3))
    while out.shape[0] < n_obs:
        m = 2 * n_obs
        xy = rng.uniform(lo, hi, (m, 2))
        r = rng.uniform(cfg["rmin"], cfg["rmax"], m)
        ok = np.ones(m, dtype=bool)
        for cx, cy in (cfg["start"], cfg["goal"]):
            ok &= np.hypot(xy[:, 0] - cx, xy[:, 1] - cy) > r + 0.5
        out = np.vstack([out, np.column_stack([xy[ok], r[ok]])])
    return np.ascontiguousarray(out[:n_obs])


def c2_rows(query_ids, n_obs: int = 256, cfg=C2) -> np.ndarray:
    """Prepared obstacle rows [Q, n_obs, 4] = x, y, size + rr, (size + rr) ** 2 (Python pow, rrt_04:1227)."""
    rr = cfg["robot_radius"]
    rows = np.empty((len(query_ids), n_obs, 4), dtype=np.float64)
    for k, q in enumerate(query_ids):
        o = c2_obstacles(q, n_obs, cfg)
        rows[k, :, 0:2] = o[:, 0:2]
        big = (o[:, 2] + rr).tolist()
        rows[k, :, 2] = big
        rows[k, :, 3] = [b ** 2 for b in big]
    return rows
