"""Dubins steering (10_path_planning_00_dubins_path.py == rrt_05:1021-1278) on the GPU.

`plan_dubins_path(s_x, s_y, s_yaw, g_x, g_y, g_yaw, curvature, step_size=0.1)` keeps the reference's
signature and return value `(x_list, y_list, yaw_list, modes, lengths)`; `steer_batch` evaluates many edges
(with their sampled collision test) in one launch -- the primitive RRT*-Dubins' steer / choose_parent /
rewire are made of (rrt_05:1458-1479, :1648-1775).  No CPU fallback."""
from __future__ import annotations

import numpy as np

from . import _lib

MODES = ("LSL", "RSR", "LSR", "RSL", "RLR", "LRL")   # _PATH_TYPE_MAP order (rrt_05:1797)


def steer_batch(from3, to3, curvature=1.0, step_size=0.1, obstacle_sets=None, obs_set=None, robot_radius=0.0,
                max_pts=0, device=None):
    """N Dubins edges from `from3[i]` to `to3[i]` ([N, 3] = x, y, yaw).
    obstacle_sets: list (per set) of [(x, y, size), ...]; obs_set [N]: set index per edge (default 0).
    Returns dict of numpy arrays: mode [N], lengths [N, 3], end [N, 3], n_pts [N], free [N] bool, cost [N]
    (= sum |lengths|, what steer adds to the node cost, rrt_05:1476) and pts [N, max_pts, 3] if max_pts > 0."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda" if device is None else device)
    f = np.ascontiguousarray(np.asarray(from3, dtype=np.float64).reshape(-1, 3))
    t = np.ascontiguousarray(np.asarray(to3, dtype=np.float64).reshape(-1, 3))
    n = f.shape[0]
    with torch.cuda.device(dev):
        d_f, d_t = torch.from_numpy(f).to(dev), torch.from_numpy(t).to(dev)
        d_obs = d_cnt = d_set = None
        stride = 0
        if obstacle_sets:
            stride = max(max(len(o) for o in obstacle_sets), 1)
            rows = np.zeros((len(obstacle_sets), stride, 4), dtype=np.float64)
            for si, obs in enumerate(obstacle_sets):
                for j, (ox, oy, size) in enumerate(obs):
                    rows[si, j] = (ox, oy, size + robot_radius, (size + robot_radius) ** 2)  # rrt_05:1636
            d_obs = torch.from_numpy(rows).to(dev)
            d_cnt = torch.tensor([len(o) for o in obstacle_sets], dtype=torch.int32, device=dev)
            if obs_set is not None:
                d_set = torch.from_numpy(np.ascontiguousarray(obs_set, dtype=np.int32)).to(dev)
        mode = torch.empty(n, dtype=torch.int32, device=dev)
        lengths = torch.empty((n, 3), dtype=torch.float64, device=dev)
        end = torch.empty((n, 3), dtype=torch.float64, device=dev)
        n_pts = torch.empty(n, dtype=torch.int32, device=dev)
        free = torch.empty(n, dtype=torch.uint8, device=dev)
        pts = torch.zeros((n, max_pts, 3), dtype=torch.float64, device=dev) if max_pts > 0 else None
        ptr = lambda x: None if x is None else x.data_ptr()  # noqa: E731
        _lib.check(_lib.lib().rrtk_dubins_steer_dev(
            n, float(curvature), float(step_size), ptr(d_f), ptr(d_t), ptr(d_set), ptr(d_obs), stride, ptr(d_cnt),
            ptr(mode), ptr(lengths), ptr(end), ptr(n_pts), ptr(free), ptr(pts), max_pts,
            torch.cuda.current_stream().cuda_stream), "rrtk_dubins_steer_dev")
        out = dict(mode=mode.cpu().numpy(), lengths=lengths.cpu().numpy(), end=end.cpu().numpy(),
                   n_pts=n_pts.cpu().numpy(), free=free.cpu().numpy().astype(bool))
        out["cost"] = np.abs(out["lengths"]).sum(axis=1)
        if pts is not None:
            out["pts"] = pts.cpu().numpy()
    return out


def plan_dubins_path(s_x, s_y, s_yaw, g_x, g_y, g_yaw, curvature, step_size=0.1, selected_types=None):
    """Drop-in for rrt_05:1021-1109 / dub00 (all six words; `selected_types` is not supported)."""
    if selected_types is not None:
        raise NotImplementedError("selected_types: only the default (all six words) is implemented")
    r = steer_batch([[s_x, s_y, s_yaw]], [[g_x, g_y, g_yaw]], curvature, step_size, max_pts=1)
    n = int(r["n_pts"][0])
    r = steer_batch([[s_x, s_y, s_yaw]], [[g_x, g_y, g_yaw]], curvature, step_size, max_pts=max(n, 1))
    p = r["pts"][0, :n]
    mode = list(MODES[int(r["mode"][0])])
    return p[:, 0].copy(), p[:, 1].copy(), p[:, 2].copy(), mode, list(r["lengths"][0])
