"""Reeds-Shepp steering (10_path_planning_00_reeds_shepp_path.py == rrt_06:1021-1437) behind the reference's function name.

`reeds_shepp_path_planning(sx, sy, syaw, gx, gy, gyaw, maxc, step_size=0.2)` returns the reference's 5-tuple
(x, y, yaw, ctypes, lengths) or five Nones; `steer_batch` evaluates N edges in one launch (device tensors out)."""
from __future__ import annotations

import numpy as np

from . import _lib

TYPE_CHARS = "LSR"


def steer_batch(from3, to3, maxc, step_size=0.2, obstacle_sets=None, obs_set=None, robot_radius=0.0, max_pts=0,
                device=None):
    """N Reeds-Shepp edges.  from3 / to3 [N, 3].  obstacle_sets: list of [(x, y, size), ...] (optional), obs_set [N] picks
    one per edge.  Returns a dict of device tensors: types [N, 5] int32, lengths [N, 5], L [N], n_paths [N], end [N, 3],
    n_pts [N], free [N] uint8 and, if max_pts > 0, pts [N, max_pts, 4] (x, y, yaw, direction)."""
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
            rows = np.zeros((len(obstacle_sets), stride, 4))
            for i, obs in enumerate(obstacle_sets):
                for j, (ox, oy, size) in enumerate(obs):
                    rows[i, j] = (ox, oy, size + robot_radius, (size + robot_radius) ** 2)
            d_obs = torch.from_numpy(rows).to(dev)
            d_cnt = torch.tensor([len(o) for o in obstacle_sets], dtype=torch.int32, device=dev)
            if obs_set is not None:
                d_set = torch.from_numpy(np.ascontiguousarray(obs_set, dtype=np.int32)).to(dev)
        out = dict(types=torch.empty((n, 5), dtype=torch.int32, device=dev),
                   lengths=torch.empty((n, 5), dtype=torch.float64, device=dev),
                   L=torch.empty((n,), dtype=torch.float64, device=dev),
                   n_paths=torch.empty((n,), dtype=torch.int32, device=dev),
                   end=torch.zeros((n, 3), dtype=torch.float64, device=dev),
                   n_pts=torch.empty((n,), dtype=torch.int32, device=dev),
                   free=torch.empty((n,), dtype=torch.uint8, device=dev),
                   pts=torch.zeros((n, max_pts, 4), dtype=torch.float64, device=dev) if max_pts > 0 else None)
        ptr = lambda x: None if x is None else x.data_ptr()  # noqa: E731
        _lib.check(_lib.lib().rrtk_reeds_shepp_steer_dev(
            n, float(maxc), float(step_size), d_f.data_ptr(), d_t.data_ptr(), ptr(d_set), ptr(d_obs), stride, ptr(d_cnt),
            out["types"].data_ptr(), out["lengths"].data_ptr(), out["L"].data_ptr(), out["n_paths"].data_ptr(),
            out["end"].data_ptr(), out["n_pts"].data_ptr(), out["free"].data_ptr(), ptr(out["pts"]), int(max_pts),
            torch.cuda.current_stream().cuda_stream), "rrtk_reeds_shepp_steer_dev")
    return out


def reeds_shepp_path_planning(sx, sy, syaw, gx, gy, gyaw, maxc, step_size=0.2):
    """Drop-in for rs00:496-515."""
    first = steer_batch([[sx, sy, syaw]], [[gx, gy, gyaw]], maxc, step_size)
    n = int(first["n_pts"][0].item())
    if n == 0:
        return None, None, None, None, None
    out = steer_batch([[sx, sy, syaw]], [[gx, gy, gyaw]], maxc, step_size, max_pts=n)
    pts = out["pts"][0].cpu().numpy()
    ty = out["types"][0].cpu().numpy()
    k = int((ty >= 0).sum())
    return (pts[:, 0].tolist(), pts[:, 1].tolist(), pts[:, 2].tolist(), [TYPE_CHARS[v] for v in ty[:k]],
            out["lengths"][0, :k].cpu().numpy().tolist())
