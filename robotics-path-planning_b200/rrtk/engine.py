"""Host-side driver of the batched RRT / RRT* kernel (rrtk_rrtstar_run_dev).

Everything that depends only on the scenario -- not on the tree -- is evaluated here on the host with
the same Python float operations the reference uses, then handed to the kernel as tables:
  * per obstacle  R2 = (size + robot_radius) ** 2            (rrt_04:1227)
  * per tree size r2[k] = min(ccd * sqrt(log(k) / k), expand_dis) ** 2, k = len(node_list) + 1
                                                             (rrt_04:1329-1336)
The tree itself (nearest, steer, collision, near, choose_parent, rewire, goal search) is computed on
the GPU only."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _lib

_NEAR_R2_CACHE: dict = {}


def near_r2_table(node_cap: int, connect_circle_dist: float, expand_dis) -> np.ndarray:
    """near_r2[k] for k = 0 .. node_cap + 1 (index = number of nodes + 1); [0] is unused."""
    key = (int(node_cap), float(connect_circle_dist), None if expand_dis is None else float(expand_dis))
    tab = _NEAR_R2_CACHE.get(key)
    if tab is None:
        tab = np.zeros(node_cap + 2, dtype=np.float64)
        for k in range(1, node_cap + 2):
            r = connect_circle_dist * math.sqrt(math.log(k) / k)
            if expand_dis is not None:
                r = min(r, expand_dis)
            tab[k] = r ** 2
        _NEAR_R2_CACHE[key] = tab
    return tab


def pack_obstacles(obstacle_lists, robot_radius: float, stride: int | None = None):
    """obstacle_lists: sequence (per query) of iterables of (x, y, size).
    Returns (rows [Q, stride, 4] float64 = x, y, size + rr, (size + rr) ** 2 ; n_obs [Q] int32)."""
    q = len(obstacle_lists)
    counts = [len(o) for o in obstacle_lists]
    stride = max(max(counts, default=0), 1) if stride is None else stride
    rows = np.zeros((q, stride, 4), dtype=np.float64)
    for i, obs in enumerate(obstacle_lists):
        for j, (ox, oy, size) in enumerate(obs):
            rows[i, j, 0] = ox
            rows[i, j, 1] = oy
            rows[i, j, 2] = size + robot_radius
            rows[i, j, 3] = (size + robot_radius) ** 2  # Python pow, as the reference evaluates it
    return rows, np.asarray(counts, dtype=np.int32)


@dataclass
class BatchResult:
    """Device-resident result of a batched run (torch tensors on the GPU)."""
    xy: "object"          # [Q, node_cap, 2] float64
    cost: "object"        # [Q, node_cap]    float64
    parent: "object"      # [Q, node_cap]    int32, -1 = root
    n_nodes: "object"     # [Q] int32
    iters_done: "object"  # [Q] int32
    goal_index: "object"  # [Q] int32, -1 = no path
    status: "object"      # [Q] int32 (RRTK_Q_* bits)
    trace: "object"       # [Q, max_iter, 8] int32 or None
    start_goal: "object"  # [Q, 4] float64
    workspace: "object" = None  # [Q * workspace_ints(p) + WS_TAIL_INTS] int32 scratch (children lists, frontier, obstacle cells, queue counter)

    def paths_device(self, path_cap: int | None = None):
        """generate_final_course for every query on the GPU -> (path [Q, path_cap, 2], length [Q])."""
        torch = _lib.require_cuda()
        q, cap = self.parent.shape
        path_cap = cap + 1 if path_cap is None else path_cap
        path = torch.empty((q, path_cap, 2), dtype=torch.float64, device=self.xy.device)
        plen = torch.empty((q,), dtype=torch.int32, device=self.xy.device)
        _lib.check(_lib.lib().rrtk_extract_paths_dev(
            q, cap, path_cap, self.start_goal.data_ptr(), self.xy.data_ptr(), self.parent.data_ptr(),
            self.goal_index.data_ptr(), path.data_ptr(), plen.data_ptr(),
            torch.cuda.current_stream().cuda_stream), "rrtk_extract_paths_dev")
        return path, plen

    def paths(self):
        """list (per query) of `[[x, y], ...]` goal -> start, or None (what `planning()` returns)."""
        path, plen = self.paths_device()
        path = path.cpu().numpy()
        plen = plen.cpu().numpy()
        return [None if n == 0 else path[i, :n].tolist() for i, n in enumerate(plen)]


def make_params(n_queries, max_iter, node_cap, obs_stride, expand_dis, path_resolution,
                play_area=None, search_until_max_iter=True, sampler=_lib.SAMPLER_STREAM,
                goal_sample_rate=5, min_rand=0.0, max_rand=0.0, seed=0, near_cap=256,
                rrt_only=False, near_r_max=0.0, exec_mode=None, query_base=0) -> _lib.RRTStarParams:
    p = _lib.RRTStarParams()
    p.n_queries, p.max_iter, p.node_cap = int(n_queries), int(max_iter), int(node_cap)
    p.obs_stride, p.near_cap = int(obs_stride), int(near_cap)
    p.search_until_max_iter = int(bool(search_until_max_iter))
    p.sampler, p.goal_sample_rate = int(sampler), int(goal_sample_rate)
    p.has_play_area = 0 if play_area is None else 1
    if play_area is not None:
        for i in range(4):
            p.play_area[i] = float(play_area[i])
    p.rrt_only = int(bool(rrt_only))
    p.expand_dis, p.path_resolution = float(expand_dis), float(path_resolution)
    p.min_rand, p.max_rand = float(min_rand), float(max_rand)
    p.seed = int(seed) & ((1 << 64) - 1)
    p.near_r_max = float(near_r_max)          # 0 = the table is clipped to expand_dis (rrt_04:1333-1335)
    p.exec_mode = default_exec_mode() if exec_mode is None else int(exec_mode)
    p.query_base = int(query_base)            # global index of query 0 (shards of one batch share the sampler streams)
    return p


def default_exec_mode() -> int:
    """RRTK_EXEC_*: AUTO unless the environment variable RRTK_EXEC (auto | warp | cta) says otherwise (tuning, tests)."""
    import os
    return {"auto": _lib.EXEC_AUTO, "warp": _lib.EXEC_WARP, "cta": _lib.EXEC_CTA}[os.environ.get("RRTK_EXEC", "auto").lower()]


def near_r_max_of(table: np.ndarray, expand_dis) -> float:
    """Upper bound of the near radius for rrtk_rrtstar_params.near_r_max: 0 when the table is clipped to expand_dis."""
    if expand_dis is not None or table is None or len(table) < 2:
        return 0.0
    return math.sqrt(float(np.max(table[1:]))) * (1.0 + 1e-12)


def set_obstacle_grid(p: _lib.RRTStarParams, xmin, xmax, ymin, ymax) -> None:
    """Cell grid for the per-iteration obstacle cull (struct rrtk_rrtstar_params.grid_*).  The box must contain
    every sample, start and goal; nodes outside it are handled by the full scan, so this only affects speed."""
    reach = max(p.expand_dis, p.near_r_max) + p.path_resolution
    if not (math.isfinite(xmin) and math.isfinite(xmax) and math.isfinite(ymin) and math.isfinite(ymax)) \
            or not reach > 0.0 or xmax < xmin or ymax < ymin:
        p.grid_nx = p.grid_ny = 0
        return
    pad = 1e-6 * (1.0 + abs(xmin) + abs(xmax) + abs(ymin) + abs(ymax))
    xmin, ymin, w, h = xmin - pad, ymin - pad, (xmax - xmin) + 2 * pad, (ymax - ymin) + 2 * pad
    cell = max(reach, w / 64.0, h / 64.0)
    p.grid_nx, p.grid_ny = max(1, min(64, math.ceil(w / cell))), max(1, min(64, math.ceil(h / cell)))
    p.grid_x0, p.grid_y0, p.grid_cell = xmin, ymin, cell


def workspace_ints(p: _lib.RRTStarParams) -> int:
    """RRTK_RRTSTAR_WS_INTS (include/rrtk.h)."""
    return 4 * p.node_cap + 4 * ((p.node_cap + 1) // 2) + 4 * ((17 * p.grid_nx * p.grid_ny + 3) // 4)


def run_dev(p: _lib.RRTStarParams, start_goal, obstacles, n_obs, near_r2, sample_stream=None,
            sobol_offset=None, want_trace=False, out: BatchResult | None = None) -> BatchResult:
    """Launch the planning kernel on device tensors (float64 / int32 / int64, contiguous, CUDA)."""
    torch = _lib.require_cuda()
    dev = start_goal.device
    q, cap = p.n_queries, p.node_cap
    for t, dt in ((start_goal, torch.float64), (obstacles, torch.float64), (n_obs, torch.int32)):
        if t.dtype != dt or not t.is_contiguous() or not t.is_cuda:
            raise _lib.RrtkError("run_dev: tensors must be contiguous CUDA tensors of the documented dtype")
    if out is None:
        out = BatchResult(
            xy=torch.empty((q, cap, 2), dtype=torch.float64, device=dev),
            cost=torch.empty((q, cap), dtype=torch.float64, device=dev),
            parent=torch.empty((q, cap), dtype=torch.int32, device=dev),
            n_nodes=torch.empty((q,), dtype=torch.int32, device=dev),
            iters_done=torch.empty((q,), dtype=torch.int32, device=dev),
            goal_index=torch.empty((q,), dtype=torch.int32, device=dev),
            status=torch.empty((q,), dtype=torch.int32, device=dev),
            trace=torch.zeros((q, p.max_iter, 8), dtype=torch.int32, device=dev) if want_trace else None,
            start_goal=start_goal,
            workspace=torch.empty((q * workspace_ints(p) + _lib.WS_TAIL_INTS,), dtype=torch.int32, device=dev))
    ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    rc = _lib.lib().rrtk_rrtstar_run_dev(
        C.byref(p), ptr(start_goal), ptr(obstacles), ptr(n_obs), ptr(near_r2), ptr(sample_stream),
        ptr(sobol_offset), ptr(out.xy), ptr(out.cost), ptr(out.parent), ptr(out.n_nodes),
        ptr(out.iters_done), ptr(out.goal_index), ptr(out.status), ptr(out.trace), ptr(out.workspace),
        torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "rrtk_rrtstar_run_dev")
    return out


def sample_stream_dev(p: _lib.RRTStarParams, start_goal, sobol_offset=None):
    """Materialise the in-kernel sampler: [Q, max_iter, 2] float64 on the GPU."""
    torch = _lib.require_cuda()
    out = torch.empty((p.n_queries, p.max_iter, 2), dtype=torch.float64, device=start_goal.device)
    rc = _lib.lib().rrtk_sample_stream_dev(C.byref(p), start_goal.data_ptr(),
                                           None if sobol_offset is None else sobol_offset.data_ptr(),
                                           out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "rrtk_sample_stream_dev")
    return out


def steer_collide(from_xy, to_xy, obstacle_lists, extend_length=float("inf"), path_resolution=0.5, robot_radius=0.0,
                  play_area=None, obs_set=None, device=None):
    """steer (rrt_04:1086-1115) + check_collision (:1216-1230) + check_if_outside_play_area (:1204-1214) for N edges in
    one launch (rrtk_steer_collide_dev).  from_xy / to_xy [N, 2]; obstacle_lists: one list of (x, y, r) for all edges, or
    several with `obs_set[r]` choosing the list of edge r.  Returns dict(new_xy [N, 2], dist, n_points, free, inside)."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda" if device is None else device)
    f = np.ascontiguousarray(from_xy, dtype=np.float64).reshape(-1, 2)
    t = np.ascontiguousarray(to_xy, dtype=np.float64).reshape(-1, 2)
    n = f.shape[0]
    if len(obstacle_lists) == 0 or (len(obstacle_lists[0]) == 3 and np.ndim(obstacle_lists[0][0]) == 0):
        obstacle_lists = [obstacle_lists]
    rows, counts = pack_obstacles([list(o) for o in obstacle_lists], robot_radius)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    with torch.cuda.device(dev):
        d_f, d_t, d_rows, d_cnt = up(f), up(t), up(rows if rows.size else np.zeros((1, 1, 4))), up(counts)
        d_set = up(np.asarray(obs_set, dtype=np.int32)) if obs_set is not None else None
        d_play = up(np.asarray(play_area, dtype=np.float64)) if play_area is not None else None
        new_xy = torch.empty((n, 2), dtype=torch.float64, device=dev)
        dist = torch.empty(n, dtype=torch.float64, device=dev)
        npts = torch.empty(n, dtype=torch.int32, device=dev)
        free = torch.empty(n, dtype=torch.uint8, device=dev)
        inside = torch.empty(n, dtype=torch.uint8, device=dev)
        rc = _lib.lib().rrtk_steer_collide_dev(
            n, d_f.data_ptr(), d_t.data_ptr(), float(extend_length), float(path_resolution),
            d_set.data_ptr() if d_set is not None else None, d_rows.data_ptr(), rows.shape[1], d_cnt.data_ptr(),
            d_play.data_ptr() if d_play is not None else None, new_xy.data_ptr(), dist.data_ptr(), npts.data_ptr(),
            free.data_ptr(), inside.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "rrtk_steer_collide_dev")
        return dict(new_xy=new_xy.cpu().numpy(), dist=dist.cpu().numpy(), n_points=npts.cpu().numpy(),
                    free=free.cpu().numpy().astype(bool), inside=inside.cpu().numpy().astype(bool))


def steer_points(from_xy, to_xy, extend_length=float("inf"), path_resolution=0.5, device=None):
    """path_x / path_y of steer (rrt_04:1086-1115) for N edges in one launch (rrtk_steer_points_dev).  from_xy / to_xy
    [N, 2]; extend_length a scalar or [N].  Returns (points [N, pt_cap, 2] float64, n_points [N]) as numpy arrays; row r
    holds n_points[r] points."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda" if device is None else device)
    f = np.ascontiguousarray(from_xy, dtype=np.float64).reshape(-1, 2)
    t = np.ascontiguousarray(to_xy, dtype=np.float64).reshape(-1, 2)
    n = f.shape[0]
    ext = np.broadcast_to(np.asarray(extend_length, dtype=np.float64), (n,)).copy()
    reach = np.minimum(ext, np.hypot(t[:, 0] - f[:, 0], t[:, 1] - f[:, 1]) * (1.0 + 1e-12) + 1e-12) if n else ext
    pt_cap = int(np.floor(reach.max() / path_resolution)) + 4 if n else 1
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    with torch.cuda.device(dev):
        pts = torch.empty((n, pt_cap, 2), dtype=torch.float64, device=dev)
        npts = torch.empty(n, dtype=torch.int32, device=dev)
        d_f, d_t, d_e = up(f), up(t), up(ext)
        rc = _lib.lib().rrtk_steer_points_dev(n, d_f.data_ptr(), d_t.data_ptr(), d_e.data_ptr(), 0.0, float(path_resolution),
                                              pt_cap, pts.data_ptr(), npts.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "rrtk_steer_points_dev")
        return pts.cpu().numpy(), npts.cpu().numpy()


def points_collide(point_lists, obstacle_list, robot_radius=0.0, device=None):
    """check_collision (rrt_04:1216-1230) of N point lists (each [[x, y], ...]) against one obstacle list; returns a bool
    array, True = safe (rrtk_points_collide_dev)."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda" if device is None else device)
    n = len(point_lists)
    pt_cap = max(1, max((len(p) for p in point_lists), default=1))
    pts = np.zeros((n, pt_cap, 2), dtype=np.float64)
    cnt = np.zeros(n, dtype=np.int32)
    for r, p in enumerate(point_lists):
        a = np.asarray(p, dtype=np.float64).reshape(-1, 2)
        pts[r, :len(a)] = a
        cnt[r] = len(a)
    rows, counts = pack_obstacles([list(obstacle_list)], robot_radius)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    with torch.cuda.device(dev):
        d_p, d_c, d_rows, d_cnt = up(pts), up(cnt), up(rows if rows.size else np.zeros((1, 1, 4))), up(counts)
        free = torch.empty(n, dtype=torch.uint8, device=dev)
        rc = _lib.lib().rrtk_points_collide_dev(n, d_p.data_ptr(), d_c.data_ptr(), pt_cap, None, d_rows.data_ptr(), rows.shape[1],
                                                d_cnt.data_ptr(), free.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "rrtk_points_collide_dev")
        return free.cpu().numpy().astype(bool)


def nearest_index(xy, samples, device=None):
    """get_nearest_node_index (rrt_04:1196-1202) of B samples over the nodes xy [n, 2]: first minimum of the squared
    distance, FP64 (rrtk_nearest_f64_dev).  Returns int32 [B]."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda" if device is None else device)
    a = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
    s = np.ascontiguousarray(samples, dtype=np.float64).reshape(-1, 2)
    with torch.cuda.device(dev):
        d_a, d_s = torch.from_numpy(a).to(dev), torch.from_numpy(s).to(dev)
        idx = torch.empty(len(s), dtype=torch.int32, device=dev)
        rc = _lib.lib().rrtk_nearest_f64_dev(d_a.data_ptr(), len(a), d_s.data_ptr(), len(s), idx.data_ptr(), None,
                                             torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "rrtk_nearest_f64_dev")
        return idx.cpu().numpy()


def near_indices(xy, cx, cy, r2, device=None):
    """find_near_nodes (rrt_04:1314-1338) around (cx, cy) with squared radius r2 over the nodes xy [n, 2]: the reference's
    list, `.index()` mapping included (rrtk_near_f64_dev).  Returns a Python list of ints."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda" if device is None else device)
    a = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
    n = len(a)
    with torch.cuda.device(dev):
        d_a = torch.from_numpy(a).to(dev)
        out = torch.empty(n, dtype=torch.int32, device=dev)
        d2 = torch.empty(n, dtype=torch.float64, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        rc = _lib.lib().rrtk_near_f64_dev(d_a.data_ptr(), n, float(cx), float(cy), float(r2), out.data_ptr(), d2.data_ptr(), n,
                                          cnt.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "rrtk_near_f64_dev")
        return out[:int(cnt.item())].cpu().tolist()
