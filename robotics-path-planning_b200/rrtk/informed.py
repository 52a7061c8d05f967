"""Informed RRT* (10_path_planning_01_rrt_07_informed_rrt_star.py) behind the reference's API.

`InformedRRTStar` is the class the reference calls `RRT` in rrt_07:1027-1285: same constructor, entry point
`informed_rrt_star_search(animation=...)`, returns the best path (goal -> start) or None; `node_list` holds
`Node` objects whose `parent` is an int index (rrt_07:1020-1025, :1133).  The search runs in one kernel launch.

Sampling: the reference draws from Python's `random` lazily -- `randint` + two `uniform`s while no solution
exists, two `random()`s afterwards (rrt_07:1145-1191).  Which branch runs depends on the tree, so the GPU path
pre-draws BOTH per iteration (same distributions; the draw order differs from the reference's once a solution
exists).  Parity tests inject the draws on both sides instead."""
from __future__ import annotations

import ctypes as C
import math
import random

import numpy as np

from . import _lib, sampling


class Node:
    """rrt_07:1020-1025."""

    def __init__(self, x, y):
        self.x = x
        self.y = y
        self.cost = 0.0
        self.parent = None


def rotation_to_world_frame(start, goal):
    """The matrix C of rrt_07:1054-1068 (numpy SVD, evaluated exactly like the reference); 2 x 2 block."""
    c_min = math.hypot(start[0] - goal[0], start[1] - goal[1])
    a1 = np.array([[(goal[0] - start[0]) / c_min], [(goal[1] - start[1]) / c_min], [0]])
    id1_t = np.array([1.0, 0.0, 0.0]).reshape(1, 3)
    m = a1 @ id1_t
    u, s, vh = np.linalg.svd(m, True, True)
    c = u @ np.diag([1.0, 1.0, np.linalg.det(u) * np.linalg.det(np.transpose(vh))]) @ vh
    return [float(c[0, 0]), float(c[0, 1]), float(c[1, 0]), float(c[1, 1])]


_NEAR_CACHE: dict = {}


def near_table(node_cap: int) -> np.ndarray:
    """[node_cap + 1, 2] = (r, r ** 2) with r = 50 * sqrt(log(n) / n), n = len(node_list) (rrt_07:1138-1139)."""
    tab = _NEAR_CACHE.get(node_cap)
    if tab is None:
        tab = np.zeros((node_cap + 1, 2), dtype=np.float64)
        for n in range(1, node_cap + 1):
            r = 50.0 * math.sqrt(math.log(n) / n)
            tab[n] = (r, r ** 2)
        _NEAR_CACHE[node_cap] = tab
    return tab


def run_batch(starts, goals, obstacle_lists, expand_dis, max_iter, free_samples, ball_draws, path_cap=1024,
              device=None, timing=None, exec_mode="auto"):
    """Q informed searches in one launch.  free_samples / ball_draws: [Q, max_iter, 2].
    Returns dict of numpy arrays / lists (trees trimmed to n_nodes).
    `timing`: optional dict that receives `kernel_ms` (CUDA events around the launch).
    `exec_mode`: "warp" (a warp per query), "cta" (a CTA per query: batches resident all at once), "auto"; same trees."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda" if device is None else device)
    starts = np.asarray(starts, dtype=np.float64).reshape(-1, 2)
    goals = np.asarray(goals, dtype=np.float64).reshape(-1, 2)
    q = starts.shape[0]
    cap = max_iter + 1
    stride = max(max((len(o) for o in obstacle_lists), default=0), 1)
    rows = np.zeros((q, stride, 4), dtype=np.float64)
    counts = np.zeros(q, dtype=np.int32)
    for i, obs in enumerate(obstacle_lists):
        counts[i] = len(obs)
        for j, (ox, oy, size) in enumerate(obs):
            rows[i, j] = (ox, oy, size, size ** 2)   # size ** 2 as Python evaluates it (rrt_07:1267)
    rot = np.array([rotation_to_world_frame(s, g) for s, g in zip(starts, goals)], dtype=np.float64)
    p = _lib.InformedParams()
    p.n_queries, p.max_iter, p.node_cap, p.obs_stride, p.path_cap = q, max_iter, cap, stride, path_cap
    p.expand_dis = float(expand_dis)
    p.exec_mode = {"auto": 0, "warp": 1, "cta": 2}[exec_mode]
    fs = np.asarray(free_samples, dtype=np.float64)
    bound = max(float(np.abs(starts).max()), float(np.abs(goals).max()), float(np.abs(fs).max()) if fs.size else 0.0,
                float(np.abs(rows[:, :, :3]).max()) if rows.size else 0.0, 1.0)
    p.coord_bound = 2.0 * bound + 4.0 * float(expand_dis)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    with torch.cuda.device(dev):
        d_sg, d_rot, d_obs, d_cnt = t(np.hstack([starts, goals])), t(rot), t(rows), t(counts)
        d_near = t(near_table(cap))
        d_free = t(np.asarray(free_samples, dtype=np.float64).reshape(q, max_iter, 2))
        d_ball = t(np.asarray(ball_draws, dtype=np.float64).reshape(q, max_iter, 2))
        xy = torch.empty((q, cap, 2), dtype=torch.float64, device=dev)
        cost = torch.empty((q, cap), dtype=torch.float64, device=dev)
        parent = torch.empty((q, cap), dtype=torch.int32, device=dev)
        n_nodes = torch.empty((q,), dtype=torch.int32, device=dev)
        path = torch.zeros((q, path_cap, 2), dtype=torch.float64, device=dev)
        plen = torch.empty((q,), dtype=torch.int32, device=dev)
        c_best = torch.empty((q,), dtype=torch.float64, device=dev)
        status = torch.empty((q,), dtype=torch.int32, device=dev)
        ws_idx = torch.empty((q * cap + _lib.WS_TAIL_INTS,), dtype=torch.int32, device=dev)
        ws_d = torch.empty((q, cap), dtype=torch.float64, device=dev)
        if timing is not None:
            ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev[0].record()
        rc = _lib.lib().rrtk_informed_run_dev(
            C.byref(p), d_sg.data_ptr(), d_rot.data_ptr(), d_obs.data_ptr(), d_cnt.data_ptr(), d_near.data_ptr(),
            d_free.data_ptr(), d_ball.data_ptr(), xy.data_ptr(), cost.data_ptr(), parent.data_ptr(),
            n_nodes.data_ptr(), path.data_ptr(), plen.data_ptr(), c_best.data_ptr(), status.data_ptr(),
            ws_idx.data_ptr(), ws_d.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "rrtk_informed_run_dev")
        if timing is not None:
            ev[1].record()
            ev[1].synchronize()
            timing["kernel_ms"] = ev[0].elapsed_time(ev[1])
        n = n_nodes.cpu().numpy()
        pl = plen.cpu().numpy()
        h_xy, h_cost, h_par, h_path = xy.cpu().numpy(), cost.cpu().numpy(), parent.cpu().numpy(), path.cpu().numpy()
        st = status.cpu().numpy()
        cb = c_best.cpu().numpy()
    out = []
    for i in range(q):
        k = int(n[i])
        out.append(dict(x=h_xy[i, :k, 0].copy(), y=h_xy[i, :k, 1].copy(), cost=h_cost[i, :k].copy(),
                        parent=h_par[i, :k].copy(), n=k, c_best=float(cb[i]), status=int(st[i]),
                        path=None if pl[i] == 0 else h_path[i, :min(int(pl[i]), path_cap)].tolist()))
    return out


class TreeRun:
    """Device-resident result of `run_tree` (one large Informed RRT* tree)."""

    def __init__(self, xy, cost, parent, path, res, max_iter):
        self.xy, self.cost, self.parent, self.path_dev, self._res, self.max_iter = xy, cost, parent, path, res, max_iter
        self._info = None

    @property
    def info(self) -> dict:
        """The kernel's result struct (synchronises)."""
        if self._info is None:
            raw = bytes(self._res.cpu().numpy().tobytes())
            r = _lib.InformedTreeResult.from_buffer_copy(raw)
            self._info = {k: getattr(r, k) for k, _ in _lib.InformedTreeResult._fields_}
            for k in ("cycles", "cycles_max", "cycles_negmin"):
                self._info[k] = list(self._info[k])
        return self._info

    def arrays(self) -> dict:
        """Host copies trimmed to n_nodes, same keys as `run_batch` entries."""
        i = self.info
        k = i["n_nodes"]
        xy = self.xy[:k].cpu().numpy()
        pl = i["path_len"]
        return dict(x=xy[:, 0].copy(), y=xy[:, 1].copy(), cost=self.cost[:k].cpu().numpy(),
                    parent=self.parent[:k].cpu().numpy(), n=k, c_best=i["c_best"], status=i["status"],
                    path=None if pl == 0 else self.path_dev[:min(pl, self.path_dev.shape[0])].cpu().numpy().tolist())


def run_tree(start, goal, obstacle_list, expand_dis, max_iter, free_samples, ball_draws, node_cap=None,
             path_cap=4096, grid=0, device=None, batch=8) -> TreeRun:
    """ONE informed search with the whole GPU on it (BASELINE config 3; rrtk_informed_tree_run_dev).
    free_samples / ball_draws: [max_iter, 2] (numpy or CUDA tensors).  Enqueues and returns; results stay on the GPU."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda" if device is None else device)
    cap = int(max_iter) + 1 if node_cap is None else int(node_cap)
    rows = np.array([[ox, oy, size, size ** 2] for ox, oy, size in obstacle_list], dtype=np.float64).reshape(-1, 4)
    p = _lib.InformedTreeParams()
    p.max_iter, p.node_cap, p.n_obs, p.path_cap, p.grid = int(max_iter), cap, rows.shape[0], int(path_cap), int(grid)
    p.batch = int(batch)   # samples per pass (identical results for every value, see include/rrtk.h)
    p.expand_dis = float(expand_dis)
    sg = [float(start[0]), float(start[1]), float(goal[0]), float(goal[1])]
    rot = rotation_to_world_frame(start, goal)
    for i in range(4):
        p.start_goal[i] = sg[i]
        p.rot[i] = rot[i]

    def t(a):
        if isinstance(a, torch.Tensor):
            return a.to(dev, torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    with torch.cuda.device(dev):
        d_free, d_ball = t(free_samples).reshape(-1, 2), t(ball_draws).reshape(-1, 2)
        if d_free.shape[0] < max_iter or d_ball.shape[0] < max_iter:
            raise _lib.RrtkError("run_tree: need max_iter rows of free_samples and ball_draws")
        bound = max([abs(v) for v in sg] + [float(d_free.abs().max().item()) if max_iter else 0.0] +
                    [abs(float(v)) for v in rows[:, :3].ravel()] + [1.0])
        p.coord_bound = 2.0 * bound + 4.0 * float(expand_dis)
        d_obs = t(rows)
        d_near = t(near_table(cap))
        xy = torch.empty((cap, 2), dtype=torch.float64, device=dev)
        cost = torch.empty((cap,), dtype=torch.float64, device=dev)
        parent = torch.empty((cap,), dtype=torch.int32, device=dev)
        path = torch.zeros((path_cap, 2), dtype=torch.float64, device=dev)
        res = torch.zeros((C.sizeof(_lib.InformedTreeResult),), dtype=torch.uint8, device=dev)
        nbytes = _lib.lib().rrtk_informed_tree_workspace_bytes(cap, int(grid))
        if nbytes < 0:
            _lib.check(int(nbytes), "rrtk_informed_tree_workspace_bytes")
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        rc = _lib.lib().rrtk_informed_tree_run_dev(
            C.byref(p), d_obs.data_ptr() if rows.shape[0] else None, d_near.data_ptr(), d_free.data_ptr(),
            d_ball.data_ptr(), xy.data_ptr(), cost.data_ptr(), parent.data_ptr(), path.data_ptr(), res.data_ptr(),
            ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "rrtk_informed_tree_run_dev")
        run = TreeRun(xy, cost, parent, path, res, max_iter)
        run._keep = (d_free, d_ball, d_obs, d_near, ws)   # inputs stay alive until the kernel has run
    return run


class InformedRRTStar:
    """rrt_07's `RRT` (Informed RRT*), same constructor keywords and defaults (rrt_07:1029-1042)."""

    def __init__(self, start, goal, obstacle_list, rand_area, expand_dis=0.5, goal_sample_rate=10,
                 max_iter=200, sobol_sampler=False):
        self.start = Node(start[0], start[1])
        self.goal = Node(goal[0], goal[1])
        self.min_rand = rand_area[0]
        self.max_rand = rand_area[1]
        self.expand_dis = expand_dis
        self.goal_sample_rate = goal_sample_rate
        self.max_iter = max_iter
        self.obstacle_list = obstacle_list
        self.node_list = None
        self.sobol_sampler = sobol_sampler
        self.sobol_inter_ = 0
        self.c_best = float("inf")
        self._arrays = None

    def informed_rrt_star_search(self, animation=True, free_samples=None, ball_draws=None):
        n = int(self.max_iter)
        goal = (float(self.goal.x), float(self.goal.y))
        if free_samples is None:
            free_samples, _, nxt = sampling.draw_stream(n, goal, self.min_rand, self.max_rand,
                                                        self.goal_sample_rate, self.sobol_sampler,
                                                        self.sobol_inter_, random)
            self.sobol_inter_ = nxt
        if ball_draws is None:
            ball_draws = np.array([[random.random(), random.random()] for _ in range(n)], dtype=np.float64)
        r = run_batch([[self.start.x, self.start.y]], [goal], [list(self.obstacle_list)], self.expand_dis, n,
                      np.asarray(free_samples, dtype=np.float64)[:n], np.asarray(ball_draws, dtype=np.float64)[:n])[0]
        if r["status"] & _lib.Q_PATH_OVERFLOW:
            raise _lib.RrtkError("best path longer than path_cap")
        self._arrays = r
        self.c_best = r["c_best"]
        nodes = [Node(float(x), float(y)) for x, y in zip(r["x"], r["y"])]
        for i, nd in enumerate(nodes):
            nd.cost = float(r["cost"][i])
            nd.parent = None if r["parent"][i] < 0 else int(r["parent"][i])
        self.node_list = nodes
        return r["path"]

    def tree_arrays(self):
        return self._arrays

    @staticmethod
    def get_path_len(path):
        """rrt_07:1193-1203."""
        path_len = 0
        for i in range(1, len(path)):
            path_len += math.hypot(path[i][0] - path[i - 1][0], path[i][1] - path[i - 1][1])
        return path_len
