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
              device=None):
    """Q informed searches in one launch.  free_samples / ball_draws: [Q, max_iter, 2].
    Returns dict of numpy arrays / lists (trees trimmed to n_nodes)."""
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
        ws_idx = torch.empty((q, cap), dtype=torch.int32, device=dev)
        ws_d = torch.empty((q, cap), dtype=torch.float64, device=dev)
        rc = _lib.lib().rrtk_informed_run_dev(
            C.byref(p), d_sg.data_ptr(), d_rot.data_ptr(), d_obs.data_ptr(), d_cnt.data_ptr(), d_near.data_ptr(),
            d_free.data_ptr(), d_ball.data_ptr(), xy.data_ptr(), cost.data_ptr(), parent.data_ptr(),
            n_nodes.data_ptr(), path.data_ptr(), plen.data_ptr(), c_best.data_ptr(), status.data_ptr(),
            ws_idx.data_ptr(), ws_d.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "rrtk_informed_run_dev")
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
