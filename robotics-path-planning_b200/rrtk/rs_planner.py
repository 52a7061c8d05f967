"""RRT*-Reeds-Shepp (10_path_planning_01_rrt_06_rrt_star_reeds_shepp_path.py) behind the reference's API.

`RRTStarReedsShepp` is the class rrt_06 calls `RRT` (rrt_06:1444-1913): same constructor keywords,
`planning(animation=True, search_until_max_iter=True)` returns the sampled final course [[x, y, yaw], ...] (goal -> start)
or None.  Quirks kept: the sampler has no goal bias (:1658-1666), costs in choose_parent / rewire / propagate are
Euclidean (the later calc_new_cost wins), `try_goal_path` appends a goal-connecting node after every accepted node
(:1572-1582), goal index 0 counts as "not found".  No CPU fallback."""
from __future__ import annotations

import random
from math import pi

import numpy as np

from . import _lib, reeds_shepp
from .dubins_planner import run_batch as _run_batch


class Node:
    """rrt_06:1445-1458; path_x / path_y / path_yaw are regenerated on first access from the edge's pose pair."""

    def __init__(self, x, y, yaw):
        self.x, self.y, self.yaw = x, y, yaw
        self.parent = None
        self.cost = 0.0
        self._edge = None
        self._course = None

    def _pts(self):
        if self._course is None:
            if self._edge is None:
                self._course = ([], [], [])
            else:
                f, t, maxc, step = self._edge
                x, y, yaw, _, _ = reeds_shepp.reeds_shepp_path_planning(f[0], f[1], f[2], t[0], t[1], t[2], maxc, step)
                self._course = (list(x), list(y), list(yaw))
        return self._course

    path_x = property(lambda self: self._pts()[0])
    path_y = property(lambda self: self._pts()[1])
    path_yaw = property(lambda self: self._pts()[2])


def run_batch(starts, goals, obstacle_lists, expand_dis, max_iter, streams, robot_radius=0.0, connect_circle_dist=50.0,
              curvature=1.0, goal_yaw_th=np.deg2rad(1.0), goal_xy_th=0.5, search_until_max_iter=True, step_size=0.2,
              near_cap=256, device=None, timing=None, rs_cost=False, exec_mode="auto"):
    """Q RRT*-Reeds-Shepp queries in one launch (rrtk_rrtstar_rs_run_dev).  starts / goals [Q, 3]; streams [Q, max_iter, 3].
    rs_cost=True: rrt_10's variant (Reeds-Shepp-length costs; pass expand_dis=inf for its unclipped near radius)."""
    return _run_batch(starts, goals, obstacle_lists, expand_dis, max_iter, streams, robot_radius, connect_circle_dist,
                      curvature, goal_yaw_th, goal_xy_th, search_until_max_iter, near_cap, device, steer="rs",
                      step_size=step_size, timing=timing, rs_cost=rs_cost, exec_mode=exec_mode)


def final_course(tree, start, goal, curvature, step_size):
    """generate_final_course (rrt_06:1643-1651): reversed (x, y, yaw) samples of every edge from the goal node to the root,
    regenerated on the GPU from the stored pose pairs."""
    gi = tree["goal_index"]
    if gi < 0:
        return None
    chain = []
    i = gi
    while tree["parent"][i] >= 0:
        chain.append(i)
        i = int(tree["parent"][i])
    path = [[float(goal[0]), float(goal[1]), float(goal[2])]]
    if chain:
        r = reeds_shepp.steer_batch(tree["edge_from"][chain], tree["edge_to"][chain], curvature, step_size)
        mp = int(r["n_pts"].max().item())
        r = reeds_shepp.steer_batch(tree["edge_from"][chain], tree["edge_to"][chain], curvature, step_size, max_pts=mp)
        pts, npts = r["pts"].cpu().numpy(), r["n_pts"].cpu().numpy()
        for j in range(len(chain)):
            path.extend(pts[j, :int(npts[j])][::-1, 0:3].tolist())
    path.append([float(start[0]), float(start[1]), float(start[2])])
    return path


class RRTStarReedsShepp:
    """rrt_06's `RRT`: RRT* with Reeds-Shepp steering, same constructor keywords and defaults (rrt_06:1467-1525)."""

    Node = Node

    def __init__(self, start, goal, obstacle_list, rand_area, expand_dis=3.0, path_resolution=0.5, goal_sample_rate=5,
                 max_iter=500, play_area=None, robot_radius=0.0, sobol_sampler=True, connect_circle_dist=50.0,
                 search_until_max_iter=False, curvature=1.0, goal_yaw_th=np.deg2rad(1.0), goal_xy_th=0.5, step_size=0.2,
                 near_cap=256):
        self.start = Node(start[0], start[1], start[2])
        self.end = Node(goal[0], goal[1], goal[2])
        self.min_rand, self.max_rand = rand_area[0], rand_area[1]
        self.play_area = play_area           # accepted; rrt_06 never checks it
        self.expand_dis, self.path_resolution = expand_dis, path_resolution
        self.goal_sample_rate, self.max_iter = goal_sample_rate, max_iter
        self.obstacle_list = obstacle_list
        self.node_list = []
        self.robot_radius = robot_radius
        self.sobol_sampler, self.sobol_inter_ = sobol_sampler, 0
        self.connect_circle_dist = connect_circle_dist
        self.search_until_max_iter = search_until_max_iter
        self.curvature, self.goal_yaw_th, self.goal_xy_th, self.step_size = curvature, goal_yaw_th, goal_xy_th, step_size
        self.near_cap = near_cap
        self._tree = None

    def set_random_seed(self, seed):
        random.seed(seed)

    def get_random_node(self):
        """rrt_06:1658-1666 (no goal bias)."""
        return (random.uniform(self.min_rand, self.max_rand), random.uniform(self.min_rand, self.max_rand),
                random.uniform(-pi, pi))

    def planning(self, animation=True, search_until_max_iter=True, sample_stream=None):
        n = int(self.max_iter)
        if sample_stream is None:
            sample_stream = np.array([self.get_random_node() for _ in range(n)], dtype=np.float64)
        stream = np.asarray(sample_stream, dtype=np.float64).reshape(-1, 3)[:n]
        start = (self.start.x, self.start.y, self.start.yaw)
        goal = (self.end.x, self.end.y, self.end.yaw)
        t = run_batch([start], [goal], [list(self.obstacle_list)], self.expand_dis, n, stream[None], self.robot_radius,
                      self.connect_circle_dist, self.curvature, self.goal_yaw_th, self.goal_xy_th, search_until_max_iter,
                      self.step_size, self.near_cap)[0]
        if t["status"] & _lib.Q_NEAR_OVERFLOW:
            raise _lib.RrtkError("near list overflow: raise near_cap")
        self._tree = t
        nodes = [Node(float(x), float(y), float(w)) for x, y, w in zip(t["x"], t["y"], t["yaw"])]
        for i, nd in enumerate(nodes):
            nd.cost = float(t["cost"][i])
            if t["parent"][i] >= 0:
                nd.parent = nodes[t["parent"][i]]
                nd._edge = (t["edge_from"][i], t["edge_to"][i], self.curvature, self.step_size)
        self.node_list = nodes
        return final_course(t, start, goal, self.curvature, self.step_size)

    def tree_arrays(self):
        return self._tree
