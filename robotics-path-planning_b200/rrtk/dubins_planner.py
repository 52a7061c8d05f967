"""RRT*-Dubins (10_path_planning_01_rrt_05_rrt_star_dubins_path.py) behind the reference's API.

`RRTStarDubins` is the class rrt_05 calls `RRT` (rrt_05:1335-1779): same constructor keywords,
`planning(animation=True, search_until_max_iter=True)` returns the sampled final course (goal -> start) or
None.  The reference's quirks are kept (see csrc/rrtk_rrtstar_dubins.cu); in particular `planning` always uses
the uniform sampler whatever `sobol_sampler` says (rrt_05:1426).  No CPU fallback."""
from __future__ import annotations

import ctypes as C
import random
from math import pi

import numpy as np

import math

from . import _lib, dubins, engine
from .planners import RRT as _RRT2D, RRTStar as _RRTStar2D


class Node:
    """rrt_05:1336-1349.  path_x / path_y / path_yaw are regenerated on first access from the pose pair
    the edge was planned between."""

    def __init__(self, x, y, yaw):
        self.x, self.y, self.yaw = x, y, yaw
        self.parent = None
        self.cost = 0.0
        self._edge = None      # (from pose, to pose, curvature)
        self._course = None

    def _pts(self):
        if self._course is None:
            if self._edge is None:
                self._course = ([], [], [])
            else:
                f, t, kappa = self._edge
                x, y, yaw, _, _ = dubins.plan_dubins_path(f[0], f[1], f[2], t[0], t[1], t[2], kappa)
                self._course = (list(x), list(y), list(yaw))
        return self._course

    path_x = property(lambda self: self._pts()[0])
    path_y = property(lambda self: self._pts()[1])
    path_yaw = property(lambda self: self._pts()[2])


def run_batch(starts, goals, obstacle_lists, expand_dis, max_iter, streams, robot_radius=0.0,
              connect_circle_dist=50.0, curvature=1.0, goal_yaw_th=np.deg2rad(1.0), goal_xy_th=0.5,
              search_until_max_iter=True, near_cap=256, device=None, steer="dubins", step_size=0.1, timing=None,
              rs_cost=False, exec_mode="auto"):
    """Q RRT*-Dubins (steer="dubins") or RRT*-Reeds-Shepp (steer="rs", rrt_06) queries in one launch.
    starts/goals [Q, 3]; streams [Q, max_iter, 3].  Returns a list of dicts (numpy arrays trimmed to n_nodes).
    `timing`: optional dict that receives `kernel_ms` (CUDA events around the launch).
    `exec_mode`: "warp" (a warp per query), "cta" (a CTA per query: small batches), "auto"; same results."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda" if device is None else device)
    starts = np.asarray(starts, dtype=np.float64).reshape(-1, 3)
    goals = np.asarray(goals, dtype=np.float64).reshape(-1, 3)
    q = starts.shape[0]
    cap = max_iter + 1 if steer == "dubins" else 2 * max_iter + 1   # try_goal_path can append a second node
    rows, counts = engine.pack_obstacles(obstacle_lists, robot_radius)
    p = _lib.DubinsParams()
    p.n_queries, p.max_iter, p.node_cap, p.obs_stride, p.near_cap = q, max_iter, cap, rows.shape[1], near_cap
    p.search_until_max_iter = int(bool(search_until_max_iter))
    p.curvature, p.step_size = float(curvature), 0.1 if steer == "dubins" else float(step_size)
    p.goal_xy_th, p.goal_yaw_th = float(goal_xy_th), float(goal_yaw_th)
    p.rs_cost = int(bool(rs_cost))          # rrt_10's RRTStarReedsShepp: Reeds-Shepp-length costs (steer="rs" only)
    p.exec_mode = {"auto": 0, "warp": 1, "cta": 2}[exec_mode]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    with torch.cuda.device(dev):
        d_sg, d_obs, d_cnt = t(np.hstack([starts, goals])), t(rows), t(counts)
        d_r2 = t(engine.near_r2_table(cap, connect_circle_dist, expand_dis))
        d_st = t(np.asarray(streams, dtype=np.float64).reshape(q, max_iter, 3))
        f64 = lambda *s: torch.empty(s, dtype=torch.float64, device=dev)  # noqa: E731
        i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)    # noqa: E731
        xy, yaw, cost, parent = f64(q, cap, 2), f64(q, cap), f64(q, cap), i32(q, cap)
        ef, et = f64(q, cap, 3), f64(q, cap, 3)
        n_nodes, iters, gi, status, ws = i32(q), i32(q), i32(q), i32(q), i32(q * 4 * cap + _lib.WS_TAIL_INTS)
        entry = _lib.lib().rrtk_rrtstar_dubins_run_dev if steer == "dubins" else _lib.lib().rrtk_rrtstar_rs_run_dev
        if timing is not None:
            ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev[0].record()
        rc = entry(
            C.byref(p), d_sg.data_ptr(), d_obs.data_ptr(), d_cnt.data_ptr(), d_r2.data_ptr(), d_st.data_ptr(),
            xy.data_ptr(), yaw.data_ptr(), cost.data_ptr(), parent.data_ptr(), ef.data_ptr(), et.data_ptr(),
            n_nodes.data_ptr(), iters.data_ptr(), gi.data_ptr(), status.data_ptr(), ws.data_ptr(),
            torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "rrtk_rrtstar_dubins_run_dev" if steer == "dubins" else "rrtk_rrtstar_rs_run_dev")
        if timing is not None:
            ev[1].record()
            ev[1].synchronize()
            timing["kernel_ms"] = ev[0].elapsed_time(ev[1])
        h = {k: v.cpu().numpy() for k, v in dict(xy=xy, yaw=yaw, cost=cost, parent=parent, ef=ef, et=et, n=n_nodes,
                                                 it=iters, gi=gi, st=status).items()}
    out = []
    for i in range(q):
        k = int(h["n"][i])
        out.append(dict(x=h["xy"][i, :k, 0].copy(), y=h["xy"][i, :k, 1].copy(), yaw=h["yaw"][i, :k].copy(),
                        cost=h["cost"][i, :k].copy(), parent=h["parent"][i, :k].copy(),
                        edge_from=h["ef"][i, :k].copy(), edge_to=h["et"][i, :k].copy(), n=k,
                        iters_done=int(h["it"][i]), goal_index=int(h["gi"][i]), status=int(h["st"][i])))
    return out


def final_course(tree, start, goal, curvature):
    """generate_final_course (rrt_05:1512-1521): the reversed course samples of every edge from the goal node up
    to the root, regenerated on the GPU from the stored pose pairs."""
    gi = tree["goal_index"]
    if gi < 0:
        return None
    chain = []
    i = gi
    while tree["parent"][i] >= 0:
        chain.append(i)
        i = int(tree["parent"][i])
    path = [[float(goal[0]), float(goal[1])]]
    if chain:
        r = dubins.steer_batch(tree["edge_from"][chain], tree["edge_to"][chain], curvature, 0.1, max_pts=1)
        mp = int(r["n_pts"].max())
        r = dubins.steer_batch(tree["edge_from"][chain], tree["edge_to"][chain], curvature, 0.1, max_pts=mp)
        for j in range(len(chain)):
            pts = r["pts"][j, :int(r["n_pts"][j])]
            path.extend(pts[::-1, 0:2].tolist())
    path.append([float(start[0]), float(start[1])])
    return path



class _DubinsSteps:
    """The reference's per-step methods of the Dubins planners (rrt_05:1458-1479, :1605-1638, :1648-1779; rrt_03 shares the
    first group), device-backed like rrtk.RRT's: `steer` is one launch of the Dubins steering kernel, the scans and the
    point-list collision test are the FP64 primitives of include/rrtk.h.  planning() is the fused kernel and refuses to run
    when one of these is overridden (it would be ignored); an overridden get_random_node is honoured."""

    _FUSED_METHODS = ("steer", "check_collision", "check_if_outside_play_area", "get_nearest_node_index",
                      "calc_distance_and_angle", "choose_parent", "rewire", "find_near_nodes", "calc_new_cost",
                      "propagate_cost_to_leaves", "search_best_goal_node")

    check_collision = staticmethod(_RRT2D.check_collision)
    check_if_outside_play_area = staticmethod(_RRT2D.check_if_outside_play_area)
    get_nearest_node_index = staticmethod(_RRT2D.get_nearest_node_index)
    calc_distance_and_angle = staticmethod(_RRT2D.calc_distance_and_angle)
    _overridden = _RRT2D._overridden
    find_near_nodes = _RRTStar2D.find_near_nodes
    calc_new_cost = _RRTStar2D.calc_new_cost
    propagate_cost_to_leaves = _RRTStar2D.propagate_cost_to_leaves

    def _check_overrides(self, base):
        bad = self._overridden([m for m in self._FUSED_METHODS if hasattr(base, m)], base)
        if bad:
            raise _lib.RrtkError("planning() runs the whole loop as one fused GPU kernel; the overridden method(s) " + ", ".join(bad)
                                 + " would be ignored.  Call the per-step methods yourself (they are device-backed), or plan "
                                 "with the stock class.")

    def steer(self, from_node, to_node):
        """rrt_05:1458-1479: the Dubins course from_node -> to_node as a new node (None when there is no course)."""
        px, py, pyaw, mode, lengths = dubins.plan_dubins_path(from_node.x, from_node.y, from_node.yaw, to_node.x, to_node.y,
                                                              to_node.yaw, self.curvature)
        if len(px) <= 1:
            return None
        nd = Node(float(px[-1]), float(py[-1]), float(pyaw[-1]))
        nd._course = (list(px), list(py), list(pyaw))
        nd.cost = from_node.cost + sum([abs(c) for c in lengths])
        nd.parent = from_node
        return nd

    def _course_ok(self, a, b):
        """steer(a, b) and whether its course is collision free (None / False when there is no course)."""
        e = self.steer(a, b)
        return e, bool(e) and self.check_collision(e, self.obstacle_list, self.robot_radius)

    def choose_parent(self, new_node, near_inds):
        """rrt_05:1648-1689: the near node with the cheapest (Euclidean, :1777-1779) cost over a collision-free course;
        first minimum in list order, None when every course is blocked."""
        if not near_inds:
            return None
        costs = [self.calc_new_cost(self.node_list[i], new_node) if self._course_ok(self.node_list[i], new_node)[1] else math.inf
                 for i in near_inds]
        best = min(costs)
        if best == math.inf:
            return None
        chosen = self.steer(self.node_list[near_inds[costs.index(best)]], new_node)
        chosen.cost = best
        return chosen

    def rewire(self, new_node, near_inds):
        """rrt_05:1741-1775: in list order, a near node that new_node reaches cheaper over a collision-free course is replaced
        by the end of that course (it MOVES there), its children follow and their costs are refreshed."""
        for i in near_inds:
            old = self.node_list[i]
            via, free = self._course_ok(new_node, old)
            if via is None:
                continue
            via.cost = self.calc_new_cost(new_node, old)
            if free and old.cost > via.cost:
                for child in self.node_list:
                    if child.parent is old:
                        child.parent = via
                self.node_list[i] = via
                self.propagate_cost_to_leaves(via)


class RRTStarDubins(_DubinsSteps):
    """rrt_05's `RRT`: RRT* with Dubins steering, same constructor keywords and defaults (rrt_05:1358-1375)."""

    Node = Node

    def __init__(self, start, goal, obstacle_list, rand_area, expand_dis=3.0, path_resolution=0.5,
                 goal_sample_rate=5, max_iter=500, play_area=None, robot_radius=0.0, sobol_sampler=True,
                 connect_circle_dist=50.0, search_until_max_iter=False, curvature=1.0,
                 goal_yaw_th=np.deg2rad(1.0), goal_xy_th=0.5, near_cap=256):
        self.start = Node(start[0], start[1], start[2])
        self.end = Node(goal[0], goal[1], goal[2])
        self.min_rand, self.max_rand = rand_area[0], rand_area[1]
        self.play_area = play_area           # accepted; rrt_05 never checks it (:1430-1432)
        self.expand_dis, self.path_resolution = expand_dis, path_resolution
        self.goal_sample_rate, self.max_iter = goal_sample_rate, max_iter
        self.obstacle_list = obstacle_list
        self.node_list = []
        self.robot_radius = robot_radius
        self.sobol_sampler, self.sobol_inter_ = sobol_sampler, 0
        self.connect_circle_dist = connect_circle_dist
        self.search_until_max_iter = search_until_max_iter
        self.curvature, self.goal_yaw_th, self.goal_xy_th = curvature, goal_yaw_th, goal_xy_th
        self.near_cap = near_cap
        self._tree = None

    def get_random_node(self):
        """rrt_05:1528-1538 (the sampler `planning` uses)."""
        if random.randint(0, 100) > self.goal_sample_rate:
            return (random.uniform(self.min_rand, self.max_rand), random.uniform(self.min_rand, self.max_rand),
                    random.uniform(-pi, pi))
        return (self.end.x, self.end.y, self.end.yaw)

    def planning(self, animation=True, search_until_max_iter=True, sample_stream=None):
        self._check_overrides(RRTStarDubins)
        n = int(self.max_iter)
        rng_state = None
        if sample_stream is None:
            rng_state = random.getstate()
            sample_stream = np.array([self.get_random_node() for _ in range(n)], dtype=np.float64)
        stream = np.asarray(sample_stream, dtype=np.float64).reshape(-1, 3)[:n]
        start = (self.start.x, self.start.y, self.start.yaw)
        goal = (self.end.x, self.end.y, self.end.yaw)
        t = run_batch([start], [goal], [list(self.obstacle_list)], self.expand_dis, n, stream[None],
                      self.robot_radius, self.connect_circle_dist, self.curvature, self.goal_yaw_th,
                      self.goal_xy_th, search_until_max_iter, self.near_cap)[0]
        if t["status"] & _lib.Q_NEAR_OVERFLOW:
            raise _lib.RrtkError("near list overflow: raise near_cap")
        self._tree = t
        if rng_state is not None and t["iters_done"] < n:
            # early exit: leave `random` where the reference's lazily drawing loop would have left it
            random.setstate(rng_state)
            for _ in range(t["iters_done"]):
                self.get_random_node()
        nodes = [Node(float(x), float(y), float(w)) for x, y, w in zip(t["x"], t["y"], t["yaw"])]
        for i, nd in enumerate(nodes):
            nd.cost = float(t["cost"][i])
            if t["parent"][i] >= 0:
                nd.parent = nodes[t["parent"][i]]
                nd._edge = (t["edge_from"][i], t["edge_to"][i], self.curvature)
        self.node_list = nodes
        return final_course(t, start, goal, self.curvature)

    def tree_arrays(self):
        return self._tree


def run_rrt_batch(starts, goals, obstacle_lists, max_iter, streams, robot_radius=0.0, curvature=1.0,
                  goal_yaw_th=np.deg2rad(1.0), goal_xy_th=0.5, search_until_max_iter=True, play_area=None, device=None,
                  timing=None):
    """Q RRT-Dubins queries (rrt_03 semantics, rrtk_rrt_dubins_run_dev) in one launch.  starts / goals [Q, 3]; streams
    [Q, max_iter, 3]; play_area = None or (xmin, xmax, ymin, ymax) for the whole batch.  Returns a list of dicts like
    run_batch (status has Q_NONE_STEER where the reference would raise AttributeError)."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda" if device is None else device)
    starts = np.asarray(starts, dtype=np.float64).reshape(-1, 3)
    goals = np.asarray(goals, dtype=np.float64).reshape(-1, 3)
    q, cap = starts.shape[0], max_iter + 1
    rows, counts = engine.pack_obstacles(obstacle_lists, robot_radius)
    p = _lib.DubinsParams()
    p.n_queries, p.max_iter, p.node_cap, p.obs_stride, p.near_cap = q, max_iter, cap, rows.shape[1], 32
    p.search_until_max_iter = int(bool(search_until_max_iter))
    p.curvature, p.step_size = float(curvature), 0.1
    p.goal_xy_th, p.goal_yaw_th = float(goal_xy_th), float(goal_yaw_th)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    with torch.cuda.device(dev):
        d_sg, d_obs, d_cnt = t(np.hstack([starts, goals])), t(rows), t(counts)
        d_st = t(np.asarray(streams, dtype=np.float64).reshape(q, max_iter, 3))
        d_play = None if play_area is None else t(np.asarray(play_area, dtype=np.float64).reshape(4))
        f64 = lambda *s: torch.empty(s, dtype=torch.float64, device=dev)  # noqa: E731
        i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)    # noqa: E731
        xy, yaw, cost, parent = f64(q, cap, 2), f64(q, cap), f64(q, cap), i32(q, cap)
        ef, et = f64(q, cap, 3), f64(q, cap, 3)
        n_nodes, iters, gi, status, ws = i32(q), i32(q), i32(q), i32(q), i32(_lib.WS_TAIL_INTS)
        if timing is not None:
            ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev[0].record()
        rc = _lib.lib().rrtk_rrt_dubins_run_dev(
            C.byref(p), d_sg.data_ptr(), d_obs.data_ptr(), d_cnt.data_ptr(), None if d_play is None else d_play.data_ptr(),
            d_st.data_ptr(), xy.data_ptr(), yaw.data_ptr(), cost.data_ptr(), parent.data_ptr(), ef.data_ptr(), et.data_ptr(),
            n_nodes.data_ptr(), iters.data_ptr(), gi.data_ptr(), status.data_ptr(), ws.data_ptr(),
            torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "rrtk_rrt_dubins_run_dev")
        if timing is not None:
            ev[1].record()
            ev[1].synchronize()
            timing["kernel_ms"] = ev[0].elapsed_time(ev[1])
        h = {k: v.cpu().numpy() for k, v in dict(xy=xy, yaw=yaw, cost=cost, parent=parent, ef=ef, et=et, n=n_nodes,
                                                 it=iters, gi=gi, st=status).items()}
    out = []
    for i in range(q):
        k = int(h["n"][i])
        out.append(dict(x=h["xy"][i, :k, 0].copy(), y=h["xy"][i, :k, 1].copy(), yaw=h["yaw"][i, :k].copy(),
                        cost=h["cost"][i, :k].copy(), parent=h["parent"][i, :k].copy(),
                        edge_from=h["ef"][i, :k].copy(), edge_to=h["et"][i, :k].copy(), n=k,
                        iters_done=int(h["it"][i]), goal_index=int(h["gi"][i]), status=int(h["st"][i])))
    return out


class RRTDubins(_DubinsSteps):
    """rrt_03's `RRT`: plain RRT with Dubins steering, same constructor keywords and defaults (rrt_03:1370-1383).
    `planning(animation=True, search_until_max_iter=True)` returns the sampled final course (goal -> start) or None.
    Unlike rrt_05, `sobol_sampler` is honoured (:1430-1433): x, y from points of the 3-D Sobol sequence mapped to
    rand_area, yaw = -pi + q * pi (:1545-1562); the goal coin and the uniform sampler draw from `random` exactly as the
    reference does.  The play-area test applies to the new node's end pose (:1437).  No CPU fallback."""

    Node = Node

    def __init__(self, start, goal, obstacle_list, rand_area, goal_sample_rate=10, max_iter=200, play_area=None,
                 robot_radius=0.0, sobol_sampler=False, curvature=1.0, goal_yaw_th=np.deg2rad(1.0), goal_xy_th=0.5):
        self.start = Node(start[0], start[1], start[2])
        self.end = Node(goal[0], goal[1], goal[2])
        self.min_rand, self.max_rand = rand_area[0], rand_area[1]
        self.play_area = play_area
        self.goal_sample_rate, self.max_iter = goal_sample_rate, max_iter
        self.obstacle_list = obstacle_list
        self.node_list = []
        self.robot_radius = robot_radius
        self.sobol_inter_, self.sobol_sampler = 0, sobol_sampler
        self.curvature, self.goal_yaw_th, self.goal_xy_th = curvature, goal_yaw_th, goal_xy_th
        self._tree = None

    def _draw_stream(self, n):
        """The samples `get_random_node` / `get_random_node_sobol` return over n iterations (rrt_03:1528-1562), drawing
        from `random` call for call like the reference; returns (stream [n, 3], is_goal [n])."""
        from . import sampling
        goal = (self.end.x, self.end.y, self.end.yaw)
        stream = np.empty((n, 3), dtype=np.float64)
        is_goal = np.zeros(n, dtype=bool)
        if self.sobol_sampler:
            for i in range(n):
                is_goal[i] = not (random.randint(0, 100) > self.goal_sample_rate)
            k = int((~is_goal).sum())
            pts = sampling.sobol_points(3, self.sobol_inter_, k)
            xy = self.min_rand + pts[:, 0:2] * (self.max_rand - self.min_rand)
            stream[~is_goal] = np.column_stack([xy, -pi + pts[:, 2] * pi])
            stream[is_goal] = goal
            return stream, is_goal
        for i in range(n):
            if random.randint(0, 100) > self.goal_sample_rate:
                stream[i] = (random.uniform(self.min_rand, self.max_rand), random.uniform(self.min_rand, self.max_rand),
                             random.uniform(-pi, pi))
            else:
                is_goal[i] = True
                stream[i] = goal
        return stream, is_goal

    def planning(self, animation=True, search_until_max_iter=True, sample_stream=None):
        self._check_overrides(RRTDubins)
        n = int(self.max_iter)
        rng_state, is_goal = None, None
        if sample_stream is None:
            rng_state = random.getstate()
            sample_stream, is_goal = self._draw_stream(n)
        stream = np.asarray(sample_stream, dtype=np.float64).reshape(-1, 3)[:n]
        start = (self.start.x, self.start.y, self.start.yaw)
        goal = (self.end.x, self.end.y, self.end.yaw)
        t = run_rrt_batch([start], [goal], [list(self.obstacle_list)], n, stream[None], self.robot_radius, self.curvature,
                          self.goal_yaw_th, self.goal_xy_th, search_until_max_iter, self.play_area)[0]
        self._tree = t
        done = t["iters_done"]
        if is_goal is not None:
            # leave `random` and sobol_inter_ where the reference's lazily drawing loop would have left them
            self.sobol_inter_ += int((~is_goal[:done]).sum()) if self.sobol_sampler else 0
            if done < n:
                random.setstate(rng_state)
                for i in range(done):
                    if random.randint(0, 100) > self.goal_sample_rate and not self.sobol_sampler:
                        random.uniform(0, 1), random.uniform(0, 1), random.uniform(0, 1)
        nodes = [Node(float(x), float(y), float(w)) for x, y, w in zip(t["x"], t["y"], t["yaw"])]
        for i, nd in enumerate(nodes):
            nd.cost = float(t["cost"][i])
            if t["parent"][i] >= 0:
                nd.parent = nodes[t["parent"][i]]
                nd._edge = (t["edge_from"][i], t["edge_to"][i], self.curvature)
        self.node_list = nodes
        if t["status"] & _lib.Q_NONE_STEER:
            raise AttributeError("'NoneType' object has no attribute 'x'")       # what rrt_03:1626 raises here
        return final_course(t, start, goal, self.curvature)

    def tree_arrays(self):
        return self._tree
