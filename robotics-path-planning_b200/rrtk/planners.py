"""The reference's planner classes, backed by the GPU kernels.

Same constructor keywords, same entry points and the same return values as the classes of
gouldberg/robotics-path-planning:
  * `RRT`      -- rrt_01:16-233 (== rrt_02 with `sobol_sampler=True`, == rrt_10:104-342)
  * `RRTStar`  -- the class the reference calls `RRT` in rrt_04:932-1384 (== `RRTStar`, rrt_10:350-577)
`planning(animation=...)` returns `[[x, y], ...]` ordered goal -> start, or None; afterwards
`node_list` holds `Node` objects (`x, y, path_x, path_y, parent, cost`).  `animation` is accepted and
ignored (the GPU path never draws).  The whole loop runs in one kernel launch; there is no CPU
fallback."""
from __future__ import annotations

import inspect
import math
import random

import numpy as np

from . import _lib, engine, sampling


class Node:
    """Tree vertex (rrt_04:933-942): x, y, path_x, path_y (the sampled edge from the parent, as steer built it), parent,
    cost."""

    def __init__(self, x, y):
        self.x = x
        self.y = y
        self.path_x = []
        self.path_y = []
        self.parent = None
        self.cost = 0.0

    def __repr__(self):
        return f"Node(x={self.x!r}, y={self.y!r}, cost={self.cost!r})"


class AreaBounds:
    """rrt_04:944-949."""

    def __init__(self, area):
        self.xmin = float(area[0])
        self.xmax = float(area[1])
        self.ymin = float(area[2])
        self.ymax = float(area[3])


class RRT:
    """Basic RRT (rrt_01:16-101).  `sobol_sampler=True` gives rrt_02's sampler (rrt_02:1077-1089)."""

    Node = Node
    AreaBounds = AreaBounds
    _rrt_only = True

    def __init__(self, start, goal, obstacle_list, rand_area, expand_dis=3.0, path_resolution=0.5,
                 goal_sample_rate=5, max_iter=500, play_area=None, robot_radius=0.0,
                 sobol_sampler=False):
        self.start = Node(start[0], start[1])
        self.end = Node(goal[0], goal[1])
        self.min_rand = rand_area[0]
        self.max_rand = rand_area[1]
        self.play_area = AreaBounds(play_area) if play_area is not None else None
        self.expand_dis = expand_dis
        self.path_resolution = path_resolution
        self.goal_sample_rate = goal_sample_rate
        self.max_iter = max_iter
        self.obstacle_list = obstacle_list
        self.node_list = []
        self.robot_radius = robot_radius
        self.sobol_sampler = sobol_sampler
        self.sobol_inter_ = 0
        # filled by planning()
        self.sample_stream = None   # [iters, 2] the samples the loop consumed
        self.iters_done = 0
        self.goal_index = None
        self.status = 0
        self.trace = None
        self._arrays = None

    # ---- configuration hooks shared with RRTStar ----
    def _near_table(self, node_cap):
        return None

    def _search_until_max_iter(self):
        return False

    def _play_tuple(self):
        a = self.play_area
        return None if a is None else (a.xmin, a.xmax, a.ymin, a.ymax)

    # ---- what the fused kernel replaces.  A subclass (or an instance attribute) that overrides one of these cannot
    # take effect inside rrtstar_kernel: planning() refuses to run rather than ignore it.  The samplers are the
    # exception: an overridden get_random_node / get_random_node_sobol is called once per iteration to build the
    # sample stream (that is how the reference's own golden-vector generator injects samples).
    _FUSED_METHODS = ("steer", "check_collision", "check_if_outside_play_area", "get_nearest_node_index",
                      "calc_distance_and_angle", "calc_dist_to_goal", "generate_final_course")
    _SAMPLER_METHODS = ("get_random_node", "get_random_node_sobol")

    def _overridden(self, names, base):
        out = []
        for nm in names:
            if nm in self.__dict__:
                out.append(nm)
                continue
            mine = inspect.getattr_static(type(self), nm, None)
            theirs = inspect.getattr_static(base, nm, None)
            if mine is not theirs:
                out.append(nm)
        return out

    def _check_overrides(self):
        bad = self._overridden(self._FUSED_METHODS, self._api_base())
        if bad:
            raise _lib.RrtkError(
                "planning() runs the whole loop as one fused GPU kernel (rrtk_rrtstar_run_dev); the overridden method(s) "
                + ", ".join(bad) + " would be ignored.  Call the per-step methods yourself (they are device-backed), or "
                "plan with the stock class.")

    @classmethod
    def _api_base(cls):
        return RRT

    def planning(self, animation=True, sample_stream=None, want_trace=False):
        """Run the whole planning loop on the GPU.  `sample_stream` ([max_iter, 2]) injects the
        samples instead of drawing them from `random` (used by the parity tests)."""
        self._check_overrides()
        torch = _lib.require_cuda()
        max_iter = int(self.max_iter)
        goal = (float(self.end.x), float(self.end.y))
        rng_state, sobol_first = None, self.sobol_inter_
        own = "get_random_node_sobol" if self.sobol_sampler else "get_random_node"
        if sample_stream is None and self._overridden((own,), self._api_base()):
            # the caller's own sampler: one call per iteration, as the reference's loop makes them (rrt_04:1046-1049)
            sample_stream = np.array([[float(nd.x), float(nd.y)] for nd in (getattr(self, own)() for _ in range(max_iter))],
                                     dtype=np.float64).reshape(max_iter, 2)
        if sample_stream is None:
            rng_state = random.getstate()
            stream, _, nxt = sampling.draw_stream(max_iter, goal, self.min_rand, self.max_rand,
                                                  self.goal_sample_rate, self.sobol_sampler,
                                                  self.sobol_inter_, random)
            self.sobol_inter_ = nxt
        else:
            stream = np.ascontiguousarray(np.asarray(sample_stream, dtype=np.float64).reshape(-1, 2))
            if stream.shape[0] < max_iter:
                raise ValueError("sample_stream is shorter than max_iter")
            stream = stream[:max_iter]
        node_cap = max_iter + 1
        rows, n_obs = engine.pack_obstacles([list(self.obstacle_list)], self.robot_radius)
        p = engine.make_params(1, max_iter, node_cap, rows.shape[1], self.expand_dis,
                               self.path_resolution, self._play_tuple(),
                               self._search_until_max_iter(), _lib.SAMPLER_STREAM,
                               self.goal_sample_rate, self.min_rand, self.max_rand,
                               near_cap=self._near_cap(node_cap), rrt_only=self._rrt_only,
                               near_r_max=engine.near_r_max_of(self._near_table(node_cap), self.expand_dis))
        if max_iter > 0:   # obstacle cell grid over everything a node can be (samples, start, goal)
            bx = [float(stream[:, 0].min()), float(stream[:, 0].max()), float(self.start.x), goal[0]]
            by = [float(stream[:, 1].min()), float(stream[:, 1].max()), float(self.start.y), goal[1]]
            engine.set_obstacle_grid(p, min(bx), max(bx), min(by), max(by))
        dev = torch.device("cuda")
        sg = torch.tensor([[float(self.start.x), float(self.start.y), goal[0], goal[1]]],
                          dtype=torch.float64, device=dev)
        near = self._near_table(node_cap)
        res = engine.run_dev(
            p, sg, torch.from_numpy(rows).to(dev), torch.from_numpy(n_obs).to(dev),
            None if near is None else torch.from_numpy(near).to(dev),
            sample_stream=torch.from_numpy(stream.reshape(1, max_iter, 2)).to(dev),
            want_trace=True)     # (the decision trace also tells how each node's path_x was made)
        n = int(res.n_nodes[0].item())
        self.iters_done = int(res.iters_done[0].item())
        if rng_state is not None and self.iters_done < max_iter:
            # early exit (first goal connection): the reference draws lazily, so leave `random` and sobol_inter_ where
            # ITS loop would have left them -- a seeded script that goes on to path_smoothing / a second planning()
            # then sees the reference's draws
            random.setstate(rng_state)
            self.sobol_inter_ = sobol_first + sampling.consume_draws(
                self.iters_done, self.goal_sample_rate, self.sobol_sampler, self.min_rand, self.max_rand, random)
        gi = int(res.goal_index[0].item())
        self.goal_index = None if gi < 0 else gi
        self.status = int(res.status[0].item())
        self.sample_stream = stream[:self.iters_done]
        trace = res.trace[0, :self.iters_done].cpu().numpy()
        self.trace = trace if want_trace else None
        xy = res.xy[0, :n].cpu().numpy()
        cost = res.cost[0, :n].cpu().numpy()
        parent = res.parent[0, :n].cpu().numpy()
        self._arrays = dict(x=xy[:, 0].copy(), y=xy[:, 1].copy(), cost=cost, parent=parent)
        self._materialise(xy, cost, parent, trace, stream)
        if self.status & _lib.Q_NEAR_OVERFLOW:
            raise _lib.RrtkError("near list overflow: raise near_cap")
        if self.goal_index is None:
            return None
        return self.generate_final_course(self.goal_index)

    def _near_cap(self, node_cap):
        return 32

    def _materialise(self, xy, cost, parent, trace, stream):
        """node_list as the reference leaves it: Node objects with parent links, costs and the path_x / path_y of the steer
        call that made each node's current edge (rrt_04:1086-1115), regenerated on the device from what the kernel
        recorded: a node whose parent is OLDER than itself still carries the edge of its creation -- the first steer
        towards the sample when choose_parent found nothing (trace status 2), else the re-steer from the chosen parent to
        the first steer's end point (:1279); a node with a YOUNGER parent was re-parented by that parent's rewire
        (:1361-1371) and carries steer(parent, node).  (A re-parented node that MOVED -- its exact steer stopped short,
        a sub-ulp event -- gets the path of the same steer call aimed at its final position.)"""
        n = len(xy)
        nodes = [Node(float(x), float(y)) for x, y in xy]
        for i, nd in enumerate(nodes):
            nd.cost = float(cost[i])
            nd.parent = None if parent[i] < 0 else nodes[parent[i]]
        self.node_list = nodes
        if n <= 1:
            return
        res = float(self.path_resolution)
        born = np.zeros(n, dtype=np.int64)                  # iteration that appended node k
        n_after = trace[:, 7]
        grew = np.flatnonzero(np.diff(np.concatenate([[1], n_after])) > 0)
        born[1:] = grew[:n - 1]
        ks = np.arange(1, n)
        it = born[1:]
        first_from = xy[trace[it, 0]]                       # the nearest node of that iteration
        first_to = stream[it]
        kept_first = (trace[it, 1] == 2) | self._rrt_only   # appended with the first steer's own edge
        rewired = parent[1:] > ks
        frm = np.where(rewired[:, None] | ~kept_first[:, None], xy[parent[1:]], first_from)
        to = np.where(rewired[:, None], xy[1:], first_to)
        ext = np.where(rewired | ~kept_first, np.inf, float(self.expand_dis))
        resteer = ~rewired & ~kept_first
        if resteer.any():                                   # target = where the first steer ended (the node before :1279)
            r = engine.steer_collide(first_from[resteer], first_to[resteer], [], float(self.expand_dis), res)
            to[resteer] = r["new_xy"]
        pts, npts = engine.steer_points(frm, to, ext, res)
        for j, k in enumerate(ks):
            m = int(npts[j])
            nodes[k].path_x = pts[j, :m, 0].tolist()
            nodes[k].path_y = pts[j, :m, 1].tolist()

    def tree_arrays(self):
        """dict(x, y, cost, parent) of the final tree as numpy arrays (parent = -1 for the root)."""
        return self._arrays

    def generate_final_course(self, goal_ind):
        """rrt_04:1117-1125."""
        path = [[self.end.x, self.end.y]]
        node = self.node_list[goal_ind]
        while node.parent is not None:
            path.append([node.x, node.y])
            node = node.parent
        path.append([node.x, node.y])
        return path

    # ---- the reference's per-step methods (rrt_04:1086-1238), device-backed: each call is one launch of the stand-alone
    # form of what the fused kernel does inside an iteration (include/rrtk.h: rrtk_steer_points_dev,
    # rrtk_points_collide_dev, rrtk_nearest_f64_dev).  They exist so that code written against the reference's class
    # keeps working; a loop built from them pays a launch and a copy per call, which is what planning() avoids. ----
    def steer(self, from_node, to_node, extend_length=float("inf")):
        """rrt_04:1086-1115: a new node at most `extend_length` from `from_node` towards `to_node`, with its sampled edge
        in path_x / path_y and parent = from_node."""
        pts, npts = engine.steer_points([[float(from_node.x), float(from_node.y)]], [[float(to_node.x), float(to_node.y)]],
                                        float(extend_length), float(self.path_resolution))
        m = int(npts[0])
        new_node = self.Node(float(pts[0, m - 1, 0]), float(pts[0, m - 1, 1]))
        new_node.path_x = pts[0, :m, 0].tolist()
        new_node.path_y = pts[0, :m, 1].tolist()
        new_node.parent = from_node
        return new_node

    @staticmethod
    def check_collision(node, obstacleList, robot_radius):
        """rrt_04:1216-1230: False when a point of the node's path lies within size + robot_radius of a circle."""
        if node is None:
            return False
        if len(obstacleList) == 0:
            return True
        return bool(engine.points_collide([list(zip(node.path_x, node.path_y))], obstacleList, robot_radius)[0])

    @staticmethod
    def check_if_outside_play_area(node, play_area):
        """rrt_04:1204-1214: True = the node is inside the play area (or there is none)."""
        if play_area is None:
            return True
        return not (node.x < play_area.xmin or node.x > play_area.xmax or node.y < play_area.ymin or node.y > play_area.ymax)

    @staticmethod
    def get_nearest_node_index(node_list, rnd_node):
        """rrt_04:1196-1202: index of the first node at minimum squared distance."""
        xy = np.array([[float(nd.x), float(nd.y)] for nd in node_list], dtype=np.float64)
        return int(engine.nearest_index(xy, [[float(rnd_node.x), float(rnd_node.y)]])[0])

    def get_random_node(self):
        """rrt_04:1132-1140 (draws from `random` exactly as the reference does)."""
        if random.randint(0, 100) > self.goal_sample_rate:
            return self.Node(random.uniform(self.min_rand, self.max_rand), random.uniform(self.min_rand, self.max_rand))
        return self.Node(self.end.x, self.end.y)

    def get_random_node_sobol(self):
        """rrt_04:1143-1155: the next 2-D Sobol point (device generator, rrt_04:230-503) mapped onto the sampling range."""
        if random.randint(0, 100) > self.goal_sample_rate:
            q = sampling.sobol_points(2, self.sobol_inter_, 1)[0]
            xy = self.min_rand + q * (self.max_rand - self.min_rand)
            self.sobol_inter_ += 1
            return self.Node(*xy)
        return self.Node(self.end.x, self.end.y)

    def calc_dist_to_goal(self, x, y):
        return math.hypot(x - self.end.x, y - self.end.y)

    @staticmethod
    def calc_distance_and_angle(from_node, to_node):
        dx = to_node.x - from_node.x
        dy = to_node.y - from_node.y
        return math.hypot(dx, dy), math.atan2(dy, dx)


class RRTStar(RRT):
    """RRT* (the class named `RRT` in rrt_04:932-1384; `RRTStar` in rrt_10:350-577)."""

    _rrt_only = False

    def __init__(self, start, goal, obstacle_list, rand_area, expand_dis=3.0, path_resolution=0.5,
                 goal_sample_rate=5, max_iter=500, play_area=None, robot_radius=0.0,
                 sobol_sampler=True, connect_circle_dist=50.0, search_until_max_iter=False,
                 near_cap=None):
        super().__init__(start, goal, obstacle_list, rand_area, expand_dis, path_resolution,
                         goal_sample_rate, max_iter, play_area, robot_radius, sobol_sampler)
        self.connect_circle_dist = connect_circle_dist
        self.goal_node = Node(goal[0], goal[1])
        self.search_until_max_iter = search_until_max_iter
        self._near_cap_user = near_cap

    def _near_table(self, node_cap):
        return engine.near_r2_table(node_cap, self.connect_circle_dist, self.expand_dis)

    def _search_until_max_iter(self):
        return self.search_until_max_iter

    def _near_cap(self, node_cap):
        if self._near_cap_user is not None:
            return int(self._near_cap_user)
        return min(1024, (node_cap + 31) // 32 * 32)

    _FUSED_METHODS = RRT._FUSED_METHODS + ("choose_parent", "rewire", "find_near_nodes", "calc_new_cost",
                                           "propagate_cost_to_leaves", "search_best_goal_node")

    @classmethod
    def _api_base(cls):
        return RRTStar

    # ---- rrt_04:1242-1384 as device-backed calls over node_list (see RRT.steer) ----
    def calc_new_cost(self, from_node, to_node):
        """rrt_04:1375-1377."""
        d, _ = self.calc_distance_and_angle(from_node, to_node)
        return from_node.cost + d

    def find_near_nodes(self, new_node):
        """rrt_04:1314-1338: the nodes within min(connect_circle_dist * sqrt(log(n) / n), expand_dis) of new_node, as the
        reference lists them (the `dist_list.index(i)` mapping included)."""
        nnode = len(self.node_list) + 1
        r = self.connect_circle_dist * math.sqrt(math.log(nnode) / nnode)
        if hasattr(self, "expand_dis"):
            r = min(r, self.expand_dis)
        xy = np.array([[float(nd.x), float(nd.y)] for nd in self.node_list], dtype=np.float64)
        return engine.near_indices(xy, float(new_node.x), float(new_node.y), r ** 2)

    def _edges(self, froms, tos):
        """steer(f, t) for every pair + its collision / play-area verdict, two launches for the whole list."""
        f = [[float(a.x), float(a.y)] for a in froms]
        t = [[float(b.x), float(b.y)] for b in tos]
        pts, npts = engine.steer_points(f, t, float("inf"), float(self.path_resolution))
        out = []
        for j, a in enumerate(froms):
            m = int(npts[j])
            nd = self.Node(float(pts[j, m - 1, 0]), float(pts[j, m - 1, 1]))
            nd.path_x, nd.path_y, nd.parent = pts[j, :m, 0].tolist(), pts[j, :m, 1].tolist(), a
            out.append(nd)
        free = engine.points_collide([list(zip(nd.path_x, nd.path_y)) for nd in out], self.obstacle_list, self.robot_radius) \
            if len(self.obstacle_list) and out else np.ones(len(out), dtype=bool)
        ok = [bool(free[j]) and self.check_if_outside_play_area(nd, self.play_area) for j, nd in enumerate(out)]
        return out, ok

    def choose_parent(self, new_node, near_inds):
        """rrt_04:1242-1282: re-steer new_node from the near node that gives the lowest cost over a collision-free edge."""
        if not near_inds:
            return None
        near = [self.node_list[i] for i in near_inds]
        edges, ok = self._edges(near, [new_node] * len(near))
        costs = [self.calc_new_cost(nd, new_node) if good else float("inf") for nd, good in zip(near, ok)]
        min_cost = min(costs)
        if min_cost == float("inf"):
            return None
        k = costs.index(min_cost)
        chosen = edges[k]
        chosen.cost = min_cost
        return chosen

    def rewire(self, new_node, near_inds):
        """rrt_04:1340-1373: re-parent every near node that new_node reaches cheaper over a collision-free edge, in list
        order, propagating the new costs to its descendants."""
        for i in near_inds:
            near_node = self.node_list[i]
            (edge_node,), (no_collision,) = self._edges([new_node], [near_node])
            edge_node.cost = self.calc_new_cost(new_node, near_node)
            if no_collision and near_node.cost > edge_node.cost:
                for node in self.node_list:
                    if node.parent is near_node:
                        node.parent = edge_node
                self.node_list[i] = edge_node
                self.propagate_cost_to_leaves(edge_node)

    def propagate_cost_to_leaves(self, parent_node):
        """rrt_04:1379-1384 (children lists instead of one scan of node_list per level; the values do not depend on the
        visiting order: a node's cost is its parent's plus the edge length)."""
        kids = {}
        for node in self.node_list:
            if node.parent is not None:
                kids.setdefault(id(node.parent), []).append(node)
        stack = [parent_node]
        while stack:
            p = stack.pop()
            for node in kids.get(id(p), ()):
                node.cost = self.calc_new_cost(p, node)
                stack.append(node)

    def search_best_goal_node(self):
        """rrt_04:1284-1312 over node_list.  Straight after planning() this is the index the kernel chose (the same
        search, run on the device); after the caller changed node_list it is evaluated again with the per-step calls."""
        if self._arrays is not None and len(self.node_list) == len(self._arrays["x"]) and not self._tree_edited():
            return self.goal_index
        dist = [self.calc_dist_to_goal(n.x, n.y) for n in self.node_list]
        goal_inds = [dist.index(d) for d in dist if d <= self.expand_dis]
        if not goal_inds:
            return None
        _, ok = self._edges([self.node_list[i] for i in goal_inds], [self.goal_node] * len(goal_inds))
        safe = [i for i, good in zip(goal_inds, ok) if good]
        if not safe:
            return None
        costs = [self.node_list[i].cost + self.calc_dist_to_goal(self.node_list[i].x, self.node_list[i].y) for i in safe]
        return safe[costs.index(min(costs))]

    def _tree_edited(self):
        a = self._arrays
        return any(nd.x != a["x"][i] or nd.y != a["y"][i] or nd.cost != a["cost"][i] for i, nd in enumerate(self.node_list))

    @staticmethod
    def planning_batch(starts, goals, obstacle_lists, rand_area, **kw):
        """Q independent queries in one launch: the constructor's keywords with starts / goals [Q, 2] and one
        obstacle list per query; returns the list of paths `planning()` would return query by query (goal -> start, or
        None).  Sampling is in-kernel (`sampler="sobol"` / `"uniform"`, per-query streams), see rrtk.RRTStarBatch."""
        from .batch import RRTStarBatch
        if "sobol_sampler" in kw:
            kw["sampler"] = "sobol" if kw.pop("sobol_sampler") else "uniform"
        return RRTStarBatch(starts, goals, obstacle_lists, rand_area, **kw).planning()
