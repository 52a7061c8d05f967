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

import math
import random

import numpy as np

from . import _lib, engine, sampling


class Node:
    """Tree vertex (rrt_04:933-942).  `path_x` / `path_y` (the sampled edge from the parent) are
    rebuilt on first access from the parent's and the node's final positions."""

    def __init__(self, x, y):
        self.x = x
        self.y = y
        self.parent = None
        self.cost = 0.0
        self._path = None
        self._res = None

    def _edge(self):
        if self._path is None:
            px, py = [], []
            if self.parent is not None and self._res:
                x, y = self.parent.x, self.parent.y
                dx, dy = self.x - x, self.y - y
                d, th = math.hypot(dx, dy), math.atan2(dy, dx)
                px.append(x)
                py.append(y)
                for _ in range(math.floor(d / self._res)):
                    x += self._res * math.cos(th)
                    y += self._res * math.sin(th)
                    px.append(x)
                    py.append(y)
                px.append(self.x)
                py.append(self.y)
            self._path = (px, py)
        return self._path

    @property
    def path_x(self):
        return self._edge()[0]

    @property
    def path_y(self):
        return self._edge()[1]

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

    def planning(self, animation=True, sample_stream=None, want_trace=False):
        """Run the whole planning loop on the GPU.  `sample_stream` ([max_iter, 2]) injects the
        samples instead of drawing them from `random` (used by the parity tests)."""
        torch = _lib.require_cuda()
        max_iter = int(self.max_iter)
        goal = (float(self.end.x), float(self.end.y))
        rng_state, sobol_first = None, self.sobol_inter_
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
            want_trace=want_trace)
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
        self.trace = None if res.trace is None else res.trace[0, :self.iters_done].cpu().numpy()
        xy = res.xy[0, :n].cpu().numpy()
        cost = res.cost[0, :n].cpu().numpy()
        parent = res.parent[0, :n].cpu().numpy()
        self._arrays = dict(x=xy[:, 0].copy(), y=xy[:, 1].copy(), cost=cost, parent=parent)
        self._materialise(xy, cost, parent)
        if self.status & _lib.Q_NEAR_OVERFLOW:
            raise _lib.RrtkError("near list overflow: raise near_cap")
        if self.goal_index is None:
            return None
        return self.generate_final_course(self.goal_index)

    def _near_cap(self, node_cap):
        return 32

    def _materialise(self, xy, cost, parent):
        nodes = [Node(float(x), float(y)) for x, y in xy]
        for i, nd in enumerate(nodes):
            nd.cost = float(cost[i])
            nd._res = self.path_resolution
            nd.parent = None if parent[i] < 0 else nodes[parent[i]]
        self.node_list = nodes

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

    def search_best_goal_node(self):
        """Index chosen by the kernel's search_best_goal_node (rrt_04:1284-1312), or None."""
        return self.goal_index

    @staticmethod
    def planning_batch(starts, goals, obstacle_lists, rand_area, **kw):
        """Q independent queries in one launch: the constructor's keywords with starts / goals [Q, 2] and one
        obstacle list per query; returns the list of paths `planning()` would return query by query (goal -> start, or
        None).  Sampling is in-kernel (`sampler="sobol"` / `"uniform"`, per-query streams), see rrtk.RRTStarBatch."""
        from .batch import RRTStarBatch
        if "sobol_sampler" in kw:
            kw["sampler"] = "sobol" if kw.pop("sobol_sampler") else "uniform"
        return RRTStarBatch(starts, goals, obstacle_lists, rand_area, **kw).planning()
