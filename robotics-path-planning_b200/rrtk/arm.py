"""Arm C-space occupancy grid (02_arm_obstacle_navigation.py) behind the reference's API.

`NLinkArm` and `get_occupancy_grid(arm, obstacles, M)` keep the reference's signatures
(arm02:79-110, :236-281); the grid itself is computed on the GPU (rrtk_arm_grid_dev), for one obstacle
set or for many at once (`get_occupancy_grids`).  No CPU fallback."""
from __future__ import annotations

from math import pi

import numpy as np

from . import _lib


class NLinkArm:
    """Planar serial arm: same constructor, attributes (`n_links`, `link_lengths`, `joint_angles`, `points`,
    `end_effector`, `lim`) and methods as arm02:236-262.  Forward kinematics is one prefix sum instead of the reference's
    per-joint loop: link k points along the sum of the first k joint angles, and a joint vector SHORTER than the link
    vector (the script passes 2 angles to a 5-link arm, arm02:98) leaves the remaining links collinear, because the
    reference's `np.sum(joint_angles[:k])` saturates.  The kernel does its own FK; this class is the host-side API shell."""

    def __init__(self, link_lengths, joint_angles):
        if len(link_lengths) != len(joint_angles):
            raise ValueError()
        self.n_links = len(link_lengths)
        self.link_lengths = np.array(link_lengths)
        self.lim = sum(link_lengths)
        self.points = [[0, 0] for _ in range(self.n_links + 1)]
        self.update_joints(np.array(joint_angles))

    def update_joints(self, joint_angles):
        self.joint_angles = joint_angles
        self.update_points()

    def _headings(self):
        # heading of link k = sum of the first min(k, len(angles)) joint angles (sequential adds, like np.sum on a
        # short slice)
        ang = np.asarray(self.joint_angles, dtype=np.float64).ravel()
        prefix = np.concatenate(([0.0], np.cumsum(ang)))
        return prefix[np.minimum(np.arange(1, self.n_links + 1), ang.size)]

    def update_points(self):
        h = self._headings()
        lengths = np.asarray(self.link_lengths, dtype=np.float64)
        xs = np.concatenate(([0.0], np.cumsum(lengths * np.cos(h))))
        ys = np.concatenate(([0.0], np.cumsum(lengths * np.sin(h))))
        base = self.points[0]
        for k in range(1, self.n_links + 1):
            self.points[k][0] = base[0] + xs[k]
            self.points[k][1] = base[1] + ys[k]
        self.end_effector = np.array(self.points[self.n_links]).T


def theta_list(M: int) -> np.ndarray:
    """The first M entries of arm02:95's theta_list, with the reference's Python float arithmetic."""
    return np.array([2 * i * pi / M for i in range(-M // 2, M // 2 + 1)][:M], dtype=np.float64)


def occupancy_grids_device(link_lengths, obstacle_sets, M, row0=0, n_rows=None, device=None, cell_by_cell=False, theta=None):
    """uint8 tensor [S, n_rows, M] on the GPU: cell (i, j) of set s is 1 iff the arm at joint angles
    (theta_list[row0 + i], theta_list[j]) touches a circle of `obstacle_sets[s]` ([S, O, 3]).
    cell_by_cell=True runs rrtk_arm_grid_cells_dev (every cell on its own, the cross-check of the row rasteriser);
    `theta` replaces the reference's theta_list (tests: any other list makes the rasteriser evaluate cell by cell)."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda" if device is None else device)
    n_rows = M - row0 if n_rows is None else n_rows
    obs = np.ascontiguousarray(np.asarray(obstacle_sets, dtype=np.float64))
    if obs.ndim != 3 or obs.shape[2] != 3:
        raise ValueError("obstacle_sets must have shape [S, O, 3]")
    S, O = obs.shape[0], obs.shape[1]
    link = np.ascontiguousarray(link_lengths, dtype=np.float64)
    with torch.cuda.device(dev):
        theta = torch.from_numpy(theta_list(M) if theta is None else np.ascontiguousarray(theta, dtype=np.float64)).to(dev)
        d_obs = torch.from_numpy(obs).to(dev)
        grid = torch.empty((S, n_rows, M), dtype=torch.uint8, device=dev)
        fn = _lib.lib().rrtk_arm_grid_cells_dev if cell_by_cell else _lib.lib().rrtk_arm_grid_dev
        _lib.check(fn(M, theta.data_ptr(), row0, n_rows, len(link), link.ctypes.data, d_obs.data_ptr(), S, O, grid.data_ptr(),
                      torch.cuda.current_stream().cuda_stream), "rrtk_arm_grid_dev")
    return grid


def get_occupancy_grids(arm, obstacle_sets, M):
    """Batched form: numpy int64 array [S, M, M] (the reference's dtype)."""
    return occupancy_grids_device(arm.link_lengths, obstacle_sets, M).cpu().numpy().astype(np.int64)


def get_occupancy_grid(arm, obstacles, M):
    """Drop-in for arm02:79-110: M x M numpy int array, 1 = collision."""
    obs = np.asarray(obstacles, dtype=np.float64).reshape(1, -1, 3)
    return get_occupancy_grids(arm, obs, M)[0]


def astar_torus_batch(grids, starts, goals, route_cap=None):
    """astar_torus (arm02:113-184) for Q queries on the GPU.  grids: uint8 CUDA tensor [Q, M, M] (0 free, 1 occupied),
    updated in place with the reference's marks (2 expanded, 3 frontier, 4 start, 5 goal, 6 route); starts / goals [Q, 2]
    (row, col).  Returns (routes [Q, route_cap, 2] int32, route_len [Q], expanded [Q]) device tensors."""
    torch = _lib.require_cuda()
    if grids.dtype != torch.uint8 or not grids.is_cuda or not grids.is_contiguous() or grids.dim() != 3:
        raise _lib.RrtkError("astar_torus_batch: grids must be a contiguous uint8 CUDA tensor [Q, M, M]")
    q, M = grids.shape[0], grids.shape[1]
    dev = grids.device
    sg = np.ascontiguousarray(np.hstack([np.asarray(starts, dtype=np.int32).reshape(q, 2),
                                         np.asarray(goals, dtype=np.int32).reshape(q, 2)]))
    if (sg < 0).any() or (sg >= M).any():
        raise IndexError("start / goal cell outside the grid")
    route_cap = M * M if route_cap is None else int(route_cap)
    with torch.cuda.device(dev):
        d_sg = torch.from_numpy(sg).to(dev)
        routes = torch.zeros((q, route_cap, 2), dtype=torch.int32, device=dev)
        rlen = torch.empty((q,), dtype=torch.int32, device=dev)
        expanded = torch.empty((q,), dtype=torch.int32, device=dev)
        heur = torch.empty((q, M * M), dtype=torch.int32, device=dev)
        parents = torch.empty((q, M * M), dtype=torch.int32, device=dev)
        heaps = torch.empty((q, M * M + 8), dtype=torch.int64, device=dev)
        _lib.check(_lib.lib().rrtk_astar_torus_dev(M, q, d_sg.data_ptr(), grids.data_ptr(), routes.data_ptr(), route_cap,
                                                   rlen.data_ptr(), expanded.data_ptr(), heur.data_ptr(),
                                                   parents.data_ptr(), heaps.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream), "rrtk_astar_torus_dev")
    return routes, rlen, expanded


def astar_torus(grid, start_node, goal_node):
    """Drop-in for arm02:113-184: returns the route as a list of (row, col) tuples from start to goal ([] if none) and
    leaves the reference's marks in `grid` (a numpy array, modified in place)."""
    torch = _lib.require_cuda()
    g = np.asarray(grid)
    d_grid = torch.from_numpy(np.ascontiguousarray(g, dtype=np.uint8)[None]).cuda()
    routes, rlen, _ = astar_torus_batch(d_grid, [list(start_node)], [list(goal_node)])
    n = int(rlen[0].item())
    grid[...] = d_grid[0].cpu().numpy().astype(g.dtype)
    if n == 0:
        print("No route found.")
        return []
    route = [(int(a), int(b)) for a, b in routes[0, :n].cpu().numpy()]
    print("The route found covers %d grid cells." % len(route))
    return route
