"""Large-tree mode: trees whose node arrays no longer fit the L2 cache (10^7 .. 10^8 nodes).

`get_nearest_node_index` (rrt_04:1196-1202, rrt_07:1210-1214) and `find_near_nodes` (rrt_04:1314-1338, rrt_07:1137-1143)
are one pass over all node positions each.  For a tree of n >= 10^7 nodes that pass is HBM-bound, so it reads an FP32 mirror
of the positions (8 B per node instead of 16) with the streaming kernels of csrc/rrtk_nn.cu -- as a FILTER.  The decision is
the reference's FP64 one:

  nearest   pass 1 (nearest_kernel):  FP32 argmin, squared distance m.
            pass 2 (near_kernel):     every node whose FP32 squared distance is <= R2(m) -- the nodes that CAN be the FP64
                                      minimum given the rounding of the mirror, of the sample and of the FP32 arithmetic.
            re-check (nearest_f64):   the reference's expression (x - sx)**2 + (y - sy)**2 in FP64 over those few
                                      candidates, first minimum = dlist.index(min(dlist)).
  near      pass 1 (near_kernel):     every node whose FP32 squared distance is <= R2(r), a superset of the FP64 hits.
            re-check (near_f64):      d2 <= r**2 in FP64 over the candidates in ascending node order, with the reference's
                                      `dist_list.index(i)` mapping.

Error band.  Coordinates are bounded by M (tracked).  The mirror and the sample are rounded to FP32: e1 = 2**-24 * M per
coordinate; dx = fl(px - sx) adds 2**-24 |dx|; d2 = fl(fma(dx, dx, fl(dy * dy))) adds two more roundings.  So for every node
|d - sqrt(d2_f32)| <= beta(d) := 3 e1 + 2**-21 sqrt(d2_f32)  (the factor 3 > 2 sqrt 2 and 2**-21 > 4 * 2**-24 leave room).
A node at FP64 distance <= U is therefore kept by an FP32 threshold of ((U + 3 e1) (1 + 2**-20))**2, rounded up.

Plumbing (torch): device memory, gathers of the handful of candidates, streams.  The scans and the FP64 re-checks are the
library's kernels; there is no CPU path."""
from __future__ import annotations

import math

import numpy as np

from . import _lib, engine


def _up32(v: float) -> float:
    """smallest float32 >= v * (1 + 2**-22)"""
    f = np.float32(v * (1.0 + 2.0 ** -22))
    if float(f) < v:
        f = np.nextafter(f, np.float32(np.inf))
    return float(np.nextafter(f, np.float32(np.inf)))


class LargeTree:
    """Node store of a large tree on one GPU: FP64 positions (the truth), FP32 mirror (what the scans read)."""

    def __init__(self, capacity: int, device=None):
        torch = _lib.require_cuda()
        self.torch = torch
        self.dev = torch.device("cuda" if device is None else device)
        self.cap = int(capacity)
        self.n = 0
        self.xy64 = torch.empty((self.cap, 2), dtype=torch.float64, device=self.dev)
        self.xy32 = torch.empty((self.cap, 2), dtype=torch.float32, device=self.dev)
        self.M = 0.0                      # bound of |coordinate| over nodes and queries seen so far
        self._scratch = torch.empty(8, dtype=torch.int64, device=self.dev)
        self._idx = torch.empty(8, dtype=torch.int32, device=self.dev)
        self._d2 = torch.empty(8, dtype=torch.float32, device=self.dev)
        self._cand = torch.empty(1 << 16, dtype=torch.int32, device=self.dev)
        self._cnt = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.stats = dict(nearest_queries=0, near_queries=0, candidates=0, max_candidates=0, scans=0, scan_ms=0.0)
        self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self.time_scans = False

    # ---- growth -------------------------------------------------------------------------------------------------
    def extend(self, xy):
        """Append nodes (numpy [k, 2] or a device tensor) -- the bulk form of `node_list.append`."""
        torch = self.torch
        t = xy if torch.is_tensor(xy) else torch.from_numpy(np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2))
        t = t.to(self.dev, torch.float64)
        k = t.shape[0]
        if self.n + k > self.cap:
            raise _lib.RrtkError("LargeTree capacity exceeded")
        self.xy64[self.n:self.n + k] = t
        self.xy32[self.n:self.n + k] = t.to(torch.float32)
        if k:
            self.M = max(self.M, float(t.abs().max().item()))
        self.n += k

    def append(self, x: float, y: float) -> int:
        self.extend(np.array([[x, y]], dtype=np.float64))
        return self.n - 1

    # ---- scans ----------------------------------------------------------------------------------------------------
    def _stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def _timed(self, fn):
        if not self.time_scans:
            fn()
            self.stats["scans"] += 1
            return
        a, b = self._ev
        a.record()
        fn()
        b.record()
        b.synchronize()
        self.stats["scans"] += 1
        self.stats["scan_ms"] += a.elapsed_time(b)

    def _band(self, sx: float, sy: float):
        M = max(self.M, abs(sx), abs(sy))
        return 3.0 * (2.0 ** -24) * M

    def _near32(self, s32, r2_32: float):
        """indices (ascending, int64 tensor) of the nodes whose FP32 squared distance to s32 is <= r2_32"""
        L = _lib.lib()
        while True:
            self._cnt.zero_()
            self._timed(lambda: _lib.check(L.rrtk_near_f32_dev(self.xy32.data_ptr(), self.n, float(s32[0]), float(s32[1]), r2_32,
                                                               self._cand.data_ptr(), self._cand.numel(), self._cnt.data_ptr(),
                                                               self._stream()), "rrtk_near_f32_dev"))
            m = int(self._cnt.item())
            if m <= self._cand.numel():
                break
            self._cand = self.torch.empty(2 * m, dtype=self.torch.int32, device=self.dev)
        self.stats["candidates"] += m
        self.stats["max_candidates"] = max(self.stats["max_candidates"], m)
        return self.torch.sort(self._cand[:m].long()).values

    def nearest(self, sx: float, sy: float) -> int:
        """get_nearest_node_index: index of the FIRST node at minimum FP64 squared distance."""
        torch, L = self.torch, _lib.lib()
        if self.n < 1:
            raise _lib.RrtkError("nearest on an empty tree")
        s32 = np.array([sx, sy], dtype=np.float32)
        smp = torch.from_numpy(s32.reshape(1, 2)).to(self.dev)
        self._timed(lambda: _lib.check(L.rrtk_nearest_f32_dev(self.xy32.data_ptr(), self.n, smp.data_ptr(), 1, self._scratch.data_ptr(),
                                                              self._idx.data_ptr(), self._d2.data_ptr(), self._stream()),
                                       "rrtk_nearest_f32_dev"))
        m32 = float(self._d2[0].item())
        e3 = self._band(sx, sy)
        U = math.sqrt(m32) * (1.0 + 2.0 ** -21) + e3            # the FP64 minimum is at most this far
        R = (U + e3) * (1.0 + 2.0 ** -20)
        cand = self._near32(s32, _up32(R * R + 1e-300))
        pts = self.xy64.index_select(0, cand).contiguous()
        q = torch.tensor([[sx, sy]], dtype=torch.float64, device=self.dev)
        _lib.check(L.rrtk_nearest_f64_dev(pts.data_ptr(), pts.shape[0], q.data_ptr(), 1, self._idx.data_ptr(), None, self._stream()),
                   "rrtk_nearest_f64_dev")
        self.stats["nearest_queries"] += 1
        return int(cand[int(self._idx[0].item())].item())

    def near(self, cx: float, cy: float, r: float):
        """find_near_nodes: the reference's list for radius r (FP64 test d2 <= r**2, ascending, `.index()` mapping)."""
        torch, L = self.torch, _lib.lib()
        s32 = np.array([cx, cy], dtype=np.float32)
        e3 = self._band(cx, cy)
        R = (r + e3) * (1.0 + 2.0 ** -20)
        cand = self._near32(s32, _up32(R * R + 1e-300))
        self.stats["near_queries"] += 1
        if cand.numel() == 0:
            return []
        pts = self.xy64.index_select(0, cand).contiguous()
        out = torch.empty(cand.numel(), dtype=torch.int32, device=self.dev)
        d2 = torch.empty(cand.numel(), dtype=torch.float64, device=self.dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=self.dev)
        _lib.check(L.rrtk_near_f64_dev(pts.data_ptr(), pts.shape[0], float(cx), float(cy), float(r ** 2), out.data_ptr(), d2.data_ptr(),
                                       cand.numel(), cnt.data_ptr(), self._stream()), "rrtk_near_f64_dev")
        k = int(cnt.item())
        return cand[out[:k].long()].cpu().tolist()


class RRTLarge:
    """Basic RRT (rrt_01:16-101 / rrt_02 with the Sobol sampler) over a LargeTree: the reference's loop with
    get_nearest_node_index answered by the HBM-streaming scans and steer + check_collision + check_if_outside_play_area by
    rrtk_steer_collide_dev, one iteration at a time.  Same constructor as rrtk.RRT plus `capacity` and an optional seed tree
    (`seed_xy`, `seed_parent`: nodes the tree already holds, e.g. a tree grown elsewhere)."""

    def __init__(self, start, goal, obstacle_list, rand_area, expand_dis=3.0, path_resolution=0.5, goal_sample_rate=5,
                 max_iter=500, play_area=None, robot_radius=0.0, capacity=None, seed_xy=None, seed_parent=None, device=None):
        torch = _lib.require_cuda()
        self.torch = torch
        self.start, self.goal = (float(start[0]), float(start[1])), (float(goal[0]), float(goal[1]))
        self.obstacle_list, self.rand_area = list(obstacle_list), rand_area
        self.expand_dis, self.path_resolution = float(expand_dis), float(path_resolution)
        self.goal_sample_rate, self.max_iter = goal_sample_rate, int(max_iter)
        self.play_area, self.robot_radius = play_area, float(robot_radius)
        n_seed = 0 if seed_xy is None else len(seed_xy)
        cap = int(capacity) if capacity is not None else n_seed + self.max_iter + 1
        self.tree = LargeTree(cap, device)
        self.parent = torch.full((cap,), -1, dtype=torch.int32, device=self.tree.dev)
        if seed_xy is None:
            self.tree.append(*self.start)
        else:
            self.tree.extend(seed_xy)
            self.parent[:n_seed] = seed_parent if torch.is_tensor(seed_parent) else torch.from_numpy(np.asarray(seed_parent, dtype=np.int32))
        self.goal_index = None
        self.iters_done = 0

    def planning(self, sample_stream, animation=False):
        """rrt_01:71-101 with the samples given ([max_iter, 2]); returns the course goal -> start or None."""
        stream = np.asarray(sample_stream, dtype=np.float64).reshape(-1, 2)
        T = self.tree
        for it in range(self.max_iter):
            rx, ry = float(stream[it, 0]), float(stream[it, 1])
            ni = T.nearest(rx, ry)
            frm = T.xy64[ni].cpu().numpy()
            r = engine.steer_collide([frm], [[rx, ry]], self.obstacle_list, self.expand_dis, self.path_resolution, self.robot_radius,
                                     self.play_area, device=T.dev)
            if bool(r["inside"][0]) and bool(r["free"][0]):
                k = T.append(float(r["new_xy"][0, 0]), float(r["new_xy"][0, 1]))
                self.parent[k] = ni
            self.iters_done = it + 1
            last = T.xy64[T.n - 1].cpu().numpy()
            if math.hypot(last[0] - self.goal[0], last[1] - self.goal[1]) <= self.expand_dis:      # rrt_01:90-96
                g = engine.steer_collide([last], [self.goal], self.obstacle_list, self.expand_dis, self.path_resolution,
                                         self.robot_radius, None, device=T.dev)
                if bool(g["free"][0]):
                    self.goal_index = T.n - 1
                    return self.generate_final_course(self.goal_index)
        return None

    def generate_final_course(self, goal_ind):
        """rrt_01:117-125 over the parent array."""
        path = [[self.goal[0], self.goal[1]]]
        i = int(goal_ind)
        par = self.parent
        xy = self.tree.xy64
        while i >= 0:
            p = xy[i].cpu().numpy()
            path.append([float(p[0]), float(p[1])])
            i = int(par[i].item())
        return path
