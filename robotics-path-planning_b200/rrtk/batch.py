"""Batched planning: Q independent RRT* queries in one launch (BASELINE config 2), and the static
partition of queries across the GPUs of one box (no per-iteration collectives)."""
from __future__ import annotations

import numpy as np

from . import _lib, engine


class RRTStarBatch:
    """Q independent RRT* problems with the reference's per-query semantics (rrt_04:1036-1084).

    starts, goals      [Q, 2]
    obstacle_lists     length-Q sequence of [(x, y, size), ...]  (or a prepared [Q, O, 4] array of
                       x, y, size + rr, (size + rr) ** 2 rows together with `n_obs`)
    sampler            "sobol" | "uniform": in-kernel generator (goal coin from a counter-based RNG
                       keyed by (seed, query_base + query, iteration); Sobol index offset `(query_base + q) * max_iter`)
                       or an explicit `sample_stream` [Q, max_iter, 2].
    query_base         global index of query 0 when this batch is a shard of a larger one: the shard then
                       plans exactly what the unsharded batch plans for the same queries.
    """

    def __init__(self, starts, goals, obstacle_lists, rand_area, expand_dis=3.0,
                 path_resolution=0.5, goal_sample_rate=5, max_iter=500, play_area=None,
                 robot_radius=0.0, sampler="sobol", connect_circle_dist=50.0,
                 search_until_max_iter=True, seed=0, near_cap=256, n_obs=None,
                 sample_stream=None, sobol_offset=None, device=None, query_base=0, exec_mode=None):
        torch = _lib.require_cuda()
        self.device = torch.device("cuda" if device is None else device)
        starts = np.asarray(starts, dtype=np.float64).reshape(-1, 2)
        goals = np.asarray(goals, dtype=np.float64).reshape(-1, 2)
        self.n_queries = q = starts.shape[0]
        self.max_iter = int(max_iter)
        self._ctor = dict(rand_area=rand_area, expand_dis=expand_dis, path_resolution=path_resolution,
                          goal_sample_rate=goal_sample_rate, play_area=play_area, sampler=sampler,
                          connect_circle_dist=connect_circle_dist, search_until_max_iter=search_until_max_iter, seed=seed,
                          near_cap=near_cap)
        self.node_cap = self.max_iter + 1
        if isinstance(obstacle_lists, np.ndarray) and obstacle_lists.ndim == 3 and obstacle_lists.shape[2] == 4:
            rows = np.ascontiguousarray(obstacle_lists, dtype=np.float64)
            counts = np.full(q, rows.shape[1], np.int32) if n_obs is None else np.asarray(n_obs, np.int32)
        else:
            rows, counts = engine.pack_obstacles(obstacle_lists, robot_radius)
        self.sampler = {"stream": _lib.SAMPLER_STREAM, "sobol": _lib.SAMPLER_SOBOL,
                        "uniform": _lib.SAMPLER_UNIFORM}["stream" if sample_stream is not None else sampler]
        self.params = engine.make_params(q, self.max_iter, self.node_cap, rows.shape[1], expand_dis,
                                         path_resolution, play_area, search_until_max_iter,
                                         self.sampler, goal_sample_rate, rand_area[0], rand_area[1],
                                         seed, near_cap, exec_mode=exec_mode, query_base=query_base)
        # obstacle cell grid: box of everything a node can be (samples, starts, goals)
        pts = [starts, goals]
        if sample_stream is not None:
            ss_ = np.asarray(sample_stream, dtype=np.float64).reshape(-1, 2)
            pts.append(np.array([ss_.min(axis=0), ss_.max(axis=0)]))
        else:
            pts.append(np.array([[rand_area[0]] * 2, [rand_area[1]] * 2], dtype=np.float64))
        allp = np.vstack(pts)
        engine.set_obstacle_grid(self.params, float(allp[:, 0].min()), float(allp[:, 0].max()),
                                 float(allp[:, 1].min()), float(allp[:, 1].max()))
        # pinned host staging (inputs of the end-to-end path) + device copies
        self.h_start_goal = torch.from_numpy(np.ascontiguousarray(np.hstack([starts, goals]))).pin_memory()
        self.h_obstacles = torch.from_numpy(rows).pin_memory()
        self.h_n_obs = torch.from_numpy(counts).pin_memory()
        self.near_r2 = torch.from_numpy(
            engine.near_r2_table(self.node_cap, connect_circle_dist, expand_dis)).to(self.device)
        if sobol_offset is None:
            sobol_offset = (np.arange(q, dtype=np.int64) + int(query_base)) * self.max_iter
        self.sobol_offset = torch.from_numpy(np.asarray(sobol_offset, dtype=np.int64)).to(self.device)
        self.sample_stream = None
        if sample_stream is not None:
            ss = np.ascontiguousarray(np.asarray(sample_stream, dtype=np.float64).reshape(q, self.max_iter, 2))
            self.sample_stream = torch.from_numpy(ss).to(self.device)
        # device copies of the scenario: allocated ONCE and refilled in place by upload(), so a BatchResult, a captured
        # CUDA graph and the path extraction all keep reading the tensors the kernel planned against
        self.start_goal = torch.empty_like(self.h_start_goal, device=self.device)
        self.obstacles = torch.empty_like(self.h_obstacles, device=self.device)
        self.n_obs = torch.empty_like(self.h_n_obs, device=self.device)
        self.upload()
        self.result = None
        self._iters_done = 0

    # ---- data movement ----
    def upload(self):
        """Host -> device copy of the scenario (start/goal, obstacle rows, counts), in place: refill the pinned staging
        tensors (h_start_goal, h_obstacles, h_n_obs) and call upload() / replay() to plan a new scenario."""
        self.start_goal.copy_(self.h_start_goal, non_blocking=True)
        self.obstacles.copy_(self.h_obstacles, non_blocking=True)
        self.n_obs.copy_(self.h_n_obs, non_blocking=True)

    def h2d_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.h_start_goal, self.h_obstacles, self.h_n_obs))

    # ---- compute ----
    def run(self, want_trace=False):
        """Enqueue the planning kernel on the current stream; results stay on the GPU."""
        self.result = engine.run_dev(self.params, self.start_goal, self.obstacles, self.n_obs,
                                     self.near_r2, self.sample_stream, self.sobol_offset,
                                     want_trace=want_trace,
                                     out=self.result if (self.result is not None and not want_trace) else None)
        return self.result

    def step(self, iters: int):
        """Incremental planning: `iters` MORE iterations on the trees the previous step() left on the GPU (the first call
        starts them).  The constructor's max_iter is the total budget; k steps of m iterations build exactly the tree of
        one run of k * m iterations (same in-kernel coin / Sobol streams, or the next slice of the given sample stream)."""
        done = getattr(self, "_iters_done", 0)
        iters = int(iters)
        if iters < 0 or done + iters > self.max_iter:
            raise _lib.RrtkError("step: the total would exceed the max_iter the batch was built for")
        p = self.params
        total = p.max_iter
        stream = None
        if self.sample_stream is not None:
            stream = self.sample_stream[:, done:done + iters].contiguous()
        p.max_iter, p.resume, p.iter_offset = iters, int(done > 0), done
        try:
            self.result = engine.run_dev(p, self.start_goal, self.obstacles, self.n_obs, self.near_r2, stream,
                                         self.sobol_offset, out=self.result if done > 0 else None)
        finally:
            p.max_iter, p.resume, p.iter_offset = total, 0, 0
        self._iters_done = done + iters
        return self.result

    # ---- one planning call as a CUDA graph: upload -> kernel -> path extraction -> download ----
    def capture(self, h_path, h_plen, path_cap: int):
        """Record `upload(); run(); paths_device(path_cap)` and the copies of the paths into the PINNED host tensors
        `h_path` [Q, path_cap, 2] / `h_plen` [Q] as one CUDA graph; `replay()` then enqueues the whole call with a single
        launch, so no host work sits between the copies and the kernel.  Inputs are read from the pinned staging tensors
        (h_start_goal, h_obstacles, h_n_obs) at replay time: refill them in place to plan a new scenario."""
        import torch
        self.run()                                   # allocate the result tensors outside the capture
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.upload()
            r = self.run()
            path, plen = r.paths_device(path_cap)
            h_path.copy_(path, non_blocking=True)
            h_plen.copy_(plen, non_blocking=True)
        self._graph = g
        return g

    def replay(self):
        """Enqueue the captured call on the current stream (asynchronous; synchronise before reading the host tensors)."""
        self._graph.replay()

    def planning(self, animation=False):
        """Paths for every query: list of `[[x, y], ...]` (goal -> start) or None.  The reference's near lists have no
        capacity: a query whose list outgrew `near_cap` (status bit RRTK_Q_NEAR_OVERFLOW, the kernel stops that query) is
        planned again on its own with near_cap = 1024, same samples; only if that overflows too the call raises."""
        res = self.run()
        paths = res.paths()
        over = (res.status.cpu().numpy() & _lib.Q_NEAR_OVERFLOW) != 0
        for i in np.flatnonzero(over).tolist():
            c = self._ctor
            if c["near_cap"] >= 1024:
                raise _lib.RrtkError(f"query {i}: near list overflow at near_cap = {c['near_cap']}")
            sub = RRTStarBatch(self.h_start_goal[i:i + 1, 0:2].numpy(), self.h_start_goal[i:i + 1, 2:4].numpy(),
                               self.h_obstacles[i:i + 1].numpy(), c["rand_area"], c["expand_dis"], c["path_resolution"],
                               c["goal_sample_rate"], self.max_iter, c["play_area"], 0.0, c["sampler"], c["connect_circle_dist"],
                               c["search_until_max_iter"], c["seed"], 1024, self.h_n_obs[i:i + 1].numpy(),
                               None if self.sample_stream is None else self.sample_stream[i:i + 1].cpu().numpy(),
                               self.sobol_offset[i:i + 1].cpu().numpy(), self.device, self.params.query_base + i)
            r1 = sub.run()
            if int(r1.status[0].item()) & _lib.Q_NEAR_OVERFLOW:
                raise _lib.RrtkError(f"query {i}: near list overflow at near_cap = 1024")
            paths[i] = r1.paths()[0]
        return paths

    def materialised_stream(self):
        """The samples the in-kernel sampler produces, [Q, max_iter, 2] (for the CPU oracle)."""
        if self.sample_stream is not None:
            return self.sample_stream
        return engine.sample_stream_dev(self.params, self.start_goal, self.sobol_offset)


def shard_range(n_items: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous block partition of `n_items` independent queries over `world_size` ranks."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
