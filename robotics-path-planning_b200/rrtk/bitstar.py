"""BIT* (10_path_planning_01_rrt_08_batch_informed_rrt_star.py) behind the reference's API.

`BITStar` is rrt_08:138-611: same constructor keywords, `plan(animation=True)` returns the path start -> goal as a list of
`[x, y]` (the interior points are the 0.01-grid coordinates of the tree vertices, like the reference's), or `[]`.
`run_batch` plans Q queries in one launch of rrtk_bitstar_run_dev (one warp per query).  No CPU fallback."""
from __future__ import annotations

import ctypes as C
import math
import random

import numpy as np

from . import _lib, engine

_DEAD_ID = -1.0e300


def rotation(start, goal):
    """The 2 x 2 block of C in setup_planning (rrt_08:199-213), with numpy like the reference."""
    c_min = math.hypot(start[0] - goal[0], start[1] - goal[1]) / 1.5
    a1 = np.array([[(goal[0] - start[0]) / c_min], [(goal[1] - start[1]) / c_min], [0]])
    m = np.dot(a1, np.array([1.0, 0.0, 0.0]).reshape(1, 3))
    u, _, vh = np.linalg.svd(m, True, True)
    c = np.dot(np.dot(u, np.diag([1.0, 1.0, np.linalg.det(u) * np.linalg.det(np.transpose(vh))])), vh)
    return [float(c[0, 0]), float(c[0, 1]), float(c[1, 0]), float(c[1, 1])]


def run_batch(starts, goals, obstacle_lists, rand_area, max_iter, draws, sample_cap=2048, edge_cap=16384, device=None,
              timing=None, inspect=False):
    """Q BIT* queries in one launch.  starts / goals [Q, 2]; draws [Q, n_draws] unit draws (each batch of m samples consumes
    2 * (m + 1)).  Returns a list of dicts: path ([n, 2] start -> goal, empty when none), g_goal, status, counters; with
    inspect=True also the ordered containers of the run (vertices, g_vertices, edges, parent_of, samples, queues)."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda" if device is None else device)
    starts = np.asarray(starts, dtype=np.float64).reshape(-1, 2)
    goals = np.asarray(goals, dtype=np.float64).reshape(-1, 2)
    q = starts.shape[0]
    draws = np.ascontiguousarray(draws, dtype=np.float64).reshape(q, -1)
    rows, counts = engine.pack_obstacles(obstacle_lists, 0.0)
    vcap = int(max_iter) + 2
    p = _lib.BitStarParams()
    p.n_queries, p.max_iter, p.vertex_cap, p.sample_cap, p.edge_cap = q, int(max_iter), vcap, int(sample_cap), int(edge_cap)
    p.path_cap, p.obs_stride, p.n_draws = vcap + 2, rows.shape[1], draws.shape[1]
    p.min_rand, p.max_rand = float(rand_area[0]), float(rand_area[1])
    p.num_cells = float(np.ceil((rand_area[1] - rand_area[0]) / 0.01))            # RTree.__init__ (:50-52)
    rot = np.array([rotation(s, g) for s, g in zip(starts, goals)], dtype=np.float64)
    nd = 6 * sample_cap + 7 * (vcap + 2) + 3 * edge_cap                           # RRTK_BITSTAR_WS_DOUBLES
    ni = 11 * (vcap + 2) + edge_cap                                               # RRTK_BITSTAR_WS_INTS
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    with torch.cuda.device(dev):
        d_sg, d_rot, d_obs, d_cnt, d_draws = t(np.hstack([starts, goals])), t(rot), t(rows), t(counts), t(draws)
        ws_d = torch.empty((q, nd), dtype=torch.float64, device=dev)
        ws_i = torch.empty((q, ni), dtype=torch.int32, device=dev)
        path = torch.empty((q, vcap + 2, 2), dtype=torch.float64, device=dev)
        cnt = torch.empty((q, 12), dtype=torch.int32, device=dev)
        gg = torch.empty(q, dtype=torch.float64, device=dev)
        status = torch.empty(q, dtype=torch.int32, device=dev)
        if timing is not None:
            ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev[0].record()
        rc = _lib.lib().rrtk_bitstar_run_dev(C.byref(p), d_sg.data_ptr(), d_rot.data_ptr(), d_obs.data_ptr(), d_cnt.data_ptr(),
                                             d_draws.data_ptr(), ws_d.data_ptr(), ws_i.data_ptr(), path.data_ptr(),
                                             cnt.data_ptr(), gg.data_ptr(), status.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "rrtk_bitstar_run_dev")
        if timing is not None:
            ev[1].record()
            ev[1].synchronize()
            timing["kernel_ms"] = ev[0].elapsed_time(ev[1])
        h_path, h_cnt, h_gg, h_st = path.cpu().numpy(), cnt.cpu().numpy(), gg.cpu().numpy(), status.cpu().numpy()
        h_d, h_i = (ws_d.cpu().numpy(), ws_i.cpu().numpy()) if inspect else (None, None)
    names = ("n_vertices", "n_edges", "n_parent", "n_sample_slots", "n_vq", "n_eq", "path_len", "draws_used", "batches",
             "resets", "skipped", "expansions")
    out = []
    for i in range(q):
        r = dict(zip(names, (int(v) for v in h_cnt[i])))
        r.update(status=int(h_st[i]), g_goal=float(h_gg[i]), path=h_path[i, :r["path_len"]].copy())
        if inspect:
            r.update(_decode(h_d[i], h_i[i], r, vcap, sample_cap, edge_cap))
        out.append(r)
    return out


def _decode(d, w, r, vcap, scap, ecap):
    """The ordered containers of one run from its workspace (layout of csrc/rrtk_bitstar.cu)."""
    k = vcap + 2
    s_id, s_x, s_y = d[0:scap], d[scap:2 * scap], d[2 * scap:3 * scap]
    o = 3 * scap
    k_id, k_g, k_par = d[o:o + k], d[o + k:o + 2 * k], d[o + 3 * k:o + 4 * k]
    eq_x = d[o + 4 * k:o + 4 * k + ecap]
    par_order, tv, te_v, te_x, vq = (w[j * k:(j + 1) * k] for j in (1, 2, 3, 4, 5))
    eq_v = w[8 * k:8 * k + ecap]
    alive = s_id[:r["n_sample_slots"]] != _DEAD_ID
    nv, ne, npar, nvq, neq = r["n_vertices"], r["n_edges"], r["n_parent"], r["n_vq"], r["n_eq"]
    return dict(
        vertices=k_id[tv[:nv]].copy(), g_vertices=k_g[tv[:nv]].copy(),
        edges=np.stack([k_id[te_v[:ne]], k_id[te_x[:ne]]], 1) if ne else np.zeros((0, 2)),
        parent_of=np.stack([k_id[par_order[:npar]], k_par[par_order[:npar]]], 1) if npar else np.zeros((0, 2)),
        sample_ids=s_id[:r["n_sample_slots"]][alive].copy(),
        sample_xy=np.stack([s_x[:r["n_sample_slots"]][alive], s_y[:r["n_sample_slots"]][alive]], 1),
        vertex_queue=k_id[vq[:nvq]].copy(),
        edge_queue=np.stack([k_id[eq_v[:neq]], eq_x[:neq]], 1) if neq else np.zeros((0, 2)))


class BITStar:
    """rrt_08's `BITStar` (:138-184): same constructor keywords (eta, lowerLimit, upperLimit and resolution are accepted
    and, like in the reference, unused: the grid is randArea at 0.01)."""

    def __init__(self, start, goal, obstacleList, randArea, eta=2.0, maxIter=80, lowerLimit=None, upperLimit=None,
                 resolution=0.01, n_draws=None):
        self.start, self.goal = start, goal
        self.min_rand, self.max_rand = randArea[0], randArea[1]
        self.max_iIter = maxIter
        self.obstacleList = obstacleList
        self.eta = eta
        self.n_draws = n_draws
        self.result = None

    def plan(self, animation=True, draws=None):
        """`draws`: the unit draws to consume instead of Python's `random` (tests).  Without it the stream is taken from
        `random.random()` draw for draw like the reference (`uniform(a, b)` = a + (b - a) * random()); a run that needs
        more draws, samples or queue slots than provisioned is repeated from the start with more (same prefix of draws,
        so the same result)."""
        own = draws is None
        if own:
            n = self.n_draws if self.n_draws is not None else 402 * 8
            draws = [random.random() for _ in range(n)]
        draws = list(np.asarray(draws, dtype=np.float64))
        sample_cap, edge_cap = 2048, 16384
        while True:
            r = run_batch([self.start], [self.goal], [list(self.obstacleList)], [self.min_rand, self.max_rand], self.max_iIter,
                          np.asarray(draws, dtype=np.float64)[None], sample_cap=sample_cap, edge_cap=edge_cap)[0]
            st = r["status"]
            if own and st & _lib.BIT_DRAWS_EXHAUSTED:
                draws += [random.random() for _ in range(len(draws))]
            elif st & _lib.BIT_SAMPLE_OVERFLOW and sample_cap < (1 << 20):
                sample_cap *= 4
            elif st & _lib.BIT_EDGE_OVERFLOW and edge_cap < (1 << 24):
                edge_cap *= 4
            else:
                break
        self.result = r
        if st & _lib.BIT_INDEX_ERROR:
            raise IndexError("list index out of range")          # what the reference raises (best_in_vertex_queue :465)
        if st & _lib.BIT_LIVELOCK:
            raise _lib.RrtkError("every edge of the first batch is blocked: the reference planner never returns here")
        if st:
            raise _lib.RrtkError(f"BIT* run failed (status {st}: draws exhausted = 8, capacity = 1 / 2 / 4 / 32)")
        return [[float(x), float(y)] for x, y in r["path"]]
