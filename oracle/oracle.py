"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the plain-C oracle (oracle/rrtk_oracle.c).

`build()` compiles oracle/_build/liborc.so with gcc (see oracle/Makefile); building the checker
is not using it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liborc.so")
_lib = None


class Params(C.Structure):
    _fields_ = [("sx", C.c_double), ("sy", C.c_double), ("gx", C.c_double), ("gy", C.c_double),
                ("expand_dis", C.c_double), ("res", C.c_double), ("robot_radius", C.c_double),
                ("connect_circle_dist", C.c_double), ("play", C.c_double * 4),
                ("has_play", C.c_int32), ("max_iter", C.c_int32),
                ("search_until_max_iter", C.c_int32), ("n_obs", C.c_int32),
                ("math_mode", C.c_int32), ("pad_", C.c_int32)]

MATH_LIBM = 0   # the reference's arithmetic on this platform (CPython -> glibc)
MATH_CR = 1     # correctly-rounded leaf functions (what the GPU path computes)


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, "rrtk_oracle.c"),
            os.path.join(HERE, "..", "robotics-path-planning_b200", "csrc", "crmath.h"),
            os.path.join(HERE, "..", "robotics-path-planning_b200", "csrc", "crmath_consts.h")]
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        for name, nargs in (("orc_hypot", 2), ("orc_cr_hypot", 2), ("orc_cr_sin", 1),
                            ("orc_cr_cos", 1), ("orc_cr_atan2", 2), ("orc_cr_acos", 1), ("orc_cr_asin", 1), ("orc_cr_tan", 1)):
            f = getattr(_lib, name)
            f.restype = C.c_double
            f.argtypes = [C.c_double] * nargs
        _lib.orc_cr_atan2_sincos.restype = C.c_double
        _lib.orc_cr_atan2_sincos.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double),
                                             C.POINTER(C.c_double)]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def make_params(start, goal, obstacle_list, expand_dis, path_resolution, max_iter, play_area,
                robot_radius, connect_circle_dist=50.0, search_until_max_iter=True,
                math_mode=MATH_LIBM):
    p = Params()
    p.math_mode = int(math_mode)
    p.sx, p.sy, p.gx, p.gy = float(start[0]), float(start[1]), float(goal[0]), float(goal[1])
    p.expand_dis, p.res = float(expand_dis), float(path_resolution)
    p.robot_radius, p.connect_circle_dist = float(robot_radius), float(connect_circle_dist)
    p.has_play = 0 if play_area is None else 1
    if play_area is not None:
        for i in range(4):
            p.play[i] = float(play_area[i])
    p.max_iter = int(max_iter)
    p.search_until_max_iter = int(bool(search_until_max_iter))
    obs = np.ascontiguousarray(np.asarray(obstacle_list, dtype=np.float64).reshape(-1, 3))
    p.n_obs = obs.shape[0]
    return p, obs


def hypot(a: float, b: float) -> float:
    return lib().orc_hypot(a, b)


def sobol_fill(dim: int, first: int, count: int) -> np.ndarray:
    out = np.empty((count, dim), dtype=np.float64)
    rc = lib().orc_sobol_fill(C.c_int(dim), C.c_int64(first), C.c_int64(count), _p(out, C.c_double))
    if rc:
        raise ValueError("orc_sobol_fill: bad dimension")
    return out


def sobol_table(dim: int) -> np.ndarray:
    v = np.empty((dim, 30), dtype=np.uint32)
    if lib().orc_sobol_table(C.c_int(dim), _p(v, C.c_uint32)):
        raise ValueError("orc_sobol_table: bad dimension")
    return v


def rrtstar_run(params: Params, obs: np.ndarray, stream: np.ndarray, want_trace=True,
                verdict_cap=0, tie_cap=0):
    """Returns dict(x, y, cost, parent, n, iters_done, goal_index, trace, verdicts)."""
    cap = params.max_iter + 1
    stream = np.ascontiguousarray(stream, dtype=np.float64).reshape(-1, 2)
    if stream.shape[0] < params.max_iter:
        pad = np.zeros((params.max_iter - stream.shape[0], 2))
        stream = np.ascontiguousarray(np.vstack([stream, pad]))
    x = np.zeros(cap); y = np.zeros(cap); cost = np.zeros(cap)
    parent = np.full(cap, -1, dtype=np.int32)
    n = C.c_int32(); it = C.c_int32(); gi = C.c_int32()
    trace = np.zeros((params.max_iter, 8), dtype=np.int32) if want_trace else None
    verd = np.zeros(max(verdict_cap, 1), dtype=np.uint8)
    nv = C.c_int64()
    ties = np.zeros((max(tie_cap, 1), 3), dtype=np.float64)
    nt = C.c_int64()
    work = np.zeros(2, dtype=np.int64)
    lib().orc_rrtstar_run(C.byref(params), _p(obs, C.c_double), _p(stream, C.c_double),
                          _p(x, C.c_double), _p(y, C.c_double), _p(cost, C.c_double),
                          _p(parent, C.c_int32), C.byref(n), C.byref(it), C.byref(gi),
                          _p(trace, C.c_int32) if want_trace else None,
                          _p(verd, C.c_uint8) if verdict_cap else None, C.c_int64(verdict_cap),
                          C.byref(nv), _p(ties, C.c_double) if tie_cap else None, C.c_int64(tie_cap),
                          C.byref(nt), _p(work, C.c_int64))
    k = n.value
    return dict(x=x[:k], y=y[:k], cost=cost[:k], parent=parent[:k], n=k, iters_done=it.value,
                goal_index=gi.value, trace=None if trace is None else trace[:it.value],
                verdicts=verd[:min(nv.value, verdict_cap)], n_verdicts=nv.value,
                ties=ties[:min(nt.value, tie_cap)], n_ties=nt.value,
                work_pairs=int(work[0]), work_scan=int(work[1]))


def rrt_run(params: Params, obs: np.ndarray, stream: np.ndarray):
    cap = params.max_iter + 1
    stream = np.ascontiguousarray(stream, dtype=np.float64).reshape(-1, 2)
    if stream.shape[0] < params.max_iter:
        stream = np.ascontiguousarray(np.vstack([stream, np.zeros((params.max_iter - stream.shape[0], 2))]))
    x = np.zeros(cap); y = np.zeros(cap)
    parent = np.full(cap, -1, dtype=np.int32)
    n = C.c_int32(); it = C.c_int32(); gi = C.c_int32()
    lib().orc_rrt_run(C.byref(params), _p(obs, C.c_double), _p(stream, C.c_double),
                      _p(x, C.c_double), _p(y, C.c_double), _p(parent, C.c_int32),
                      C.byref(n), C.byref(it), C.byref(gi))
    k = n.value
    return dict(x=x[:k], y=y[:k], parent=parent[:k], n=k, iters_done=it.value, goal_index=gi.value)


def final_course(res: dict, goal) -> list | None:
    """generate_final_course (rrt_04:1117-1125) from an oracle result."""
    gi = res["goal_index"]
    if gi < 0:
        return None
    path = [[float(goal[0]), float(goal[1])]]
    i = gi
    while res["parent"][i] >= 0:
        path.append([float(res["x"][i]), float(res["y"][i])])
        i = int(res["parent"][i])
    path.append([float(res["x"][i]), float(res["y"][i])])
    return path


def cr_atan2_sincos(y: float, x: float):
    s = C.c_double(); c = C.c_double()
    th = lib().orc_cr_atan2_sincos(y, x, C.byref(s), C.byref(c))
    return th, s.value, c.value


def arm_theta_list(M: int) -> np.ndarray:
    """theta_list[:M] of arm02:95: [2 * i * pi / M for i in range(-M // 2, M // 2 + 1)]."""
    from math import pi
    return np.array([2 * i * pi / M for i in range(-M // 2, M // 2 + 1)][:M], dtype=np.float64)


def arm_grid(M, link_lengths, obstacles, math_mode=MATH_LIBM, row0=0, n_rows=None):
    """get_occupancy_grid (arm02:79-110) -> uint8 [n_rows, M]."""
    n_rows = M - row0 if n_rows is None else n_rows
    theta = arm_theta_list(M)
    link = np.ascontiguousarray(link_lengths, dtype=np.float64)
    obs = np.ascontiguousarray(np.asarray(obstacles, dtype=np.float64).reshape(-1, 3))
    grid = np.zeros((n_rows, M), dtype=np.uint8)
    rc = lib().orc_arm_grid(C.c_int32(M), _p(theta, C.c_double), C.c_int32(row0), C.c_int32(n_rows),
                            C.c_int32(len(link)), _p(link, C.c_double), _p(obs, C.c_double),
                            C.c_int32(obs.shape[0]), C.c_int32(math_mode), _p(grid, C.c_uint8))
    if rc:
        raise ValueError("orc_arm_grid: bad arguments")
    return grid


class InformedParams(C.Structure):
    _fields_ = [("sx", C.c_double), ("sy", C.c_double), ("gx", C.c_double), ("gy", C.c_double),
                ("expand_dis", C.c_double), ("rot", C.c_double * 4), ("max_iter", C.c_int32),
                ("n_obs", C.c_int32), ("math_mode", C.c_int32), ("path_cap", C.c_int32)]


def informed_run(start, goal, obstacle_list, expand_dis, max_iter, rot, free, ball, math_mode=MATH_LIBM,
                 path_cap=4096):
    """rrt_07 informed_rrt_star_search with injected draws -> dict(x, y, cost, parent, n, path, c_best)."""
    p = InformedParams()
    p.sx, p.sy, p.gx, p.gy = float(start[0]), float(start[1]), float(goal[0]), float(goal[1])
    p.expand_dis = float(expand_dis)
    for i in range(4):
        p.rot[i] = float(rot[i])
    p.max_iter, p.math_mode, p.path_cap = int(max_iter), int(math_mode), int(path_cap)
    obs4 = np.array([[ox, oy, s, s ** 2] for ox, oy, s in obstacle_list], dtype=np.float64).reshape(-1, 4)
    p.n_obs = obs4.shape[0]
    free = np.ascontiguousarray(free, dtype=np.float64).reshape(-1, 2)
    ball = np.ascontiguousarray(ball, dtype=np.float64).reshape(-1, 2)
    cap = max_iter + 1
    x = np.zeros(cap); y = np.zeros(cap); cost = np.zeros(cap); parent = np.full(cap, -1, np.int32)
    path = np.zeros((path_cap, 2)); n = C.c_int32(); plen = C.c_int32(); cb = C.c_double()
    lib().orc_informed_run(C.byref(p), _p(obs4, C.c_double), _p(free, C.c_double), _p(ball, C.c_double),
                           _p(x, C.c_double), _p(y, C.c_double), _p(cost, C.c_double), _p(parent, C.c_int32),
                           C.byref(n), _p(path, C.c_double), C.byref(plen), C.byref(cb))
    k = n.value
    return dict(x=x[:k], y=y[:k], cost=cost[:k], parent=parent[:k], n=k, c_best=cb.value,
                path=None if plen.value == 0 else path[:plen.value].tolist())


DUBINS_MODES = ("LSL", "RSR", "LSR", "RSL", "RLR", "LRL")


def dubins_plan(s, g, curvature=1.0, step=0.1, math_mode=MATH_LIBM, max_pts=2048):
    """plan_dubins_path (rrt_05:1021-1109) -> dict(mode, lengths [3], pts [n, 3] = x, y, yaw)."""
    L = lib()
    L.orc_dubins_plan.restype = C.c_int
    L.orc_dubins_plan.argtypes = [C.c_double] * 8 + [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                                    C.POINTER(C.c_double), C.c_int32]
    mode = C.c_int32()
    lengths = np.zeros(3)
    pts = np.zeros((max_pts, 3))
    n = L.orc_dubins_plan(float(s[0]), float(s[1]), float(s[2]), float(g[0]), float(g[1]), float(g[2]),
                          float(curvature), float(step), int(math_mode), C.byref(mode), _p(lengths, C.c_double),
                          _p(pts, C.c_double), max_pts)
    if n < 0:
        return None
    return dict(mode=mode.value, lengths=lengths, pts=pts[:min(n, max_pts)], n=n)


class DubinsParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("sx", "sy", "syaw", "gx", "gy", "gyaw", "expand_dis", "robot_radius",
                                          "connect_circle_dist", "kappa", "goal_yaw_th", "goal_xy_th")] + \
               [(n, C.c_int32) for n in ("max_iter", "n_obs", "search_until_max_iter", "math_mode")]


def rrtstar_dubins_run(start, goal, obstacle_list, expand_dis, max_iter, robot_radius, connect_circle_dist,
                       curvature, goal_yaw_th, goal_xy_th, search_until_max_iter, stream3, math_mode=MATH_LIBM):
    """rrt_05 planning() with an injected (x, y, yaw) stream."""
    p = DubinsParams()
    p.sx, p.sy, p.syaw = [float(v) for v in start]
    p.gx, p.gy, p.gyaw = [float(v) for v in goal]
    p.expand_dis, p.robot_radius, p.connect_circle_dist = float(expand_dis), float(robot_radius), float(connect_circle_dist)
    p.kappa, p.goal_yaw_th, p.goal_xy_th = float(curvature), float(goal_yaw_th), float(goal_xy_th)
    p.max_iter, p.search_until_max_iter, p.math_mode = int(max_iter), int(bool(search_until_max_iter)), int(math_mode)
    obs = np.ascontiguousarray(np.asarray(obstacle_list, dtype=np.float64).reshape(-1, 3))
    p.n_obs = obs.shape[0]
    st = np.ascontiguousarray(stream3, dtype=np.float64).reshape(-1, 3)
    cap = max_iter + 1
    x = np.zeros(cap); y = np.zeros(cap); yaw = np.zeros(cap); cost = np.zeros(cap)
    parent = np.full(cap, -1, np.int32)
    ef = np.zeros((cap, 3)); et = np.zeros((cap, 3))
    n = C.c_int32(); it = C.c_int32(); gi = C.c_int32()
    lib().orc_rrtstar_dubins_run(C.byref(p), _p(obs, C.c_double), _p(st, C.c_double), _p(x, C.c_double),
                                 _p(y, C.c_double), _p(yaw, C.c_double), _p(cost, C.c_double), _p(parent, C.c_int32),
                                 _p(ef, C.c_double), _p(et, C.c_double), C.byref(n), C.byref(it), C.byref(gi))
    k = n.value
    res = dict(x=x[:k], y=y[:k], yaw=yaw[:k], cost=cost[:k], parent=parent[:k], edge_from=ef[:k], edge_to=et[:k],
               n=k, iters_done=it.value, goal_index=gi.value)
    res["path"] = dubins_final_course(res, start, goal, curvature, math_mode)
    return res


def rrt_dubins_run(start, goal, obstacle_list, max_iter, robot_radius, curvature, goal_yaw_th, goal_xy_th,
                   search_until_max_iter, stream3, play_area=None, math_mode=MATH_LIBM):
    """rrt_03 planning() (plain RRT with Dubins steering) with an injected (x, y, yaw) stream.  Raises AttributeError
    where the reference does (steer returned None while a play area is set)."""
    p = DubinsParams()
    p.sx, p.sy, p.syaw = [float(v) for v in start]
    p.gx, p.gy, p.gyaw = [float(v) for v in goal]
    p.expand_dis, p.robot_radius, p.connect_circle_dist = 0.0, float(robot_radius), 0.0
    p.kappa, p.goal_yaw_th, p.goal_xy_th = float(curvature), float(goal_yaw_th), float(goal_xy_th)
    p.max_iter, p.search_until_max_iter, p.math_mode = int(max_iter), int(bool(search_until_max_iter)), int(math_mode)
    obs = np.ascontiguousarray(np.asarray(obstacle_list, dtype=np.float64).reshape(-1, 3))
    p.n_obs = obs.shape[0]
    st = np.ascontiguousarray(stream3, dtype=np.float64).reshape(-1, 3)
    play = None if play_area is None else np.ascontiguousarray(play_area, dtype=np.float64)
    cap = max_iter + 1
    x = np.zeros(cap); y = np.zeros(cap); yaw = np.zeros(cap); cost = np.zeros(cap)
    parent = np.full(cap, -1, np.int32)
    ef = np.zeros((cap, 3)); et = np.zeros((cap, 3))
    n = C.c_int32(); it = C.c_int32(); gi = C.c_int32()
    L = lib()
    L.orc_rrt_dubins_run.restype = C.c_int
    rc = L.orc_rrt_dubins_run(C.byref(p), _p(obs, C.c_double), _p(st, C.c_double),
                              None if play is None else _p(play, C.c_double), _p(x, C.c_double), _p(y, C.c_double),
                              _p(yaw, C.c_double), _p(cost, C.c_double), _p(parent, C.c_int32), _p(ef, C.c_double),
                              _p(et, C.c_double), C.byref(n), C.byref(it), C.byref(gi))
    if rc == -2:
        raise AttributeError("'NoneType' object has no attribute 'x'")
    k = n.value
    res = dict(x=x[:k], y=y[:k], yaw=yaw[:k], cost=cost[:k], parent=parent[:k], edge_from=ef[:k], edge_to=et[:k],
               n=k, iters_done=it.value, goal_index=gi.value)
    res["path"] = dubins_final_course(res, start, goal, curvature, math_mode)
    return res


def dubins_final_course(res, start, goal, curvature, math_mode):
    """generate_final_course (rrt_05:1512-1521): reversed course samples of every edge up to the root."""
    gi = res["goal_index"]
    if gi < 0:
        return None
    path = [[float(goal[0]), float(goal[1])]]
    i = gi
    while res["parent"][i] >= 0:
        c = dubins_plan(res["edge_from"][i], res["edge_to"][i], curvature, 0.1, math_mode, max_pts=8192)
        for k in range(c["n"] - 1, -1, -1):
            path.append([float(c["pts"][k, 0]), float(c["pts"][k, 1])])
        i = int(res["parent"][i])
    path.append([float(start[0]), float(start[1])])
    return path


def path_smoothing(path, draws, obstacle_list, cap=None):
    """rrt_04:1447-1479 with injected uniform draws -> (path list, status: 0 ok, 1 = reference raises ZeroDivisionError)."""
    path = np.ascontiguousarray(path, dtype=np.float64).reshape(-1, 2)
    draws = np.ascontiguousarray(draws, dtype=np.float64).reshape(-1, 2)
    obs = np.ascontiguousarray(np.asarray(obstacle_list, dtype=np.float64).reshape(-1, 3))
    cap = cap or (path.shape[0] + draws.shape[0] + 2)
    buf = np.zeros((cap, 2))
    buf[:path.shape[0]] = path
    n = C.c_int32(path.shape[0]); done = C.c_int32()
    L = lib()
    L.orc_path_smoothing.restype = C.c_int
    rc = L.orc_path_smoothing(_p(buf, C.c_double), C.byref(n), cap, _p(draws, C.c_double), draws.shape[0],
                              _p(obs, C.c_double), obs.shape[0], C.byref(done))
    return buf[:n.value].tolist(), rc


def astar_torus(grid, start, goal, cap=None):
    """arm02:113-184 -> (route [n, 2] int32 start -> goal (empty = none), final grid uint8 with the 2..6 marks)."""
    g = np.ascontiguousarray(np.asarray(grid), dtype=np.uint8).copy()
    M = g.shape[0]
    cap = cap or M * M
    route = np.zeros((cap, 2), dtype=np.int32)
    L = lib()
    L.orc_astar_torus.restype = C.c_int
    n = L.orc_astar_torus(_p(g, C.c_uint8), M, int(start[0]), int(start[1]), int(goal[0]), int(goal[1]),
                          _p(route, C.c_int32), cap)
    return route[:max(n, 0)].copy(), g


def astar_heuristic(M, goal):
    out = np.zeros((M, M), dtype=np.int64)
    lib().orc_astar_heuristic(M, int(goal[0]), int(goal[1]), _p(out, C.c_int64))
    return out


def astar_heuristic_closed(M, goal):
    L = lib()
    L.orc_astar_heuristic_closed.restype = C.c_int64
    return np.array([[L.orc_astar_heuristic_closed(M, int(goal[0]), int(goal[1]), i, j) for j in range(M)]
                     for i in range(M)], dtype=np.int64)


RS_TYPES = "LSR"


def reeds_shepp(s, g, maxc, step_size=0.2, math_mode=MATH_LIBM, max_pts=4096):
    """reeds_shepp_path_planning (rs00:496-515) -> dict(types, lengths, L, n_paths, pts [n, 4] = x, y, yaw, direction) or None."""
    L = lib()
    L.orc_reeds_shepp.restype = C.c_int
    L.orc_reeds_shepp.argtypes = [C.c_double] * 8 + [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                                    C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_int32]
    types = np.zeros(5, dtype=np.int32); lengths = np.zeros(5); nseg = C.c_int32(); bl = C.c_double(); npaths = C.c_int32()
    pts = np.zeros((max_pts, 4))
    n = L.orc_reeds_shepp(float(s[0]), float(s[1]), float(s[2]), float(g[0]), float(g[1]), float(g[2]), float(maxc),
                          float(step_size), int(math_mode), _p(types, C.c_int32), _p(lengths, C.c_double), C.byref(nseg),
                          C.byref(bl), C.byref(npaths), _p(pts, C.c_double), max_pts)
    if n == 0:
        return None
    k = nseg.value
    return dict(types=[RS_TYPES[t] for t in types[:k]], lengths=lengths[:k].tolist(), L=bl.value, n_paths=npaths.value,
                pts=pts[:min(n, max_pts)].copy(), n=n)


class RSParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("sx", "sy", "syaw", "gx", "gy", "gyaw", "expand_dis", "robot_radius",
                                          "connect_circle_dist", "kappa", "goal_yaw_th", "goal_xy_th", "step_size")] + \
               [(k, C.c_int32) for k in ("max_iter", "n_obs", "search_until_max_iter", "math_mode", "cost_mode", "pad_")]


def rrtstar_rs_run(start, goal, obstacle_list, expand_dis, max_iter, robot_radius, connect_circle_dist, curvature,
                   goal_yaw_th, goal_xy_th, search_until_max_iter, stream3, step_size=0.2, math_mode=MATH_LIBM, rs_cost=False):
    """rrt_06 planning() with an injected (x, y, yaw) stream -> tree arrays, goal_index and the final course [n, 3].
    rs_cost=True: rrt_10's RRTStarReedsShepp (Reeds-Shepp length costs; pass expand_dis=inf for its unclipped radius)."""
    p = RSParams()
    p.cost_mode = int(bool(rs_cost))
    p.sx, p.sy, p.syaw = [float(v) for v in start]
    p.gx, p.gy, p.gyaw = [float(v) for v in goal]
    p.expand_dis, p.robot_radius, p.connect_circle_dist = float(expand_dis), float(robot_radius), float(connect_circle_dist)
    p.kappa, p.goal_yaw_th, p.goal_xy_th, p.step_size = float(curvature), float(goal_yaw_th), float(goal_xy_th), float(step_size)
    p.max_iter, p.search_until_max_iter, p.math_mode = int(max_iter), int(bool(search_until_max_iter)), int(math_mode)
    obs = np.ascontiguousarray(np.asarray(obstacle_list, dtype=np.float64).reshape(-1, 3))
    p.n_obs = obs.shape[0]
    st = np.ascontiguousarray(stream3, dtype=np.float64).reshape(-1, 3)
    cap = 2 * max_iter + 1
    x = np.zeros(cap); y = np.zeros(cap); yaw = np.zeros(cap); cost = np.zeros(cap)
    parent = np.full(cap, -1, np.int32)
    ef = np.zeros((cap, 3)); et = np.zeros((cap, 3))
    n = C.c_int32(); it = C.c_int32(); gi = C.c_int32()
    lib().orc_rrtstar_rs_run(C.byref(p), _p(obs, C.c_double), _p(st, C.c_double), _p(x, C.c_double), _p(y, C.c_double),
                             _p(yaw, C.c_double), _p(cost, C.c_double), _p(parent, C.c_int32), _p(ef, C.c_double),
                             _p(et, C.c_double), C.byref(n), C.byref(it), C.byref(gi))
    k = n.value
    res = dict(x=x[:k], y=y[:k], yaw=yaw[:k], cost=cost[:k], parent=parent[:k], edge_from=ef[:k], edge_to=et[:k],
               n=k, iters_done=it.value, goal_index=gi.value)
    res["path"] = rs_final_course(res, start, goal, curvature, step_size, math_mode)
    return res


def rs_final_course(res, start, goal, curvature, step_size, math_mode):
    """generate_final_course (rrt_06:1643-1651): reversed (x, y, yaw) course samples of every edge up to the root."""
    gi = res["goal_index"]
    if gi < 0:
        return None
    path = [[float(goal[0]), float(goal[1]), float(goal[2])]]
    i = gi
    while res["parent"][i] >= 0:
        r = reeds_shepp(res["edge_from"][i], res["edge_to"][i], curvature, step_size, math_mode, max_pts=8192)
        path.extend(r["pts"][::-1, 0:3].tolist())
        i = int(res["parent"][i])
    path.append([float(start[0]), float(start[1]), float(start[2])])
    return path


class CLParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("target_speed", "yaw_th", "invalid_travel_ratio", "robot_radius")] + \
               [(k, C.c_int32) for k in ("n_obs", "math_mode", "traj_cap", "pad_")]


def closed_loop(course, obstacle_list, robot_radius=0.0, target_speed=10.0 / 3.6, yaw_th=np.deg2rad(3.0),
                invalid_travel_ratio=5.0, math_mode=MATH_LIBM, traj_cap=2048):
    """check_tracking_path_is_feasible (rrt_10:1521-1559) for one final course given in DRIVING order (start -> goal,
    rows x, y, yaw).  Returns dict(bits, traj [n, 7] = x, y, yaw, v, t, a, d)."""
    c = np.ascontiguousarray(course, dtype=np.float64).reshape(-1, 3)
    obs = np.ascontiguousarray(np.asarray(obstacle_list, dtype=np.float64).reshape(-1, 3))
    p = CLParams()
    p.target_speed, p.yaw_th, p.invalid_travel_ratio, p.robot_radius = float(target_speed), float(yaw_th), \
        float(invalid_travel_ratio), float(robot_radius)
    p.n_obs, p.math_mode, p.traj_cap = obs.shape[0], int(math_mode), int(traj_cap)
    traj = np.zeros((traj_cap, 7))
    n = C.c_int32(); bits = C.c_int32()
    rc = lib().orc_closed_loop(C.byref(p), c.shape[0], _p(c, C.c_double), _p(obs, C.c_double), _p(traj, C.c_double),
                               C.byref(n), C.byref(bits))
    if rc:
        raise RuntimeError("traj_cap too small")
    return dict(bits=bits.value, traj=traj[:n.value].copy())


def closed_loop_best(courses, *a, **k):
    """search_best_feasible_path (rrt_10:1494-1519): position of the winning course (-1 = none) and every result."""
    res = [closed_loop(c, *a, **k) for c in courses]
    best, best_time = -1, float("inf")
    for i, r in enumerate(res):
        if r["bits"] == 0 and best_time >= r["traj"][-1, 4]:
            best_time, best = r["traj"][-1, 4], i
    return best, res


class BitParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("sx", "sy", "gx", "gy", "min_rand", "max_rand", "lower", "resolution", "num_cells")] + \
               [("rot", C.c_double * 4)] + [(k, C.c_int32) for k in ("max_iter", "n_obs", "math_mode", "n_draws")]


def bit_rotation(start, goal):
    """The 2 x 2 block of C in BITStar.setup_planning (rrt_08:199-213), evaluated with numpy like the reference."""
    c_min = math.hypot(start[0] - goal[0], start[1] - goal[1]) / 1.5
    a1 = np.array([[(goal[0] - start[0]) / c_min], [(goal[1] - start[1]) / c_min], [0]])
    m = np.dot(a1, np.array([1.0, 0.0, 0.0]).reshape(1, 3))
    u, _, vh = np.linalg.svd(m, True, True)
    c = np.dot(np.dot(u, np.diag([1.0, 1.0, np.linalg.det(u) * np.linalg.det(np.transpose(vh))])), vh)
    return [float(c[0, 0]), float(c[0, 1]), float(c[1, 0]), float(c[1, 1])]


def bitstar_plan(start, goal, obstacle_list, rand_area, max_iter, draws, math_mode=MATH_LIBM, vcap=None, scap=8192,
                 ecap=1 << 18):
    """BITStar.plan (rrt_08:236-331) on a recorded stream of unit draws.  Returns dict(path, vertices, g_vertices, edges,
    parent_of, sample_ids, sample_xy, vertex_queue, edge_queue, g_goal, draws_used, status)."""
    p = BitParams()
    p.sx, p.sy, p.gx, p.gy = float(start[0]), float(start[1]), float(goal[0]), float(goal[1])
    p.min_rand, p.max_rand = float(rand_area[0]), float(rand_area[1])
    p.lower, p.resolution = float(rand_area[0]), 0.01
    p.num_cells = float(np.ceil((rand_area[1] - rand_area[0]) / 0.01))
    p.rot[:] = bit_rotation(start, goal)
    obs = np.ascontiguousarray(np.asarray(obstacle_list, dtype=np.float64).reshape(-1, 3))
    d = np.ascontiguousarray(draws, dtype=np.float64)
    p.max_iter, p.n_obs, p.math_mode, p.n_draws = int(max_iter), obs.shape[0], int(math_mode), d.size
    vcap = max_iter + 2 if vcap is None else vcap
    pcap = vcap + 2
    vert = np.zeros(vcap); gv = np.zeros(vcap); edges = np.zeros((vcap, 2)); par = np.zeros((vcap, 2))
    sid = np.zeros(scap); sxy = np.zeros((scap, 2)); vq = np.zeros(vcap); eq = np.zeros((ecap, 2)); path = np.zeros((pcap, 2))
    counts = np.zeros(12, np.int32)
    gg = C.c_double()
    f = lib().orc_bitstar_plan
    f.restype = C.c_int
    rc = f(C.byref(p), _p(obs, C.c_double), _p(d, C.c_double), vcap, scap, ecap, pcap, _p(vert, C.c_double),
           _p(gv, C.c_double), _p(edges, C.c_double), _p(par, C.c_double), _p(sid, C.c_double), _p(sxy, C.c_double),
           _p(vq, C.c_double), _p(eq, C.c_double), _p(path, C.c_double), _p(counts, C.c_int32), C.byref(gg))
    nv, ne, npar, ns, nvq, neq, plen, used = [int(v) for v in counts[:8]]
    return dict(status=rc, path=path[:plen].copy(), vertices=vert[:nv].copy(), g_vertices=gv[:nv].copy(),
                edges=edges[:ne].copy(), parent_of=par[:npar].copy(), sample_ids=sid[:ns].copy(), sample_xy=sxy[:ns].copy(),
                vertex_queue=vq[:nvq].copy(), edge_queue=eq[:neq].copy(), g_goal=gg.value, draws_used=used,
                batches=int(counts[8]), resets=int(counts[9]), skipped=int(counts[10]), expansions=int(counts[11]))
