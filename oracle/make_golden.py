"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/* by running the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py [case ...]

Each fixture is an .npz holding the inputs (scenario parameters, the injected sample
stream) and what the reference produced for them (node positions, costs, parent indices,
path, every `check_collision` verdict in call order).  The reference classes are loaded by
`oracle/ref_loader.py` (definition block only, stub matplotlib) and driven through their own
`planning()`; only the sampler methods are replaced by a function that pops the recorded
stream, and `check_collision` is wrapped to log its verdicts.
"""
from __future__ import annotations

import json
import math
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

C1 = dict(start=[0, 0], goal=[6.0, 10.0],
          obstacle_list=[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2),
                         (8, 10, 1)],
          rand_area=[-2, 15], expand_dis=1.0, path_resolution=0.1, goal_sample_rate=5,
          max_iter=500, play_area=[0, 10, 0, 14], robot_radius=0.6,
          connect_circle_dist=50.0)  # rrt_04:1498-1546


def random_circles(seed, n, lo=-2.0, hi=15.0, rmin=0.1, rmax=0.4, keep_clear=((0.0, 0.0), (13.0, 13.0))):
    """C2-style obstacle set: seeded, rejected if within r+0.5 of start/goal."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        x, y = rng.uniform(lo, hi, 2)
        r = rng.uniform(rmin, rmax)
        if all(np.hypot(x - cx, y - cy) > r + 0.5 for cx, cy in keep_clear):
            out.append((float(x), float(y), float(r)))
    return out


def tree_arrays(node_list):
    idx = {id(n): i for i, n in enumerate(node_list)}
    par = []
    orphans = 0
    for n in node_list:
        if n.parent is None:
            par.append(-1)
        elif isinstance(n.parent, int):
            par.append(n.parent)
        else:
            j = idx.get(id(n.parent), -2)
            orphans += (j == -2)
            par.append(j)
    assert orphans == 0, "reference tree holds a parent object that left node_list"
    x = np.array([float(n.x) for n in node_list])
    y = np.array([float(n.y) for n in node_list])
    c = np.array([float(n.cost) for n in node_list]) if hasattr(node_list[0], "cost") else None
    return x, y, c, np.array(par, dtype=np.int64)


def run_rrt04(name, params, sobol_sampler, search_until_max_iter, seed, stream=None, paths=False):
    """Drive rrt_04's RRT (RRT*).  If `stream` is None the reference samples by itself
    (random.seed(seed) + its Sobol code) and the stream it produced is recorded."""
    ns = ref_loader.load("rrt_04")
    R = ns["RRT"]
    ref_loader.reset_sobol(ns)
    random.seed(seed)
    rrt = R(sobol_sampler=sobol_sampler, search_until_max_iter=search_until_max_iter, **params)
    recorded = []
    verdicts = []

    if stream is None:
        inner = rrt.get_random_node_sobol if sobol_sampler else rrt.get_random_node

        def sampler():
            n = inner()
            recorded.append((float(n.x), float(n.y)))
            return n
    else:
        it = iter(stream)

        def sampler():
            x, y = next(it)
            recorded.append((float(x), float(y)))
            return rrt.Node(float(x), float(y))
    rrt.get_random_node_sobol = sampler
    rrt.get_random_node = sampler

    orig_cc = R.check_collision

    def logged_cc(node, obstacle_list, robot_radius):
        ok = orig_cc(node, obstacle_list, robot_radius)
        verdicts.append(bool(ok))
        return ok
    rrt.check_collision = logged_cc  # instance attribute shadows the staticmethod

    t0 = time.perf_counter()
    with ref_loader.quiet():
        path = rrt.planning(animation=False)
    wall = time.perf_counter() - t0
    x, y, c, par = tree_arrays(rrt.node_list)
    meta = dict(params)
    meta.update(sobol_sampler=bool(sobol_sampler), search_until_max_iter=bool(search_until_max_iter),
                seed=seed, reference_wall_s=wall, iters=len(recorded), kind="rrt_04",
                sobol_inter=int(rrt.sobol_inter_))
    out = dict(meta=json.dumps(meta), stream=np.array(recorded, dtype=np.float64),
               x=x, y=y, cost=c, parent=par,
               verdicts=np.array(verdicts, dtype=np.uint8),
               path=np.array(path, dtype=np.float64) if path is not None else np.zeros((0, 2)))
    if paths:   # every node's path_x / path_y (the sampled edge its last steer call made), concatenated
        lens = [len(n.path_x) for n in rrt.node_list]
        out["path_off"] = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        out["path_xy"] = np.array([[float(a), float(b)] for n in rrt.node_list for a, b in zip(n.path_x, n.path_y)],
                                  dtype=np.float64).reshape(-1, 2)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: {len(x)} nodes, {len(recorded)} iterations, path "
          f"{0 if path is None else len(path)} waypoints, {len(verdicts)} verdicts, "
          f"{wall:.2f} s ({len(recorded) / wall:.1f} tree-iter/s)")


def c2_params(seed, n_obs, max_iter):
    return dict(start=[0.0, 0.0], goal=[13.0, 13.0], obstacle_list=random_circles(seed, n_obs),
                rand_area=[-2, 15], expand_dis=1.0, path_resolution=0.1, goal_sample_rate=5,
                max_iter=max_iter, play_area=None, robot_radius=0.0, connect_circle_dist=50.0)


def informed_rotation(start, goal):
    """The rotation matrix C of rrt_07:1054-1068 (numpy SVD), 2 x 2 block, row major."""
    import math
    c_min = math.hypot(start[0] - goal[0], start[1] - goal[1])
    a1 = np.array([[(goal[0] - start[0]) / c_min], [(goal[1] - start[1]) / c_min], [0]])
    id1_t = np.array([1.0, 0.0, 0.0]).reshape(1, 3)
    m = a1 @ id1_t
    u, s, vh = np.linalg.svd(m, True, True)
    c = u @ np.diag([1.0, 1.0, np.linalg.det(u) * np.linalg.det(np.transpose(vh))]) @ vh
    return [float(c[0, 0]), float(c[0, 1]), float(c[1, 0]), float(c[1, 1])]


def run_rrt07(name, params, seed):
    """Drive rrt_07's RRT.informed_rrt_star_search with injected per-iteration draws: iteration i gets
    free[i] from sample_free_space[_sobol] and ball[i] = (a, b) inside sample_unit_ball."""
    import math
    ns = ref_loader.load("rrt_07")
    R = ns["RRT"]
    rng = np.random.default_rng(seed)
    n = params["max_iter"]
    lo, hi = params["rand_area"]
    free = rng.uniform(lo, hi, (n, 2))
    coin = rng.integers(0, 101, n) <= params["goal_sample_rate"]
    free[coin] = params["goal"]
    ball = rng.random((n, 2))
    rrt = R(sobol_sampler=False, **params)
    state = dict(i=-1)
    orig_informed = rrt.informed_sample

    def informed_sample(c_max, c_min, x_center, c):
        state["i"] += 1
        return orig_informed(c_max, c_min, x_center, c)

    def free_space():
        return [float(free[state["i"], 0]), float(free[state["i"], 1])]

    def unit_ball():
        a, b = float(ball[state["i"], 0]), float(ball[state["i"], 1])
        if b < a:
            a, b = b, a
        sample = (b * math.cos(2 * math.pi * a / b), b * math.sin(2 * math.pi * a / b))
        return np.array([[sample[0]], [sample[1]], [0]])
    rrt.informed_sample = informed_sample
    rrt.sample_free_space = free_space
    rrt.sample_free_space_sobol = free_space
    rrt.sample_unit_ball = unit_ball
    t0 = time.perf_counter()
    with ref_loader.quiet():
        path = rrt.informed_rrt_star_search(animation=False)
    wall = time.perf_counter() - t0
    x, y, c, par = tree_arrays(rrt.node_list)
    meta = dict(params)
    meta.update(kind="rrt_07", seed=seed, reference_wall_s=wall, rot=informed_rotation(params["start"], params["goal"]))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), free=free, ball=ball,
                        x=x, y=y, cost=c, parent=par,
                        path=np.array(path, dtype=np.float64) if path is not None else np.zeros((0, 2)))
    print(f"{name}: {len(x)} nodes, {n} iterations, path {0 if path is None else len(path)} waypoints, "
          f"{wall:.2f} s ({n / wall:.1f} it/s)")


def run_dubins(name, n, seed):
    """plan_dubins_path of the unmodified rrt_05 on seeded random pose pairs (+ the docstring example)."""
    import math
    ns = ref_loader.load("rrt_05")
    f = ns["plan_dubins_path"]
    rng = np.random.default_rng(seed)
    rows = []
    pts_all, off = [], [0]
    for i in range(n):
        if i == 0:
            s, g, kappa = [1.0, 1.0, math.radians(45.0)], [-3.0, -3.0, math.radians(-45.0)], 1.0   # rrt_05:1067-1075
        else:
            s = [float(rng.uniform(-2, 15)), float(rng.uniform(-2, 15)), float(rng.uniform(-math.pi, math.pi))]
            if i % 3 == 0:   # close goals: the CCC words win
                g = [s[0] + float(rng.uniform(-3, 3)), s[1] + float(rng.uniform(-3, 3)), float(rng.uniform(-math.pi, math.pi))]
            else:
                g = [float(rng.uniform(-2, 15)), float(rng.uniform(-2, 15)), float(rng.uniform(-math.pi, math.pi))]
            kappa = float(rng.choice([1.0, 1.0, 0.5, 2.0]))
        x, y, yaw, mode, lengths = f(s[0], s[1], s[2], g[0], g[1], g[2], kappa)
        mi = ["LSL", "RSR", "LSR", "RSL", "RLR", "LRL"].index("".join(mode))
        rows.append(s + g + [kappa, mi] + [float(v) for v in lengths] + [len(x)])
        pts_all.append(np.column_stack([x, y, yaw]))
        off.append(off[-1] + len(x))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(dict(kind="dubins", n=n, seed=seed)),
                        cases=np.array(rows, dtype=np.float64), pts=np.vstack(pts_all), offsets=np.array(off))
    print(f"{name}: {n} pose pairs, {off[-1]} course points, modes "
          f"{np.bincount(np.array(rows)[:, 7].astype(int), minlength=6).tolist()}")


def run_rrt05(name, params, seed, search_until_max_iter=True):
    """Drive rrt_05's RRT.planning (RRT*-Dubins) with an injected (x, y, yaw) sample stream."""
    import math
    ns = ref_loader.load("rrt_05")
    R = ns["RRT"]
    rng = np.random.default_rng(seed)
    n = params["max_iter"]
    lo, hi = params["rand_area"]
    stream = np.column_stack([rng.uniform(lo, hi, (n, 2)), rng.uniform(-math.pi, math.pi, n)])
    coin = rng.integers(0, 101, n) <= params["goal_sample_rate"]
    stream[coin] = params["goal"]
    rrt = R(**params)
    it = iter(stream)
    rrt.get_random_node = lambda: rrt.Node(*[float(v) for v in next(it)])
    sys.setrecursionlimit(100000)          # steer deep-copies the whole ancestor chain (rrt_05:1468)
    t0 = time.perf_counter()
    with ref_loader.quiet():
        path = rrt.planning(animation=False, search_until_max_iter=search_until_max_iter)
    wall = time.perf_counter() - t0
    x, y, c, par = tree_arrays(rrt.node_list)
    yaw = np.array([float(nd.yaw) for nd in rrt.node_list])
    meta = dict(params)
    meta.update(kind="rrt_05", seed=seed, reference_wall_s=wall, search_until_max_iter=search_until_max_iter,
                goal_yaw_th=float(params.get("goal_yaw_th", np.deg2rad(1.0))))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), stream=stream,
                        x=x, y=y, yaw=yaw, cost=c, parent=par,
                        path=np.array(path, dtype=np.float64) if path is not None else np.zeros((0, 2)))
    print(f"{name}: {len(x)} nodes, {n} iterations, path {0 if path is None else len(path)} points, "
          f"{wall:.2f} s ({n / wall:.1f} it/s)")


def run_rrt03(name, params, seed, search_until_max_iter=True, own_sampler=False):
    """Drive rrt_03's RRT.planning (plain RRT + Dubins steering).  own_sampler: the reference samples by itself
    (sobol_sampler per params, `random.seed(seed)`) and the nodes it draws are recorded; otherwise an (x, y, yaw) stream is
    injected through get_random_node.  (rrt_03 builds _PATH_TYPE_MAP before the functions it names: see ref_loader.)"""
    import math
    ns = ref_loader.load("rrt_03")
    ref_loader.reset_sobol(ns)
    R = ns["RRT"]
    n = params["max_iter"]
    rrt = R(**params)
    drawn = []
    if own_sampler:
        random.seed(seed)
        orig = rrt.get_random_node_sobol if params.get("sobol_sampler") else rrt.get_random_node

        def logged():
            nd = orig()
            drawn.append([float(nd.x), float(nd.y), float(nd.yaw)])
            return nd
        rrt.get_random_node_sobol = logged
        rrt.get_random_node = logged
    else:
        rng = np.random.default_rng(seed)
        lo, hi = params["rand_area"]
        stream = np.column_stack([rng.uniform(lo, hi, (n, 2)), rng.uniform(-math.pi, math.pi, n)])
        coin = rng.integers(0, 101, n) <= params["goal_sample_rate"]
        stream[coin] = params["goal"]
        it = iter(stream)

        def pop():
            v = [float(t) for t in next(it)]
            drawn.append(v)
            return rrt.Node(*v)
        rrt.get_random_node = pop
        rrt.get_random_node_sobol = pop
    sys.setrecursionlimit(100000)          # steer deep-copies the whole ancestor chain (rrt_03:1465)
    t0 = time.perf_counter()
    with ref_loader.quiet():
        path = rrt.planning(animation=False, search_until_max_iter=search_until_max_iter)
    wall = time.perf_counter() - t0
    x, y, c, par = tree_arrays(rrt.node_list)
    yaw = np.array([float(nd.yaw) for nd in rrt.node_list])
    stream = np.array(drawn, dtype=np.float64).reshape(-1, 3)
    meta = dict(params)
    meta.update(kind="rrt_03", seed=seed, reference_wall_s=wall, search_until_max_iter=search_until_max_iter,
                own_sampler=own_sampler, sobol_inter_=int(rrt.sobol_inter_), iters=int(stream.shape[0]),
                goal_yaw_th=float(params.get("goal_yaw_th", np.deg2rad(1.0))))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), stream=stream,
                        x=x, y=y, yaw=yaw, cost=c, parent=par,
                        path=np.array(path, dtype=np.float64) if path is not None else np.zeros((0, 2)))
    print(f"{name}: {len(x)} nodes, {stream.shape[0]} iterations, path {0 if path is None else len(path)} points, "
          f"{wall:.2f} s")


def run_rrt06(name, params, seed, search_until_max_iter=True):
    """Drive rrt_06's RRT.planning (RRT*-Reeds-Shepp) with an injected (x, y, yaw) sample stream."""
    ns = ref_loader.load("rrt_06")
    R = ns["RRT"]
    rng = np.random.default_rng(seed)
    n = params["max_iter"]
    lo, hi = params["rand_area"]
    stream = np.column_stack([rng.uniform(lo, hi, (n, 2)), rng.uniform(-math.pi, math.pi, n)])   # no goal bias (:1658)
    rrt = R(**params)
    it = iter(stream)
    rrt.get_random_node = lambda: rrt.Node(*[float(v) for v in next(it)])
    sys.setrecursionlimit(100000)
    t0 = time.perf_counter()
    with ref_loader.quiet():
        path = rrt.planning(animation=False, search_until_max_iter=search_until_max_iter)
    wall = time.perf_counter() - t0
    x, y, c, par = tree_arrays(rrt.node_list)
    yaw = np.array([float(nd.yaw) for nd in rrt.node_list])
    meta = dict(params)
    meta.update(kind="rrt_06", seed=seed, reference_wall_s=wall, search_until_max_iter=search_until_max_iter,
                goal_yaw_th=float(params.get("goal_yaw_th", np.deg2rad(1.0))))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), stream=stream,
                        x=x, y=y, yaw=yaw, cost=c, parent=par,
                        path=np.array(path, dtype=np.float64) if path is not None else np.zeros((0, 3)))
    print(f"{name}: {len(x)} nodes, {n} iterations, path {0 if path is None else len(path)} points, "
          f"{wall:.2f} s ({n / wall:.1f} it/s)")


def run_rrt10(name, params, seed):
    """rrt_10's ClosedLoopRRTStar.planning (:1478-1492) with an injected sample stream: the RRT*-Reeds-Shepp tree with
    Reeds-Shepp-length costs, the goal indexes, and the closed-loop prediction of EVERY candidate course."""
    ns = ref_loader.load("rrt_10")
    rng = np.random.default_rng(seed)
    n = params["max_iter"]
    lo, hi = params["rand_area"]
    stream = np.column_stack([rng.uniform(lo, hi, (n, 2)), rng.uniform(-math.pi, math.pi, n)])
    c = ns["ClosedLoopRRTStar"](**params)
    it = iter(stream)
    c.get_random_node = lambda: c.Node(*[float(v) for v in next(it)])
    sys.setrecursionlimit(100000)
    t0 = time.perf_counter()
    with ref_loader.quiet():
        ns["RRTStarReedsShepp"].planning(c, animation=False)
        wall_plan = time.perf_counter() - t0
        gi = c.get_goal_indexes()
        courses = [c.generate_final_course(i) for i in gi]
        t1 = time.perf_counter()
        res = [c.check_tracking_path_is_feasible(p) for p in courses]
        wall_filter = time.perf_counter() - t1
        best = c.search_best_feasible_path(gi)
    x, y, cst, par = tree_arrays(c.node_list)
    yaw = np.array([float(nd.yaw) for nd in c.node_list])
    traj = [np.array([r[1], r[2], r[3], r[4], r[5], r[6], r[7]], dtype=np.float64).T for r in res]   # x y yaw v t a d
    off = np.cumsum([0] + [len(t) for t in traj])
    coff = np.cumsum([0] + [len(p) for p in courses])
    meta = dict(params)
    meta.update(kind="rrt_10", seed=seed, reference_wall_plan_s=wall_plan, reference_wall_filter_s=wall_filter,
                flag=bool(best[0]))
    win = np.zeros((0, 4)) if not best[0] else np.array([best[1], best[2], best[3]], dtype=np.float64).T
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), stream=stream,
                        x=x, y=y, yaw=yaw, cost=cst, parent=par, goal_idx=np.array(gi, dtype=np.int32),
                        course=np.concatenate([np.array(p, dtype=np.float64) for p in courses]) if courses else np.zeros((0, 3)),
                        course_off=coff, found=np.array([bool(r[0]) for r in res]), traj=np.concatenate(traj) if traj else
                        np.zeros((0, 7)), traj_off=off, winner_xyyaw=win)
    print(f"{name}: {len(x)} nodes, {len(gi)} goal candidates, {int(sum(bool(r[0]) for r in res))} feasible, flag {best[0]}, "
          f"plan {wall_plan:.1f} s, filter {wall_filter:.1f} s")


C10 = dict(start=[0.0, 0.0, 0.0], goal=[6.0, 9.0, float(np.deg2rad(90.0))],
           obstacle_list=[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)],
           rand_area=[-2, 20], max_iter=150, connect_circle_dist=50.0, robot_radius=0.0, target_speed=10.0 / 3.6,
           yaw_th=float(np.deg2rad(3.0)), xy_th=0.5, invalid_travel_ratio=5.0)        # rrt_10:1610-1660


C6 = dict(start=[0.0, 0.0, 0.0], goal=[10.0, 9.0, 0.0],
          obstacle_list=[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)],
          rand_area=[-2, 15], expand_dis=3.0, max_iter=300, robot_radius=0.6, connect_circle_dist=50.0,
          curvature=2.0, goal_xy_th=0.5, step_size=0.1)                      # rrt_06:2015-2083


C5D = dict(start=[0.0, 0.0, 0.0], goal=[10.0, 10.0, 0.0],
           obstacle_list=[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2)],
           rand_area=[-2, 15], expand_dis=3.0, goal_sample_rate=10, max_iter=500, robot_radius=0.0,
           connect_circle_dist=50.0, curvature=1.0, goal_xy_th=0.5)          # rrt_05:1804-1859


C3D = dict(start=[0.0, 0.0, 0.0], goal=[10.0, 10.0, 0.0],
           obstacle_list=[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2)],
           rand_area=[-2, 15], goal_sample_rate=10, max_iter=200, robot_radius=0.6, sobol_sampler=True,
           curvature=1.0, goal_xy_th=0.5)                                     # rrt_03:1664-1716


C7 = dict(start=[0.0, 0.0], goal=[6.0, 10.0],
          obstacle_list=[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)],
          rand_area=[-2, 15], expand_dis=0.5, goal_sample_rate=10, max_iter=200)   # rrt_07:1339-1378


def run_arm02(name, M, link_length, obstacles):
    """get_occupancy_grid of the unmodified arm02 (NLinkArm + detect_collision)."""
    ns = ref_loader.load("arm02")
    arm = ns["NLinkArm"](link_length, [0.0] * len(link_length))
    t0 = time.perf_counter()
    grid = ns["get_occupancy_grid"](arm, obstacles, M)
    wall = time.perf_counter() - t0
    grid = np.asarray(grid)
    meta = dict(kind="arm02", M=M, link_length=list(link_length), obstacles=[list(map(float, o)) for o in obstacles],
                reference_wall_s=wall, occupied=int(grid.sum()), dtype=str(grid.dtype))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta),
                        grid_bits=np.packbits(grid.astype(np.uint8), axis=None), shape=np.array(grid.shape))
    print(f"{name}: M={M}, {int(grid.sum())} occupied cells, {wall:.2f} s ({M * M / wall:.0f} cells/s)")


ARM_OBS = [[1.75, 0.75, 0.6], [0.55, 1.5, 0.5], [0, -1, 0.7], [0, -0.6, 0.4], [-1, 1., 0.3]]  # arm02:298
ARM_LINKS = [0.5, 0.5, 0.3, 0.5, 0.1]                                                            # arm02:302


def _rand_arm_obs(seed, n=5):
    rng = np.random.default_rng(seed)
    return [[float(x), float(y), float(r)] for (x, y), r in
            zip(rng.uniform(-2, 2, (n, 2)), rng.uniform(0.2, 0.7, n))]


class _InjectedRandom:
    """Stands in for the `random` module inside the reference namespace: uniform(a, b) = a + (b - a) * next draw,
    which is CPython's own formula, so the reference consumes a recorded stream."""

    def __init__(self, draws):
        self.it = iter(draws)

    def uniform(self, a, b):
        return a + (b - a) * next(self.it)


class _InjectedRandom2(_InjectedRandom):
    """+ random() for sample_unit_ball (rrt_08:421-430); counts the draws consumed."""

    def __init__(self, draws):
        super().__init__(draws)
        self.used = 0

    def uniform(self, a, b):
        self.used += 1
        return super().uniform(a, b)

    def random(self):
        self.used += 1
        return next(self.it)


def run_rrt08(name, params, seed, n_draws=40000):
    """BITStar.plan (rrt_08:236-331) of the unmodified reference on a recorded stream of unit draws."""
    ns = ref_loader.load("rrt_08")
    draws = np.random.default_rng(seed).random(n_draws)
    inj = _InjectedRandom2(draws.tolist())
    ns["random"] = inj
    b = ns["BITStar"](**params)
    t0 = time.perf_counter()
    err = ""
    with ref_loader.quiet():
        try:
            path = b.plan(animation=False)
        except IndexError as e:                       # both queues ran empty (best_in_vertex_queue on an empty list)
            path, err = [], "IndexError"
    wall = time.perf_counter() - t0
    vid = np.array([float(k) for k in b.tree.vertices.keys()])
    meta = dict(params)
    meta.update(kind="rrt_08", seed=seed, reference_wall_s=wall, draws_used=inj.used, error=err,
                g_goal=float(b.g_scores[b.goalId]), start_id=float(b.startId), goal_id=float(b.goalId))
    np.savez_compressed(
        os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), draws=draws[:inj.used],
        path=np.array([[float(x), float(y)] for x, y in path]) if len(path) else np.zeros((0, 2)),
        vertices=vid, g_vertices=np.array([float(b.g_scores[k]) for k in b.tree.vertices.keys()]),
        edges=np.array([[float(v), float(x)] for v, x in b.tree.edges]).reshape(-1, 2),
        parent_of=np.array([[float(k), float(v)] for k, v in b.nodes.items()]).reshape(-1, 2),
        sample_ids=np.array([float(k) for k in b.samples.keys()]),
        sample_xy=np.array([[float(v[0]), float(v[1])] for v in b.samples.values()]).reshape(-1, 2),
        vertex_queue=np.array([float(v) for v in b.vertex_queue]),
        edge_queue=np.array([[float(v), float(x)] for v, x in b.edge_queue]).reshape(-1, 2))
    print(f"{name}: {len(vid)} vertices, {len(b.tree.edges)} edges, path {len(path)} points, g(goal) {meta['g_goal']:.6f}, "
          f"{inj.used} draws, {wall:.1f} s {err}")


C8 = dict(start=[-1.0, 0.0], goal=[3.0, 8.0],
          obstacleList=[(5, 5, 0.5), (9, 6, 1), (7, 5, 1), (1, 5, 1), (3, 6, 1), (7, 9, 1)],
          randArea=[-2, 15], maxIter=80)                                          # rrt_08:644-679


def run_smoothing(name, src_fixture, max_iter, seed, obstacle_list=None):
    """path_smoothing (rrt_04:1447-1479) of the final course of an existing rrt_04 fixture."""
    ns = ref_loader.load("rrt_04")
    g = np.load(os.path.join(GOLDEN, src_fixture + ".npz"))
    m = json.loads(str(g["meta"]))
    path = [[float(x), float(y)] for x, y in g["path"]]
    obs = [tuple(float(v) for v in o) for o in (obstacle_list if obstacle_list is not None else m["obstacle_list"])]
    rng = np.random.default_rng(seed)
    draws = rng.random((max_iter, 2))
    ns["random"] = _InjectedRandom(draws.ravel().tolist())
    with ref_loader.quiet():
        out = ns["path_smoothing"]([list(pt) for pt in path], max_iter, obs)
    meta = dict(source=src_fixture, max_iter=max_iter, obstacle_list=[list(o) for o in obs])
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), path_in=np.array(path),
                        draws=draws, path_out=np.array(out, dtype=np.float64),
                        length_in=ns["get_path_length"](path), length_out=ns["get_path_length"](out))
    print(f"{name}: {len(path)} -> {len(out)} points, length {ns['get_path_length'](path):.6f} -> "
          f"{ns['get_path_length'](out):.6f}")


class _Dummy:
    """Absorbs any attribute access / call (matplotlib stand-in for astar_torus' drawing loop)."""

    def __getattr__(self, name):
        return self

    def __call__(self, *a, **k):
        return self


def run_astar(name, src_fixture, start, goal):
    """astar_torus (arm02:113-233) of the unmodified reference on the occupancy grid of an arm02 fixture."""
    ns = ref_loader.load("arm02")
    g = np.load(os.path.join(GOLDEN, src_fixture + ".npz"))
    shape = tuple(int(v) for v in g["shape"])
    grid = np.unpackbits(g["grid_bits"])[:shape[0] * shape[1]].reshape(shape).astype(np.int64)
    ns["M"] = shape[0]
    ns["plt"] = _Dummy()
    ns["from_levels_and_colors"] = lambda *a, **k: (None, None)
    work = grid.copy()
    with ref_loader.quiet():
        route = ns["astar_torus"](work, tuple(start), tuple(goal))
        hmap = ns["calc_heuristic_map"](shape[0], tuple(goal))
    meta = dict(source=src_fixture, M=shape[0], start=list(start), goal=list(goal))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta),
                        route=np.array([[int(a), int(b)] for a, b in route], dtype=np.int64).reshape(-1, 2),
                        grid_after=work.astype(np.uint8), heuristic=np.asarray(hmap, dtype=np.int64))
    print(f"{name}: M={shape[0]}, route of {len(route)} cells, {int((work == 2).sum())} expanded")


def run_reeds_shepp(name, n, seed):
    """reeds_shepp_path_planning (rs00:496-515) of the unmodified reference on n random pose pairs."""
    ns = ref_loader.load("rs00")
    rng = np.random.default_rng(seed)
    starts, goals, maxcs, steps = [], [], [], []
    types = np.full((n, 5), -1, dtype=np.int32); lengths = np.zeros((n, 5)); npts = np.zeros(n, dtype=np.int32)
    xs, ys, yaws = [], [], []
    for k in range(n):
        s = [rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-math.pi, math.pi)]
        g = [s[0] + rng.uniform(-6, 6), s[1] + rng.uniform(-6, 6), rng.uniform(-math.pi, math.pi)]
        if k % 7 == 0:
            g = [s[0] + rng.uniform(-0.5, 0.5), s[1] + rng.uniform(-0.5, 0.5), s[2] + rng.uniform(-0.3, 0.3)]
        if k % 11 == 0:
            g = [s[0] + 3.0, s[1], s[2]] if k % 2 else [s[0], s[1] + 2.0, s[2] + math.pi]
        maxc, step = [1.0, 0.5, 2.0, 0.1][k % 4], [0.2, 0.1, 0.05][k % 3]
        with ref_loader.quiet():
            r = ns["reeds_shepp_path_planning"](*s, *g, maxc, step)
        starts.append(s); goals.append(g); maxcs.append(maxc); steps.append(step)
        if r[0] is not None:
            m = len(r[3])
            types[k, :m] = ["LSR".index(c) for c in r[3]]
            lengths[k, :m] = [float(v) for v in r[4]]
            npts[k] = len(r[0])
            xs += [float(v) for v in r[0]]; ys += [float(v) for v in r[1]]; yaws += [float(v) for v in r[2]]
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(dict(kind="rs00", n=n)),
                        start=np.array(starts), goal=np.array(goals), maxc=np.array(maxcs), step=np.array(steps),
                        types=types, lengths=lengths, n_pts=npts, x=np.array(xs), y=np.array(ys), yaw=np.array(yaws))
    print(f"{name}: {n} pairs, {int((npts == 0).sum())} without a path, {len(xs)} course points")


CASES = {
    "rrt08_builtin_80": lambda: run_rrt08("rrt08_builtin_80", C8, 31),
    "rrt08_builtin_200": lambda: run_rrt08("rrt08_builtin_200", dict(C8, maxIter=200), 32),
    "rrt08_dense_120": lambda: run_rrt08("rrt08_dense_120", dict(
        C8, start=[0.0, 0.0], goal=[9.0, 9.0], maxIter=120,
        obstacleList=[(5, 5, 1.0), (3, 6, 1.5), (3, 8, 1.0), (7, 5, 1.5), (6, 8, 1.0), (8, 2, 1.0), (2, 3, 0.8)]), 33),
    "rrt08_far_goal_300": lambda: run_rrt08("rrt08_far_goal_300", dict(
        C8, start=[0.0, 0.0], goal=[14.0, 14.0], maxIter=300, obstacleList=[(5, 5, 1.0), (10, 10, 2.0)]), 100),
    "rrt10_cl_60": lambda: run_rrt10("rrt10_cl_60", dict(C10, max_iter=60), 5),
    "rrt10_cl_builtin_150": lambda: run_rrt10("rrt10_cl_builtin_150", C10, 11),
    "rrt10_cl_radius_100": lambda: run_rrt10("rrt10_cl_radius_100", dict(C10, max_iter=100, robot_radius=0.3,
                                                                         goal=[8.0, 7.5, 0.0]), 23),
    "rrt06_builtin_300": lambda: run_rrt06("rrt06_builtin_300", C6, 1),
    "rrt06_builtin_700": lambda: run_rrt06("rrt06_builtin_700", dict(C6, max_iter=700), 2),
    "rrt06_loose_400": lambda: run_rrt06("rrt06_loose_400", dict(
        C6, max_iter=400, curvature=1.0, step_size=0.2, robot_radius=0.3, goal_yaw_th=float(np.deg2rad(15.0)), goal_xy_th=1.0,
        obstacle_list=[(5, 5, 1), (9, 6, 1), (7, 5, 1), (1, 5, 1), (3, 6, 1), (7, 9, 1)]), 3),
    "rrt06_early_exit_500": lambda: run_rrt06("rrt06_early_exit_500", dict(C6, max_iter=500), 4, False),
    "rs_pairs_400": lambda: run_reeds_shepp("rs_pairs_400", 400, 41),
    "astar_script_m100": lambda: run_astar("astar_script_m100", "arm02_script_m100", (10, 50), (58, 56)),   # arm02:309-310
    "astar_script_m100_b": lambda: run_astar("astar_script_m100_b", "arm02_script_m100", (95, 3), (40, 80)),
    "astar_2link_m100": lambda: run_astar("astar_2link_m100", "arm02_2link_m100", (10, 50), (58, 56)),
    "astar_rand_m51": lambda: run_astar("astar_rand_m51", "arm02_script_m51_rand", (2, 2), (30, 44)),
    "astar_3link_m64_enclosed": lambda: run_astar("astar_3link_m64_enclosed", "arm02_3link_m64_rand", (1, 62), (60, 3)),
    "astar_rand_m51_wrap2": lambda: run_astar("astar_rand_m51_wrap2", "arm02_script_m51_rand", (15, 42), (42, 35)),
    "astar_2link_m100_wrap": lambda: run_astar("astar_2link_m100_wrap", "arm02_2link_m100", (88, 23), (96, 54)),
    "astar_same_cell": lambda: run_astar("astar_same_cell", "arm02_script_m51_rand", (7, 9), (7, 9)),
    "smooth_c1_sobol_1000": lambda: run_smoothing("smooth_c1_sobol_1000", "rrt04_c1_sobol_500", 1000, 31),
    "smooth_c1_sobol2000_300": lambda: run_smoothing("smooth_c1_sobol2000_300", "rrt04_c1_sobol_2000", 300, 32),
    "smooth_c1_uniform_1000": lambda: run_smoothing("smooth_c1_uniform_1000", "rrt04_c1_uniform_500", 1000, 33),
    "smooth_c2_o256_500": lambda: run_smoothing("smooth_c2_o256_500", "rrt04_c2_o256_2000", 500, 34),
    "smooth_c2_o64_free_200": lambda: run_smoothing("smooth_c2_o64_free_200", "rrt04_c2_o64_600", 200, 35, obstacle_list=[]),
    "rrt05_builtin_500": lambda: run_rrt05("rrt05_builtin_500", C5D, 1),
    "rrt05_builtin_1500": lambda: run_rrt05("rrt05_builtin_1500", dict(C5D, max_iter=1500), 2),
    "rrt05_loose_goal_800": lambda: run_rrt05("rrt05_loose_goal_800", dict(
        C5D, max_iter=800, goal_yaw_th=float(np.deg2rad(20.0)), goal_xy_th=1.0, robot_radius=0.3, curvature=1.5,
        obstacle_list=[(5, 5, 1), (9, 6, 1), (7, 5, 1), (1, 5, 1), (3, 6, 1), (7, 9, 1)]), 3),
    "rrt05_early_exit_600": lambda: run_rrt05("rrt05_early_exit_600", dict(
        C5D, max_iter=600, goal_yaw_th=float(np.deg2rad(30.0)), goal_xy_th=1.5), 4, False),
    # rrt_03 (plain RRT + Dubins): the script's own scene with its own 3-D Sobol sampler, then injected streams
    "rrt03_builtin_sobol_200": lambda: run_rrt03("rrt03_builtin_sobol_200", C3D, 0, True, own_sampler=True),
    "rrt03_sobol_early_600": lambda: run_rrt03("rrt03_sobol_early_600", dict(
        C3D, max_iter=600, goal_yaw_th=float(np.deg2rad(6.0)), goal_xy_th=0.9, goal_sample_rate=2), 5, False, own_sampler=True),
    "rrt03_uniform_own_400": lambda: run_rrt03("rrt03_uniform_own_400", dict(C3D, max_iter=400, sobol_sampler=False), 7, True,
                                               own_sampler=True),
    "rrt03_stream_800": lambda: run_rrt03("rrt03_stream_800", dict(
        C3D, max_iter=800, robot_radius=0.2, curvature=1.5, goal_yaw_th=float(np.deg2rad(20.0)), goal_xy_th=1.0,
        obstacle_list=[(5, 5, 1), (9, 6, 1), (7, 5, 1), (1, 5, 1), (3, 6, 1), (7, 9, 1)]), 11),
    "rrt03_play_area_500": lambda: run_rrt03("rrt03_play_area_500", dict(
        C3D, max_iter=500, play_area=[-1.0, 12.0, -1.0, 13.0], goal_sample_rate=-1), 12),   # (no goal samples: a second one
    # from a node already ON the goal pose makes steer return None, which rrt_03 turns into an AttributeError at :1626)
    "dubins_pairs_150": lambda: run_dubins("dubins_pairs_150", 150, 21),
    "rrt07_builtin_200": lambda: run_rrt07("rrt07_builtin_200", C7, 1),
    "rrt07_builtin_1000": lambda: run_rrt07("rrt07_builtin_1000", dict(C7, max_iter=1000), 2),
    "rrt07_builtin_2500": lambda: run_rrt07("rrt07_builtin_2500", dict(C7, max_iter=2500), 3),
    "rrt07_alt_800": lambda: run_rrt07("rrt07_alt_800", dict(
        C7, max_iter=800, start=[1.0, 12.0], goal=[11.5, 1.0], expand_dis=0.8,
        obstacle_list=[(5, 5, 0.5), (9, 6, 1), (7, 5, 1), (1, 5, 1), (3, 6, 1), (7, 9, 1)]), 4),
    "arm02_script_m100": lambda: run_arm02("arm02_script_m100", 100, ARM_LINKS, ARM_OBS),
    "arm02_2link_m100": lambda: run_arm02("arm02_2link_m100", 100, [1.0, 1.0], ARM_OBS),
    "arm02_script_m51_rand": lambda: run_arm02("arm02_script_m51_rand", 51, ARM_LINKS, _rand_arm_obs(5)),
    "arm02_3link_m64_rand": lambda: run_arm02("arm02_3link_m64_rand", 64, [0.8, 0.6, 0.4], _rand_arm_obs(6, 7)),
    # C1: the built-in scenario, the reference samples by itself (Sobol + random.seed(0))
    "rrt04_c1_sobol_500": lambda: run_rrt04("rrt04_c1_sobol_500", C1, True, True, 0),
    "rrt04_c1_sobol_2000": lambda: run_rrt04("rrt04_c1_sobol_2000", dict(C1, max_iter=2000), True, True, 0),
    "rrt04_c1_uniform_500": lambda: run_rrt04("rrt04_c1_uniform_500", C1, False, True, 3),
    # script default: stop at the first goal connection
    "rrt04_c1_sobol_early": lambda: run_rrt04("rrt04_c1_sobol_early", C1, True, False, 1),
    "rrt04_c1_uniform_early": lambda: run_rrt04("rrt04_c1_uniform_early", C1, False, False, 2),
    # C2 shape (random circles, no play area, robot_radius 0)
    "rrt04_c2_o64_600": lambda: run_rrt04("rrt04_c2_o64_600", c2_params(1234, 64, 600), True, True, 5),
    # the same two runs with every node's path_x / path_y kept (rrtk.Node.path_x parity)
    "paths04_c1_sobol_500": lambda: run_rrt04("paths04_c1_sobol_500", C1, True, True, 0, paths=True),
    "paths04_c2_o64_600": lambda: run_rrt04("paths04_c2_o64_600", c2_params(1234, 64, 600), True, True, 5, paths=True),
    "paths04_c1_uniform_early": lambda: run_rrt04("paths04_c1_uniform_early", C1, False, False, 2, paths=True),
    "rrt04_c2_o256_800": lambda: run_rrt04("rrt04_c2_o256_800", c2_params(1235, 256, 800), True, True, 6),
    "rrt04_c2_o256_2000": lambda: run_rrt04("rrt04_c2_o256_2000", c2_params(1236, 256, 2000), True, True, 7),
}


if __name__ == "__main__":
    if not ref_loader.available():
        sys.exit("reference not present at " + ref_loader.REF_DIR)
    os.makedirs(GOLDEN, exist_ok=True)
    want = sys.argv[1:] or list(CASES)
    for name in want:
        if name.endswith("*"):
            for k in CASES:
                if k.startswith(name[:-1]):
                    CASES[k]()
        else:
            CASES[name]()
