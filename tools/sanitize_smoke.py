"""Small invocations of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-path-planning_b200"))
import numpy as np, torch
import rrtk
from rrtk import informed, arm as A, smoothing, dubins_planner as DP, workloads as W

g = np.load(os.path.join(ROOT, "tests", "golden", "rrt04_c1_sobol_500.npz")); m = json.loads(str(g["meta"]))
r = rrtk.RRTStar(m["start"], m["goal"], m["obstacle_list"], m["rand_area"], m["expand_dis"], m["path_resolution"],
                 m["goal_sample_rate"], 300, m["play_area"], m["robot_radius"], True, m["connect_circle_dist"], True)
path = r.planning(animation=False, sample_stream=g["stream"][:300])
print("rrtstar nodes", len(r.node_list))
cfg = W.C2
Q, iters, n_obs = 64, 300, 64
qids = list(range(Q)); rows = W.c2_rows(qids, n_obs)
for mode in ("cta", "warp"):      # both executions of the RRT* loop (CTA per query: shared-memory tree + named barriers)
    os.environ["RRTK_EXEC"] = mode
    b = rrtk.RRTStarBatch(np.tile(cfg["start"], (Q, 1)), np.tile(cfg["goal"], (Q, 1)), rows, cfg["rand_area"], cfg["expand_dis"],
                          cfg["path_resolution"], cfg["goal_sample_rate"], iters, None, 0.0, "sobol", cfg["connect_circle_dist"], True, seed=3)
    res = b.run(); sp, sl = res.paths_device(64 + 50)
    print("batch", mode, int(res.n_nodes.sum()))
os.environ.pop("RRTK_EXEC")
st, _ = smoothing.smooth_batch(sp, sl, 50, b.obstacles[:, :, :3].contiguous(), b.n_obs)
print("batch ok", int((res.status != 0).sum()), int((st != 0).sum()))
rng = np.random.default_rng(1)
OBS = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]
it = 1500
free = rng.uniform(-2, 15, (it, 2)); free[rng.integers(0, 101, it) <= 10] = (6.0, 10.0); ball = rng.random((it, 2))
for grid in (0, 3):
    run = informed.run_tree([0.0, 0.0], [6.0, 10.0], OBS, 0.5, it, free, ball, grid=grid)
    print("tree", grid, run.info["n_nodes"], run.info["status"])
out = informed.run_batch([[0.0, 0.0]] * 4, [[6.0, 10.0]] * 4, [OBS] * 4, 0.5, 400, np.tile(free[:400], (4, 1, 1)), np.tile(ball[:400], (4, 1, 1)))
print("informed batch", out[0]["n"])
st3 = np.concatenate([rng.uniform(-2, 15, (4, 150, 2)), rng.uniform(-np.pi, np.pi, (4, 150, 1))], axis=2)
d = DP.run_batch([[0.0, 0.0, 0.0]] * 4, [[10.0, 10.0, 0.0]] * 4, [[(5, 5, 1), (3, 6, 2), (7, 5, 2)]] * 4, 3.0, 150, st3)
print("dubins", d[0]["n"])
sets = np.array([[[1.75, 0.75, 0.6], [0.55, 1.5, 0.5], [0, -1, 0.7], [0, -0.6, 0.4], [-1, 1., 0.3]]])
grids = A.occupancy_grids_device([1.0, 1.0], sets, 64)
routes, rlen, ex = A.astar_torus_batch(grids.clone(), [[5, 30]], [[40, 36]])
print("arm + astar", int(grids.sum()), int(rlen[0]))
# the per-step primitives behind rrtk.RRTStar's methods
nd = r.steer(r.node_list[0], r.Node(3.0, 4.0), 1.0)
print("prims", r.check_collision(nd, r.obstacle_list, 0.0), r.get_nearest_node_index(r.node_list, nd), len(r.find_near_nodes(nd)))
from rrtk import rs_planner as RS
st6 = np.concatenate([rng.uniform(-2, 15, (3, 80, 2)), rng.uniform(-np.pi, np.pi, (3, 80, 1))], axis=2)
rs = RS.run_batch([[0.0, 0.0, 0.0]] * 3, [[6.0, 7.0, 1.57]] * 3, [[(5, 5, 1), (4, 6, 1), (4, 8, 1)]] * 3, 3.0, 80, st6) if hasattr(RS, "run_batch") else None
print("rs", None if rs is None else rs[0]["n"])
torch.cuda.synchronize()
