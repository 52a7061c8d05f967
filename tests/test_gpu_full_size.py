"""GPU tests at BASELINE.json's FULL sizes (-m gpu).  The oracle cannot finish these sizes in seconds, so they are
checked through size-independent properties of the domain plus an oracle spot check on a few whole queries:

  config 2 (4096 queries x 256 circles x 2000 iterations, the bench workload, same seeds):
    every tree is a tree (root, parents in range, no cycles), every cost is its parent's cost plus the edge length
    (calc_new_cost, rrt_04:1376-1378; propagate_cost_to_leaves, :1380-1385), no node lies inside a circle
    (check_collision, rrt_04:1219-1232 tests the end point of every accepted edge), no edge is longer than steer +
    snap allow (rrt_04:1086-1115), the extracted course runs goal -> start along parents (rrt_04:1117-1125), the
    launch is deterministic although the work queue hands queries to warps in a different order each time, and
    four whole queries equal the C oracle bit for bit.
  config 5 (8192 x 8192 joint grid x 64 obstacle sets):
    theta_list (arm02:95) of M = 8192 sampled every 8th angle is theta_list of M = 1024 exactly (a power-of-two
    rescale of `2 * i * pi / M`), so grid_8192[:, ::8, ::8] must equal grid_1024 bit for bit; set 0 at M = 512 is
    compared with the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2_full():
    import torch
    import rrtk
    from rrtk import workloads as W
    assert torch.cuda.is_available(), "these tests need a GPU"
    cfg = W.C2
    Q, iters, n_obs = cfg["n_queries"], cfg["max_iter"], cfg["n_obs"]
    qids = list(range(Q))
    rows = W.c2_rows(qids, n_obs)
    starts = np.tile(np.array(cfg["start"]), (Q, 1))
    goals = np.tile(np.array(cfg["goal"]), (Q, 1))
    batch = rrtk.RRTStarBatch(starts, goals, rows, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"],
                              cfg["goal_sample_rate"], iters, None, cfg["robot_radius"], "sobol",
                              cfg["connect_circle_dist"], True, seed=0xC2,
                              sobol_offset=np.asarray(qids, dtype=np.int64) * iters)
    res = batch.run()
    torch.cuda.synchronize()
    return torch, cfg, batch, res, rows


def _valid_mask(torch, res):
    cap = res.parent.shape[1]
    return torch.arange(cap, device=res.parent.device)[None, :] < res.n_nodes[:, None].long()


def test_c2_full_every_query_ran_to_the_end(c2_full):
    torch, cfg, batch, res, _ = c2_full
    assert int((res.status != 0).sum()) == 0
    assert bool((res.iters_done == cfg["max_iter"]).all())
    n = res.n_nodes
    assert int(n.min()) >= 2 and int(n.max()) <= cfg["max_iter"] + 1
    assert int((res.goal_index >= 0).sum()) == cfg["n_queries"]      # every scene of the workload is solvable
    assert bool((res.goal_index < n).all())


def test_c2_full_trees_are_trees_with_consistent_costs(c2_full):
    torch, cfg, batch, res, _ = c2_full
    valid = _valid_mask(torch, res)
    par = res.parent.long()
    assert bool((par[:, 0] == -1).all())
    inner = valid.clone(); inner[:, 0] = False
    assert bool(((par >= 0) & (par < res.n_nodes[:, None].long()))[inner].all())
    # no cycles: pointer doubling reaches the root from every node in ceil(log2(cap)) rounds
    p = torch.where(inner, par, torch.zeros_like(par))
    for _ in range(int(np.ceil(np.log2(par.shape[1]))) + 1):
        p = torch.gather(p, 1, p)
    assert bool((p == 0).all())
    # cost[i] = cost[parent] + |xy[i] - xy[parent]|.  A rewired node whose steer does not snap back onto its old
    # position (d / resolution within an ulp of an integer) keeps the cost of the ORIGINAL distance, as in the
    # reference (rrt_04:1361), so a handful of nodes may differ by up to one resolution step; none by more.
    pc = torch.where(inner, par, torch.zeros_like(par))
    pxy = torch.gather(res.xy, 1, pc[:, :, None].expand(-1, -1, 2))
    d = torch.hypot(res.xy[..., 0] - pxy[..., 0], res.xy[..., 1] - pxy[..., 1])
    err = (res.cost - (torch.gather(res.cost, 1, pc) + d)).abs()[inner]
    assert bool((res.cost[:, 0] == 0).all())
    assert float(err.max()) <= cfg["path_resolution"] + 1e-9
    assert int((err > 1e-9).sum()) <= 1e-3 * err.numel(), int((err > 1e-9).sum())
    # edge length: steer reaches expand_dis and may snap one resolution step further (rrt_04:1103-1111); rewire and
    # choose_parent edges stay within the near radius <= expand_dis; one more step each time a parent was moved
    assert float(d[inner].max()) <= cfg["expand_dis"] + 3 * cfg["path_resolution"] + 1e-9, float(d[inner].max())
    assert float((d[inner] <= cfg["expand_dis"] + cfg["path_resolution"] + 1e-9).double().mean()) >= 0.999
    assert float(res.cost[valid].min()) == 0.0


def test_c2_full_no_node_inside_a_circle(c2_full):
    torch, cfg, batch, res, rows = c2_full
    valid = _valid_mask(torch, res)
    valid[:, 0] = False                                   # the start is never collision-checked
    obs = torch.from_numpy(rows).to(res.xy.device)        # [Q, O, 4] = x, y, r, r^2
    worst = float("inf")
    for q0 in range(0, obs.shape[0], 128):
        xy = res.xy[q0:q0 + 128]
        o = obs[q0:q0 + 128]
        dx = xy[:, :, None, 0] - o[:, None, :, 0]
        dy = xy[:, :, None, 1] - o[:, None, :, 1]
        margin = ((dx * dx + dy * dy) / o[:, None, :, 3]).amin(dim=2)      # > 1 outside every circle
        worst = min(worst, float(margin[valid[q0:q0 + 128]].min()))
    assert worst > 1.0 - 1e-12, worst


def test_c2_full_courses_follow_parents(c2_full):
    torch, cfg, batch, res, _ = c2_full
    path, plen = res.paths_device(path_cap=256)
    path = path.cpu().numpy(); plen = plen.cpu().numpy()
    assert (plen >= 2).all() and (plen <= 256).all()
    xy = res.xy.cpu().numpy(); par = res.parent.cpu().numpy(); gi = res.goal_index.cpu().numpy()
    for q in range(0, cfg["n_queries"], 97):
        want = [list(cfg["goal"])]
        i = gi[q]
        while par[q, i] >= 0:
            want.append(xy[q, i].tolist())
            i = par[q, i]
        want.append(xy[q, i].tolist())
        assert path[q, :plen[q]].tolist() == want
        assert want[-1] == list(cfg["start"])


def test_c2_full_launch_is_deterministic(c2_full):
    torch, cfg, batch, res, _ = c2_full
    first = [t.clone() for t in (res.xy, res.cost, res.parent, res.n_nodes, res.goal_index)]
    valid = _valid_mask(torch, res)
    again = batch.run()
    torch.cuda.synchronize()
    assert torch.equal(first[3], again.n_nodes) and torch.equal(first[4], again.goal_index)
    assert torch.equal(first[2][valid], again.parent[valid])
    assert torch.equal(first[1][valid], again.cost[valid])
    assert torch.equal(first[0][valid], again.xy[valid])


def test_c2_full_spot_queries_equal_the_oracle(c2_full, oracle_lib):
    """96 of the 4096 queries, spread over every wave of the work queue (plus its ends and the first query of the second
    wave of resident warps), replayed in the C oracle (one thread each: ctypes releases the GIL) and compared bit for bit."""
    torch, cfg, batch, res, _ = c2_full
    from concurrent.futures import ThreadPoolExecutor
    from rrtk import workloads as W
    O = oracle_lib
    picks = sorted(set([0, 1777, 2368, 4095] + list(range(11, cfg["n_queries"], 44))))[:96]
    assert len(picks) >= 64
    stream = batch.materialised_stream()[torch.tensor(picks, device=res.xy.device)].cpu().numpy()
    got = {k: t[torch.tensor(picks, device=t.device)].cpu().numpy() for k, t in
           dict(n=res.n_nodes, parent=res.parent, xy=res.xy, cost=res.cost, gi=res.goal_index).items()}

    def replay(j):
        q = picks[j]
        p, obs = O.make_params(cfg["start"], cfg["goal"], W.c2_obstacles(q, cfg["n_obs"]).tolist(), cfg["expand_dis"],
                               cfg["path_resolution"], cfg["max_iter"], None, cfg["robot_radius"],
                               cfg["connect_circle_dist"], True, math_mode=O.MATH_CR)
        return O.rrtstar_run(p, obs, stream[j], want_trace=False)
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 8)) as ex:
        refs = list(ex.map(replay, range(len(picks))))
    for j, ref in enumerate(refs):
        n = ref["n"]
        assert int(got["n"][j]) == n, picks[j]
        assert np.array_equal(got["parent"][j, :n], ref["parent"]), picks[j]
        assert np.array_equal(got["xy"][j, :n, 0], ref["x"]) and np.array_equal(got["xy"][j, :n, 1], ref["y"]), picks[j]
        assert np.array_equal(got["cost"][j, :n], ref["cost"]), picks[j]
        assert int(got["gi"][j]) == ref["goal_index"], picks[j]


def test_c5_full_grid_subsamples_to_the_coarse_grid(oracle_lib):
    import torch
    from rrtk import arm as A
    M, S = 8192, 64
    rng = np.random.default_rng(5)
    sets = np.concatenate([rng.uniform(-2, 2, (S, 5, 2)), rng.uniform(0.2, 0.7, (S, 5, 1))], axis=2)
    sets[0] = [[1.75, 0.75, 0.6], [0.55, 1.5, 0.5], [0, -1, 0.7], [0, -0.6, 0.4], [-1, 1., 0.3]]   # arm02:298
    link = [0.5, 0.5, 0.3, 0.5, 0.1]
    assert np.array_equal(A.theta_list(M)[::8], A.theta_list(M // 8))
    fine = A.occupancy_grids_device(link, sets, M)
    assert fine.shape == (S, M, M) and fine.dtype == torch.uint8
    assert int(fine.max()) == 1
    coarse = A.occupancy_grids_device(link, sets, M // 8)
    assert torch.equal(fine[:, ::8, ::8], coarse)
    # rows computed as 8 shards (the multi-GPU partition of config 5) equal the whole grid
    r0 = 3 * (M // 8)
    shard = A.occupancy_grids_device(link, sets, M, r0, M // 8)
    assert torch.equal(shard, fine[:, r0:r0 + M // 8])
    del shard
    want = oracle_lib.arm_grid(512, link, sets[0], oracle_lib.MATH_CR)
    assert np.array_equal(fine[0, ::16, ::16].cpu().numpy(), want)
    # the occupied fraction is resolution-independent to first order
    f_fine = float(fine[0].float().mean()); f_coarse = float(coarse[0].float().mean())
    assert abs(f_fine - f_coarse) < 5e-3


def test_c4_full_rrtstar_dubins_batch(oracle_lib):
    """Config 4 at its full size (1024 queries x 500 iterations, the bench workload of rrt_05:1804-1859's scene):
    every tree is a tree whose stored edges join the parent's pose to the node's pose (to rounding: rrt_05 steer takes
    the node pose from the last course sample), a second launch gives identical trees, and every 43rd query equals
    the C oracle bit for bit."""
    import math
    import torch
    from rrtk import dubins_planner as DP
    O = oracle_lib
    Q, iters = 1024, 500
    rng = np.random.default_rng(7)
    st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
    coin = rng.integers(0, 101, (Q, iters)) <= 10
    st[coin] = (10.0, 10.0, 0.0)
    scene = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2)]
    start, goal = [0.0, 0.0, 0.0], [10.0, 10.0, 0.0]
    res = DP.run_batch([start] * Q, [goal] * Q, [scene] * Q, 3.0, iters, st)
    again = DP.run_batch([start] * Q, [goal] * Q, [scene] * Q, 3.0, iters, st)
    torch.cuda.synchronize()
    assert len(res) == Q
    for q, (t, u) in enumerate(zip(res, again)):
        n = t["n"]
        assert t["status"] == 0 and t["iters_done"] == iters and 1 <= n <= iters + 1
        par = t["parent"]
        assert par[0] == -1 and (par[1:] >= 0).all() and (par[1:] < n).all()
        anc = np.where(par < 0, 0, par)
        for _ in range(10):
            anc = anc[anc]
        assert (anc == 0).all()                                     # no cycles
        pose = np.column_stack([t["x"], t["y"], t["yaw"]])
        # a node's pose is the last sample of its course (within rounding of the requested pose, rrt_05 steer); a
        # parent re-wired later is replaced by the end of ITS new course, again within rounding
        assert np.abs(t["edge_to"][1:] - pose[1:]).max(initial=0.0) < 1e-9
        assert np.abs(t["edge_from"][1:] - pose[par[1:]]).max(initial=0.0) < 1e-9
        assert (t["cost"][1:] >= t["cost"][par[1:]]).all() and t["cost"][0] == 0.0   # equal: a second node at the goal pose
        for c in scene:                                             # the end pose of every accepted course is tested
            assert ((t["x"][1:] - c[0]) ** 2 + (t["y"][1:] - c[1]) ** 2 > c[2] ** 2 * (1 - 1e-12)).all()
        assert -1 <= t["goal_index"] < n
        for k in ("x", "y", "yaw", "cost", "parent", "edge_from", "edge_to"):
            assert np.array_equal(t[k], u[k]), (q, k)
        assert t["goal_index"] == u["goal_index"]
    for q in range(0, Q, 43):
        ref = O.rrtstar_dubins_run(start, goal, scene, 3.0, iters, 0.0, 50.0, 1.0, np.deg2rad(1.0), 0.5, True, st[q],
                                   O.MATH_CR)
        t = res[q]
        assert t["n"] == ref["n"] and t["goal_index"] == ref["goal_index"], q
        assert np.array_equal(t["parent"], ref["parent"]), q
        for k in ("x", "y", "yaw", "cost"):
            assert np.array_equal(t[k], ref[k]), (q, k)


def test_c3_full_single_tree_to_a_million_nodes():
    """Config 3 at its full size: ONE Informed RRT* tree (rrt_07 semantics, the script's 7 circles) grown to 10^6
    nodes, the bench workload.  The oracle needs O(n^2) for this, so the tree is checked through what every rrt_07
    tree satisfies: parents in range and acyclic, costs that only go stale upwards (rewire, rrt_07:1145-1168, does not
    propagate), every edge clear of every circle (segment test, rrt_07:1216-1260), and a returned course whose
    length is c_best."""
    import torch
    from rrtk import informed as INF
    cap, iters = 1_000_001, 1_700_000
    rng = np.random.default_rng(9)
    free = rng.uniform(-2, 15, (iters, 2)); coin = rng.integers(0, 101, iters) <= 10; free[coin] = (6.0, 10.0)
    ball = rng.random((iters, 2))
    obs = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]
    run = INF.run_tree([0.0, 0.0], [6.0, 10.0], obs, 0.5, iters, free, ball, node_cap=cap)
    torch.cuda.synchronize()
    i = run.info
    n = i["n_nodes"]
    assert n == cap and i["iters_done"] <= iters
    par = run.parent[:n].long(); xy = run.xy[:n]; cost = run.cost[:n]
    assert int(par[0]) == -1 and bool(((par[1:] >= 0) & (par[1:] < n)).all())
    anc = par.clamp(min=0)
    for _ in range(21):                                    # 2^21 > 10^6: pointer doubling reaches the root
        anc = anc[anc]
    assert bool((anc == 0).all())
    p = par[1:]
    w = xy[1:] - xy[p]
    d = torch.hypot(w[:, 0], w[:, 1])
    assert float(cost[0]) == 0.0 and bool((cost[1:] >= cost[p] + d - 1e-9).all())
    # (no bound on the edge length: rrt_07's near radius 50 * sqrt(log(n) / n) is not capped by expand_dis,
    # rrt_07:1137-1143, so choose_parent / rewire edges of the early tree are long)
    l2 = (w * w).sum(dim=1).clamp(min=1e-300)
    for ox, oy, r in obs:
        c = torch.tensor([ox, oy], dtype=torch.float64, device=xy.device)
        t = (((c - xy[p]) * w).sum(dim=1) / l2).clamp(0.0, 1.0)
        dd = ((c - xy[p] - t[:, None] * w) ** 2).sum(dim=1)
        assert bool((dd > r * r - 1e-9).all())
    a = run.arrays()
    path = np.array(a["path"])
    assert np.isclose(np.hypot(*(path[1:] - path[:-1]).T).sum(), a["c_best"], rtol=1e-12)
    assert np.hypot(6.0, 10.0) < a["c_best"] < 17.0        # rrt_07's scenario: optimum ~16.9
