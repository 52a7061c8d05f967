"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI / the reference-shaped Python
classes, against the oracle on identical inputs.

Chain of evidence
  reference  == oracle[libm]   bit for bit on every fixture          (tests/test_oracle_golden.py)
  oracle[cr] == oracle[libm]   bit for bit except logged sub-ulp ties (tests/test_oracle_golden.py)
  GPU        == oracle[cr]     bit for bit, every array, every iteration trace  (this file)
so wherever the cr arithmetic reproduces the fixture, the GPU is also compared with the reference's
own output directly.  Bar: positions, costs, parents, traces and paths are compared with
array_equal -- no tolerance."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_names, load_golden, scenario_args

pytestmark = pytest.mark.gpu

RRT04 = golden_names("rrt04_")
CR_EXACT = [n for n in RRT04 if n != "rrt04_c2_o256_800"]


@pytest.fixture(autouse=True, params=["cta", "warp"])
def exec_mode(request, monkeypatch):
    """Every test of this module runs under both executions of the loop (rrtk_rrtstar_params.exec_mode): one CTA per query
    with the tree in shared memory, and one warp per query.  RRTK_EXEC is read by engine.make_params."""
    monkeypatch.setenv("RRTK_EXEC", request.param)
    return request.param


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need a GPU"
    return torch


def _planner(m):
    import rrtk
    return rrtk.RRTStar(m["start"], m["goal"], m["obstacle_list"], m["rand_area"], m["expand_dis"],
                        m["path_resolution"], m["goal_sample_rate"], m["max_iter"], m["play_area"],
                        m["robot_radius"], m["sobol_sampler"], m["connect_circle_dist"],
                        m["search_until_max_iter"])


def _pad(stream, n):
    stream = np.asarray(stream, dtype=np.float64).reshape(-1, 2)
    if len(stream) < n:
        stream = np.vstack([stream, np.zeros((n - len(stream), 2))])
    return stream


@pytest.mark.parametrize("name", RRT04)
def test_rrtstar_bitwise_vs_oracle_cr(name, torch_cuda, oracle_lib):
    O = oracle_lib
    g, m = load_golden(name)
    p, obs = O.make_params(*scenario_args(m), math_mode=O.MATH_CR)
    ref = O.rrtstar_run(p, obs, g["stream"])
    rrt = _planner(m)
    path = rrt.planning(animation=False, sample_stream=_pad(g["stream"], m["max_iter"]), want_trace=True)
    a = rrt.tree_arrays()
    assert rrt.iters_done == ref["iters_done"]
    assert np.array_equal(rrt.trace, ref["trace"]), "iteration traces differ"
    assert len(a["x"]) == ref["n"]
    assert np.array_equal(a["parent"], ref["parent"])
    assert np.array_equal(a["x"], ref["x"]) and np.array_equal(a["y"], ref["y"])
    assert np.array_equal(a["cost"], ref["cost"])
    assert (rrt.goal_index if rrt.goal_index is not None else -1) == ref["goal_index"]
    assert path == O.final_course(ref, m["goal"])


@pytest.mark.parametrize("name", CR_EXACT)
def test_rrtstar_bitwise_vs_reference_fixture(name, torch_cuda):
    """Directly against what the unmodified reference produced (no oracle in between)."""
    g, m = load_golden(name)
    rrt = _planner(m)
    path = rrt.planning(animation=False, sample_stream=_pad(g["stream"], m["max_iter"]))
    a = rrt.tree_arrays()
    assert rrt.iters_done == m["iters"]
    assert np.array_equal(a["parent"], g["parent"])
    assert np.array_equal(a["x"], g["x"]) and np.array_equal(a["y"], g["y"])
    assert np.array_equal(a["cost"], g["cost"])
    assert np.array_equal(np.array(path, float), g["path"])
    # Node objects as the reference exposes them
    nl = rrt.node_list
    assert nl[0].parent is None and all(n.parent is nl[p] for n, p in zip(nl[1:], a["parent"][1:]))


def test_planning_with_own_sampler_matches_reference(torch_cuda):
    """No injected stream: the planner draws the goal coins from `random` and the Sobol points from the
    GPU generator exactly like the reference did after random.seed(0) (rrt04_c1_sobol_500)."""
    import random
    g, m = load_golden("rrt04_c1_sobol_500")
    random.seed(m["seed"])
    rrt = _planner(m)
    path = rrt.planning(animation=False)
    assert np.array_equal(rrt.sample_stream, g["stream"])
    assert rrt.sobol_inter_ == m["sobol_inter"]
    assert np.array_equal(rrt.tree_arrays()["parent"], g["parent"])
    assert np.array_equal(np.array(path, float), g["path"])
    assert len(path) == 24 and len(rrt.node_list) == 152          # SURVEY.md 8c anchor
    assert rrt.node_list[-1].cost == 10.096284324107767


def _random_scenario(rng, n_obs, max_iter, play=False, rr=0.0, expand=1.0, res=0.1):
    obs = []
    while len(obs) < n_obs:
        x, y = rng.uniform(-2, 15, 2)
        r = rng.uniform(0.1, 0.5)
        if min(np.hypot(x, y), np.hypot(x - 13, y - 13)) > r + rr + 0.5:
            obs.append((float(x), float(y), float(r)))
    goal_rate = 5
    stream = rng.uniform(-2, 15, (max_iter, 2))
    coin = rng.integers(0, 101, max_iter) <= goal_rate
    stream[coin] = (13.0, 13.0)
    return dict(start=[0.0, 0.0], goal=[13.0, 13.0], obstacle_list=obs, rand_area=[-2, 15],
                expand_dis=expand, path_resolution=res, goal_sample_rate=goal_rate, max_iter=max_iter,
                play_area=[-1.0, 14.0, -1.0, 14.0] if play else None, robot_radius=rr,
                connect_circle_dist=50.0, search_until_max_iter=True, sobol_sampler=False), stream


@pytest.mark.parametrize("seed,n_obs,max_iter,play,rr,expand,res", [
    (1, 0, 300, False, 0.0, 1.0, 0.1),      # no obstacles at all
    (2, 40, 500, True, 0.3, 1.0, 0.1),
    (3, 256, 700, False, 0.0, 1.0, 0.1),
    (4, 100, 400, False, 0.2, 3.0, 0.5),    # reference defaults: expand 3.0 / resolution 0.5
    (5, 300, 300, True, 0.1, 2.0, 0.25),    # > CULL_CAP survivors possible -> full-list fallback
    (6, 64, 1500, False, 0.0, 0.5, 0.1),
])
def test_rrtstar_random_scenarios_bitwise(seed, n_obs, max_iter, play, rr, expand, res, torch_cuda, oracle_lib):
    O = oracle_lib
    rng = np.random.default_rng(seed)
    m, stream = _random_scenario(rng, n_obs, max_iter, play, rr, expand, res)
    p, obs = O.make_params(*scenario_args(m), math_mode=O.MATH_CR)
    if n_obs == 0:
        obs = np.zeros((0, 3))
    ref = O.rrtstar_run(p, obs, stream)
    rrt = _planner(m)
    path = rrt.planning(animation=False, sample_stream=stream, want_trace=True)
    a = rrt.tree_arrays()
    assert np.array_equal(rrt.trace, ref["trace"])
    assert np.array_equal(a["parent"], ref["parent"])
    assert np.array_equal(a["x"], ref["x"]) and np.array_equal(a["y"], ref["y"])
    assert np.array_equal(a["cost"], ref["cost"])
    assert path == O.final_course(ref, m["goal"])
    # production path (no trace: rewire edges are evaluated lazily) must give the same tree
    rrt2 = _planner(m)
    path2 = rrt2.planning(animation=False, sample_stream=stream)
    b = rrt2.tree_arrays()
    assert all(np.array_equal(a[k], b[k]) for k in ("x", "y", "cost", "parent")) and path2 == path


def test_batch_matches_single_queries_and_oracle(torch_cuda, oracle_lib):
    """Q queries in one launch (in-kernel Sobol sampler + counter-based coins) == each query alone ==
    oracle[cr] fed the materialised stream; also checks the host restatement of the coins."""
    import rrtk
    from rrtk import sampling
    O = oracle_lib
    Q, iters, n_obs = 24, 400, 96
    rng = np.random.default_rng(7)
    obstacle_lists, starts, goals = [], [], []
    for q in range(Q):
        m, _ = _random_scenario(np.random.default_rng(100 + q), n_obs, iters)
        obstacle_lists.append(m["obstacle_list"])
        starts.append(m["start"]); goals.append(m["goal"])
    b = rrtk.RRTStarBatch(starts, goals, obstacle_lists, [-2, 15], expand_dis=1.0, path_resolution=0.1,
                          goal_sample_rate=5, max_iter=iters, sampler="sobol", seed=99)
    res = b.run(want_trace=True)
    stream = b.materialised_stream().cpu().numpy()
    n_nodes = res.n_nodes.cpu().numpy()
    xy = res.xy.cpu().numpy(); cost = res.cost.cpu().numpy(); parent = res.parent.cpu().numpy()
    trace = res.trace.cpu().numpy(); gi = res.goal_index.cpu().numpy()
    assert (res.status.cpu().numpy() == 0).all()
    paths = res.paths()
    for q in range(Q):
        coins = sampling.kernel_coins(99, q, iters, 5)
        assert np.array_equal(np.all(stream[q] == np.array(goals[q]), axis=1), coins)
        nong = ~coins
        pts = O.sobol_fill(2, q * iters, int(nong.sum()))
        assert np.array_equal(stream[q][nong], -2 + pts * 17)
        p, obs = O.make_params(starts[q], goals[q], obstacle_lists[q], 1.0, 0.1, iters, None, 0.0, 50.0, True,
                               math_mode=O.MATH_CR)
        ref = O.rrtstar_run(p, obs, stream[q])
        n = ref["n"]
        assert n_nodes[q] == n
        assert np.array_equal(trace[q], ref["trace"])
        assert np.array_equal(parent[q, :n], ref["parent"])
        assert np.array_equal(xy[q, :n, 0], ref["x"]) and np.array_equal(xy[q, :n, 1], ref["y"])
        assert np.array_equal(cost[q, :n], ref["cost"])
        assert gi[q] == ref["goal_index"]
        assert paths[q] == O.final_course(ref, goals[q])


def test_rrt_basic_bitwise(torch_cuda, oracle_lib):
    """Basic RRT loop (rrt_01:71-101) against the oracle, with and without a play area."""
    import rrtk
    O = oracle_lib
    for seed, play in ((11, False), (12, True)):
        m, stream = _random_scenario(np.random.default_rng(seed), 60, 600, play, 0.2, 1.0, 0.1)
        p, obs = O.make_params(*scenario_args(m), math_mode=O.MATH_CR)
        ref = O.rrt_run(p, obs, stream)
        rrt = rrtk.RRT(m["start"], m["goal"], m["obstacle_list"], m["rand_area"], m["expand_dis"],
                       m["path_resolution"], m["goal_sample_rate"], m["max_iter"], m["play_area"],
                       m["robot_radius"])
        path = rrt.planning(animation=False, sample_stream=stream)
        a = rrt.tree_arrays()
        assert rrt.iters_done == ref["iters_done"]
        assert np.array_equal(a["parent"], ref["parent"])
        assert np.array_equal(a["x"], ref["x"]) and np.array_equal(a["y"], ref["y"])
        assert (path is None) == (ref["goal_index"] < 0)
        if path is not None:
            assert path == O.final_course(ref, m["goal"])


def test_crmath_on_device(torch_cuda):
    """The device build of csrc/crmath.h against the mpmath fixture (tests/golden/crmath.npz)."""
    torch = torch_cuda
    from rrtk import _lib
    g = np.load(os.path.join(GOLDEN, "crmath.npz"))
    L = _lib.lib()
    s = torch.cuda.current_stream().cuda_stream

    def probe(kind, a, b=None):
        ta = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        tb = None if b is None else torch.from_numpy(np.ascontiguousarray(b)).cuda()
        out = torch.empty_like(ta)
        _lib.check(L.rrtk_crmath_probe_dev(kind, ta.numel(), ta.data_ptr(), None if tb is None else tb.data_ptr(),
                                           out.data_ptr(), s))
        return out.cpu().numpy()
    assert np.array_equal(probe(0, g["y"], g["x"]), g["hypot"])
    assert np.array_equal(probe(1, g["y"], g["x"]), g["atan2"])
    assert np.array_equal(probe(4, g["y"], g["x"]), g["sin_t"])
    assert np.array_equal(probe(5, g["y"], g["x"]), g["cos_t"])
    assert np.array_equal(probe(2, g["arg"]), g["sin"])
    assert np.array_equal(probe(3, g["arg"]), g["cos"])


def test_sobol_device(torch_cuda, oracle_lib):
    from rrtk import sampling
    O = oracle_lib
    for dim, first, count in ((2, 0, 20000), (3, 12345, 5000), (40, 7, 300), (1, 2 ** 29, 64)):
        assert np.array_equal(sampling.sobol_points(dim, first, count), O.sobol_fill(dim, first, count))
    assert sampling.sobol_points(2, 1000, 1)[0].tolist() == [0.2197265625, 0.0966796875]
    assert sampling.sobol_points(2, 0, 0).shape == (0, 2)


def test_host_buffer_entry_point(torch_cuda, oracle_lib):
    """rrtk_rrtstar_run_host: plain host pointers in, host arrays out (the end-to-end C-ABI call)."""
    from rrtk import _lib, engine
    O = oracle_lib
    m, stream = _random_scenario(np.random.default_rng(21), 50, 300)
    rows, n_obs = engine.pack_obstacles([m["obstacle_list"]], 0.0)
    cap = m["max_iter"] + 1
    p = engine.make_params(1, m["max_iter"], cap, rows.shape[1], 1.0, 0.1)
    near = engine.near_r2_table(cap, 50.0, 1.0)
    sg = np.array([[0.0, 0.0, 13.0, 13.0]])
    xy = np.zeros((1, cap, 2)); cost = np.zeros((1, cap)); parent = np.zeros((1, cap), np.int32)
    nn = np.zeros(1, np.int32); itd = np.zeros(1, np.int32); gi = np.zeros(1, np.int32); st = np.zeros(1, np.int32)
    ptr = lambda a: a.ctypes.data  # noqa: E731
    rc = _lib.lib().rrtk_rrtstar_run_host(C.byref(p), ptr(sg), ptr(rows), ptr(n_obs), ptr(near),
                                          ptr(np.ascontiguousarray(stream)), None, ptr(xy), ptr(cost),
                                          ptr(parent), ptr(nn), ptr(itd), ptr(gi), ptr(st), None)
    _lib.check(rc)
    po, obs = O.make_params(*scenario_args(m), math_mode=O.MATH_CR)
    ref = O.rrtstar_run(po, obs, stream)
    n = ref["n"]
    assert nn[0] == n and np.array_equal(parent[0, :n], ref["parent"])
    assert np.array_equal(xy[0, :n, 0], ref["x"]) and np.array_equal(cost[0, :n], ref["cost"])


@pytest.mark.gpu
def test_steer_collide_primitive_matches_the_port():
    """rrtk_steer_collide_dev against the Python port of steer / check_collision / check_if_outside_play_area (which is
    bit-identical to the reference): same point counts and verdicts, new nodes equal up to the last bit of cos / sin."""
    import pyport as P
    from rrtk import engine
    rng = np.random.default_rng(21)
    n = 4000
    f = rng.uniform(-2, 15, (n, 2)); t = f + rng.uniform(-4, 4, (n, 2))
    t[:50] = f[:50]                                              # zero-length edges
    obs = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]
    play = [0.0, 10.0, 0.0, 14.0]
    for ext, res, rr in ((1.0, 0.1, 0.6), (float("inf"), 0.5, 0.0), (3.0, 1.0, 0.0)):
        r = engine.steer_collide(f, t, obs, ext, res, rr, play)
        bad = 0
        for k in range(n):
            ex, ey, px, py = P.steer_points(f[k, 0], f[k, 1], t[k, 0], t[k, 1], ext, res)
            assert r["n_points"][k] == len(px)
            assert abs(r["new_xy"][k, 0] - ex) <= 1e-12 and abs(r["new_xy"][k, 1] - ey) <= 1e-12
            assert r["dist"][k] == math.hypot(t[k, 0] - f[k, 0], t[k, 1] - f[k, 1])
            bad += r["free"][k] != P.collision_free(px, py, obs, rr)
            bad += r["inside"][k] != P.inside_play_area(ex, ey, play)
        assert bad == 0


@pytest.mark.gpu
@pytest.mark.parametrize("sampler", ["sobol", "uniform", "stream"])
def test_incremental_steps_build_the_same_trees(sampler):
    """rrtk_rrtstar_run_dev with resume = 1: 4 steps of 150 iterations == one run of 600, bit for bit."""
    import rrtk
    from rrtk import workloads as W
    cfg = W.C2
    Q, iters = 12, 600
    rows = W.c2_rows(list(range(Q)), 64)
    starts = np.tile(np.array(cfg["start"]), (Q, 1)); goals = np.tile(np.array(cfg["goal"]), (Q, 1))
    stream = None
    if sampler == "stream":
        stream = np.random.default_rng(3).uniform(-2, 15, (Q, iters, 2))
    mk = lambda: rrtk.RRTStarBatch(starts, goals, rows, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"],  # noqa: E731
                                   cfg["goal_sample_rate"], iters, None, cfg["robot_radius"], sampler,
                                   cfg["connect_circle_dist"], True, seed=7, sample_stream=stream)
    a = mk(); ra = a.run()
    b = mk()
    for _ in range(4):
        rb = b.step(150)
    n = ra.n_nodes.cpu().numpy()
    assert np.array_equal(n, rb.n_nodes.cpu().numpy()) and n.min() > 100
    assert (rb.iters_done.cpu().numpy() == iters).all()          # cumulative over the steps
    assert np.array_equal(ra.goal_index.cpu().numpy(), rb.goal_index.cpu().numpy())
    xa, xb = ra.xy.cpu().numpy(), rb.xy.cpu().numpy()
    ca, cb = ra.cost.cpu().numpy(), rb.cost.cpu().numpy()
    pa, pb = ra.parent.cpu().numpy(), rb.parent.cpu().numpy()
    for q in range(Q):
        k = n[q]
        assert np.array_equal(xa[q, :k], xb[q, :k]) and np.array_equal(ca[q, :k], cb[q, :k]) and np.array_equal(pa[q, :k], pb[q, :k])
    assert a.planning() == b.result.paths()


@pytest.mark.gpu
def test_planning_batch_entry_point():
    import rrtk
    obs = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]
    paths = rrtk.RRTStar.planning_batch([[0.0, 0.0]] * 4, [[6.0, 10.0]] * 4, [obs] * 4, [-2, 15], expand_dis=1.0,
                                        path_resolution=0.1, max_iter=400, robot_radius=0.6, sobol_sampler=True,
                                        search_until_max_iter=True, play_area=[0, 10, 0, 14])
    assert len(paths) == 4 and all(p is None or (p[0] == [6.0, 10.0] and p[-1] == [0.0, 0.0]) for p in paths)
    assert any(p is not None for p in paths)


@pytest.mark.gpu
def test_captured_graph_replay_matches_plain_call():
    """RRTStarBatch.capture / replay: the whole call as one CUDA graph gives the plain call's paths, also after the pinned
    inputs are refilled in place."""
    import torch
    import rrtk
    from rrtk import workloads as W
    cfg = W.C2
    Q, iters, cap = 16, 300, 128
    rows = W.c2_rows(list(range(Q)), 64)
    starts = np.tile(np.array(cfg["start"]), (Q, 1)); goals = np.tile(np.array(cfg["goal"]), (Q, 1))
    mk = lambda: rrtk.RRTStarBatch(starts, goals, rows, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"],  # noqa: E731
                                   cfg["goal_sample_rate"], iters, None, cfg["robot_radius"], "sobol",
                                   cfg["connect_circle_dist"], True, seed=3)
    a = mk()
    want_path, want_len = a.run().paths_device(cap)
    b = mk()
    h_path = torch.empty((Q, cap, 2), dtype=torch.float64).pin_memory()
    h_plen = torch.empty((Q,), dtype=torch.int32).pin_memory()
    b.capture(h_path, h_plen, cap)
    for _ in range(2):
        h_plen.zero_()
        b.replay()
        torch.cuda.synchronize()
        assert torch.equal(h_plen, want_len.cpu())
        for q in range(Q):
            k = int(h_plen[q])
            assert torch.equal(h_path[q, :k], want_path[q, :k].cpu())
    # a new scenario through the same graph: swap two queries' obstacle rows in the pinned staging tensor
    b.h_obstacles[[0, 1]] = b.h_obstacles[[1, 0]]
    b.replay()
    torch.cuda.synchronize()
    assert torch.equal(h_plen[2:], want_len.cpu()[2:])


@pytest.mark.gpu
def test_replay_and_upload_follow_new_goals():
    """Refilling h_start_goal in place changes what the NEXT replay() / upload() + run() plans to AND the goal waypoint
    paths_device() writes as path[0] (the result keeps reading the tensor the kernel planned against)."""
    import torch
    import rrtk
    from rrtk import workloads as W
    cfg = W.C2
    Q, iters, cap = 8, 400, 160
    rows = W.c2_rows(list(range(Q)), 32)
    starts = np.tile(np.array(cfg["start"]), (Q, 1))
    goals = np.tile(np.array(cfg["goal"]), (Q, 1))
    goals2 = goals.copy()
    goals2[:, 0] -= 2.5
    goals2[:, 1] -= 1.25
    mk = lambda g: rrtk.RRTStarBatch(starts, g, rows, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"],  # noqa: E731
                                     cfg["goal_sample_rate"], iters, None, cfg["robot_radius"], "sobol",
                                     cfg["connect_circle_dist"], True, seed=5)
    want_path, want_len = mk(goals2).run().paths_device(cap)
    assert int((want_len > 0).sum()) > 0
    # (1) through the captured graph
    b = mk(goals)
    h_path = torch.empty((Q, cap, 2), dtype=torch.float64).pin_memory()
    h_plen = torch.empty((Q,), dtype=torch.int32).pin_memory()
    b.capture(h_path, h_plen, cap)
    b.h_start_goal[:, 2:4] = torch.from_numpy(goals2)
    b.replay()
    torch.cuda.synchronize()
    assert torch.equal(h_plen, want_len.cpu())
    for q in range(Q):
        k = int(h_plen[q])
        assert torch.equal(h_path[q, :k], want_path[q, :k].cpu())
        if k:
            assert h_path[q, 0].tolist() == goals2[q].tolist()
    # (2) plain upload() + run() on an existing result
    c = mk(goals)
    c.run()
    c.h_start_goal[:, 2:4] = torch.from_numpy(goals2)
    c.upload()
    got_path, got_len = c.run().paths_device(cap)
    assert torch.equal(got_len, want_len) and torch.equal(got_path[got_len > 0][:, 0], want_path[want_len > 0][:, 0])


@pytest.mark.gpu
def test_incremental_steps_leave_finished_queries_alone():
    """resume = 1 with search_until_max_iter = False: a query that connected to the goal in an earlier step is not grown
    any further (its tree, goal index and cumulative iteration count stay), so stepping gives the single run's result."""
    import rrtk
    from rrtk import workloads as W
    cfg = W.C2
    Q, iters = 10, 900
    rows = W.c2_rows(list(range(Q)), 24)
    starts = np.tile(np.array(cfg["start"]), (Q, 1))
    goals = np.tile(np.array([6.0, 7.0]), (Q, 1))
    mk = lambda: rrtk.RRTStarBatch(starts, goals, rows, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"],  # noqa: E731
                                   cfg["goal_sample_rate"], iters, None, cfg["robot_radius"], "sobol",
                                   cfg["connect_circle_dist"], False, seed=11)
    ra = mk().run()
    it_a = ra.iters_done.cpu().numpy()
    assert (it_a < iters).any(), "the scene should let some queries finish early"
    b = mk()
    for _ in range(6):
        rb = b.step(150)
    assert np.array_equal(it_a, rb.iters_done.cpu().numpy())
    assert np.array_equal(ra.n_nodes.cpu().numpy(), rb.n_nodes.cpu().numpy())
    assert np.array_equal(ra.goal_index.cpu().numpy(), rb.goal_index.cpu().numpy())
    n = ra.n_nodes.cpu().numpy()
    xa, xb, pa, pb = ra.xy.cpu().numpy(), rb.xy.cpu().numpy(), ra.parent.cpu().numpy(), rb.parent.cpu().numpy()
    for q in range(Q):
        assert np.array_equal(xa[q, :n[q]], xb[q, :n[q]]) and np.array_equal(pa[q, :n[q]], pb[q, :n[q]])


@pytest.mark.gpu
def test_cta_and_warp_executions_agree_on_a_ragged_batch():
    """The two executions of the loop on one batch with ragged obstacle counts, a play area and a near radius that is not
    clipped for the first nodes: identical trees, costs and goal indices."""
    import rrtk
    from rrtk import _lib, engine, workloads as W
    cfg = W.C2
    Q, iters = 40, 700
    rng = np.random.default_rng(21)
    rows = W.c2_rows(list(range(Q)), 48)
    n_obs = rng.integers(0, 49, Q).astype(np.int32)
    starts = np.tile(np.array(cfg["start"]), (Q, 1)); goals = rng.uniform(4.0, 13.0, (Q, 2))
    out = {}
    for mode in ("warp", "cta"):
        os.environ["RRTK_EXEC"] = mode
        b = rrtk.RRTStarBatch(starts, goals, rows, cfg["rand_area"], 1.5, 0.25, cfg["goal_sample_rate"], iters,
                              [-1.0, 14.0, -1.5, 14.5], cfg["robot_radius"], "uniform", 3.0, True, seed=13, n_obs=n_obs)
        assert b.params.exec_mode == {"warp": _lib.EXEC_WARP, "cta": _lib.EXEC_CTA}[mode]
        r = b.run()
        out[mode] = [t.cpu().numpy() for t in (r.n_nodes, r.goal_index, r.status, r.xy, r.cost, r.parent)]
    a, c = out["warp"], out["cta"]
    assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1]) and np.array_equal(a[2], c[2])
    for q in range(Q):
        k = a[0][q]
        assert np.array_equal(a[3][q, :k], c[3][q, :k]) and np.array_equal(a[4][q, :k], c[4][q, :k])
        assert np.array_equal(a[5][q, :k], c[5][q, :k])


def test_batch_planning_replans_near_list_overflows(torch_cuda):
    """The reference's near lists have no capacity.  A batch built with a near_cap its queries outgrow still returns the
    paths of the uncapped run: planning() re-plans the overflowed queries on their own with near_cap = 1024."""
    import rrtk
    from rrtk import _lib, workloads as W
    cfg = W.C2
    Q, iters, n_obs = 6, 900, 64
    rows = W.c2_rows(list(range(Q)), n_obs)
    starts = np.tile(np.array(cfg["start"]), (Q, 1)); goals = np.tile(np.array(cfg["goal"]), (Q, 1))
    mk = lambda cap: rrtk.RRTStarBatch(starts, goals, rows, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"],  # noqa: E731
                                       cfg["goal_sample_rate"], iters, None, cfg["robot_radius"], "sobol",
                                       cfg["connect_circle_dist"], True, seed=11, near_cap=cap)
    small, big = mk(32), mk(256)
    want = big.planning()
    assert int((big.result.status != 0).sum()) == 0
    got = small.planning()
    assert int(((small.result.status & _lib.Q_NEAR_OVERFLOW) != 0).sum()) > 0, "the scenario should overflow a 32-entry list"
    assert got == want
