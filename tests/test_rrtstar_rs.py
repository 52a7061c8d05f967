"""RRT*-Reeds-Shepp planning loop (rrt_06:1444-1913): oracle ports against fixtures made by the unmodified reference (CPU),
the CUDA kernel bit-exact against oracle[cr] and topology-exact against the reference (GPU)."""
import sys

import numpy as np
import pytest

from conftest import golden_names, load_golden

RRT06 = golden_names("rrt06_")


def _args(m):
    return (m["start"], m["goal"], m["obstacle_list"], m["expand_dis"], m["max_iter"], m["robot_radius"],
            m["connect_circle_dist"], m["curvature"], m["goal_yaw_th"], m["goal_xy_th"], m["search_until_max_iter"])


@pytest.mark.parametrize("name", [n for n in RRT06 if n != "rrt06_builtin_700"])
def test_python_port_matches_reference(name):
    import pyport
    sys.setrecursionlimit(100000)
    g, m = load_golden(name)
    port = pyport.RRTStarRSPort(*_args(m), step_size=m["step_size"])
    path = port.planning([tuple(r) for r in g["stream"]])
    assert np.array_equal(np.array(port.x), g["x"]) and np.array_equal(np.array(port.yaw), g["yaw"])
    assert np.array_equal(np.array(port.cost), g["cost"]) and np.array_equal(np.array(port.parent), g["parent"])
    assert (path is None and len(g["path"]) == 0) or np.array_equal(np.array(path), g["path"])


@pytest.mark.parametrize("name", RRT06)
def test_c_oracle_libm_matches_reference_bitwise(name, oracle_lib):
    O = oracle_lib
    g, m = load_golden(name)
    r = O.rrtstar_rs_run(*_args(m), g["stream"], m["step_size"], O.MATH_LIBM)
    assert r["n"] == len(g["x"]) and np.array_equal(r["parent"], g["parent"])
    for k in ("x", "y", "yaw", "cost"):
        assert np.array_equal(r[k], g[k]), k
    assert (r["path"] is None and len(g["path"]) == 0) or np.array_equal(np.array(r["path"]), g["path"])


@pytest.mark.parametrize("name", RRT06)
def test_c_oracle_cr_mode_same_topology(name, oracle_lib):
    O = oracle_lib
    g, m = load_golden(name)
    r = O.rrtstar_rs_run(*_args(m), g["stream"], m["step_size"], O.MATH_CR)
    assert r["n"] == len(g["x"]) and np.array_equal(r["parent"], g["parent"])
    for k in ("x", "y", "yaw", "cost"):
        assert np.allclose(r[k], g[k], rtol=0, atol=1e-9), k


def test_fixtures_contain_goal_connections():
    """try_goal_path nodes (pose == goal) exist, so the extra append and its Reeds-Shepp cost are exercised."""
    hits = 0
    for name in RRT06:
        g, m = load_golden(name)
        hits += int(((np.abs(g["x"] - m["goal"][0]) < 1e-6) & (np.abs(g["y"] - m["goal"][1]) < 1e-6)).sum())
    assert hits >= 3


# ---------------------------------------------------------------------------------------------- GPU
def _planner(m):
    import rrtk
    return rrtk.RRTStarReedsShepp(m["start"], m["goal"], m["obstacle_list"], m["rand_area"], m["expand_dis"],
                                  max_iter=m["max_iter"], robot_radius=m["robot_radius"],
                                  connect_circle_dist=m["connect_circle_dist"], curvature=m["curvature"],
                                  goal_yaw_th=m["goal_yaw_th"], goal_xy_th=m["goal_xy_th"], step_size=m["step_size"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", RRT06)
def test_gpu_bitwise_vs_oracle_cr(name, oracle_lib):
    O = oracle_lib
    g, m = load_golden(name)
    ref = O.rrtstar_rs_run(*_args(m), g["stream"], m["step_size"], O.MATH_CR)
    r = _planner(m)
    path = r.planning(animation=False, search_until_max_iter=m["search_until_max_iter"], sample_stream=g["stream"])
    t = r.tree_arrays()
    assert t["status"] == 0 and t["n"] == ref["n"] and t["iters_done"] == ref["iters_done"] and t["goal_index"] == ref["goal_index"]
    assert np.array_equal(t["parent"], ref["parent"])
    for k in ("x", "y", "yaw", "cost"):
        assert np.array_equal(t[k], ref[k]), k
    assert np.array_equal(t["edge_from"][1:], ref["edge_from"][1:]) and np.array_equal(t["edge_to"][1:], ref["edge_to"][1:])
    assert (path is None) == (ref["path"] is None)
    if path is not None:
        assert path == ref["path"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", RRT06)
def test_gpu_vs_reference_fixture(name):
    """Against the unmodified reference: identical node count, parents and path length; poses, costs and path points
    within 1e-9 (ulp-level libm differences in the Reeds-Shepp words; north_star allows 1e-5 relative)."""
    g, m = load_golden(name)
    r = _planner(m)
    path = r.planning(animation=False, search_until_max_iter=m["search_until_max_iter"], sample_stream=g["stream"])
    t = r.tree_arrays()
    assert t["n"] == len(g["x"]) and np.array_equal(t["parent"], g["parent"])
    for k in ("x", "y", "yaw", "cost"):
        assert np.allclose(t[k], g[k], rtol=0, atol=1e-9), k
    if len(g["path"]) == 0:
        assert path is None
    else:
        assert len(path) == len(g["path"]) and np.allclose(np.array(path), g["path"], rtol=0, atol=1e-9)
        nd = r.node_list[t["goal_index"]]
        assert len(nd.path_x) == len(nd.path_y) == len(nd.path_yaw) > 1 and nd.parent is not None


@pytest.mark.gpu
def test_gpu_batch_of_queries_matches_oracle(oracle_lib):
    from rrtk import rs_planner as RP
    O = oracle_lib
    Q, iters = 6, 150
    rng = np.random.default_rng(21)
    obs = [(5, 5, 1), (3, 6, 2), (7, 5, 2), (9, 5, 2)]
    streams = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-np.pi, np.pi, (Q, iters, 1))], axis=2)
    starts = [[0.0, 0.0, 0.0]] * Q
    goals = [[10.0, 9.0, 0.0], [8.0, 12.0, 1.0], [12.0, 2.0, -0.5], [10.0, 9.0, 3.0], [2.0, 12.0, 0.0], [13.0, 13.0, 1.5]]
    out = RP.run_batch(starts, goals, [obs] * Q, 3.0, iters, streams, robot_radius=0.4, curvature=1.5, step_size=0.15,
                       goal_yaw_th=np.deg2rad(10.0), goal_xy_th=0.8)
    for q in range(Q):
        ref = O.rrtstar_rs_run(starts[q], goals[q], obs, 3.0, iters, 0.4, 50.0, 1.5, np.deg2rad(10.0), 0.8, True, streams[q],
                               0.15, O.MATH_CR)
        t = out[q]
        assert t["n"] == ref["n"] and np.array_equal(t["parent"], ref["parent"]) and t["goal_index"] == ref["goal_index"]
        assert np.array_equal(t["x"], ref["x"]) and np.array_equal(t["cost"], ref["cost"])


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["dubins", "rs", "rs_cost"])
def test_gpu_warp_and_cta_execution_agree_with_the_oracle(kind, oracle_lib):
    """exec_mode: a warp per query and a CTA per query give the same trees bit for bit (and the oracle's), also when the
    near lists span several rounds of the team (rs_cost: unclipped near radius) and when goals are reached early."""
    from rrtk import dubins_planner as DP
    O = oracle_lib
    Q, iters = 10, 120
    rng = np.random.default_rng(5)
    obs = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (7, 5, 2), (9, 5, 2)]
    streams = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-np.pi, np.pi, (Q, iters, 1))], axis=2)
    starts = [[0.0, 0.0, 0.0]] * Q
    goals = [[10.0, 10.0, 0.0]] * (Q // 2) + [[6.0, 10.0, 1.0]] * (Q - Q // 2)
    streams[rng.integers(0, 101, (Q, iters)) <= 10] = (10.0, 10.0, 0.0)
    expand = float("inf") if kind == "rs_cost" else 3.0
    kw = dict(steer="dubins") if kind == "dubins" else dict(steer="rs", step_size=0.2, rs_cost=(kind == "rs_cost"))
    outs = {}
    for until in (True, False):
        for mode in ("warp", "cta"):
            outs[mode] = DP.run_batch(starts, goals, [obs] * Q, expand, iters, streams, robot_radius=0.3, near_cap=256,
                                      search_until_max_iter=until, exec_mode=mode, **kw)
        for q in range(Q):
            a, b = outs["warp"][q], outs["cta"][q]
            assert a["n"] == b["n"] and a["goal_index"] == b["goal_index"] and a["iters_done"] == b["iters_done"]
            for k in ("parent", "x", "y", "yaw", "cost", "edge_from", "edge_to"):
                assert np.array_equal(a[k], b[k], equal_nan=True), (kind, q, k)
        for q in (0, Q - 1):
            if kind == "dubins":
                ref = O.rrtstar_dubins_run(starts[q], goals[q], obs, expand, iters, 0.3, 50.0, 1.0, np.deg2rad(1.0), 0.5, until,
                                           streams[q], O.MATH_CR)
            else:
                ref = O.rrtstar_rs_run(starts[q], goals[q], obs, expand, iters, 0.3, 50.0, 1.0, np.deg2rad(1.0), 0.5, until,
                                       streams[q], step_size=0.2, math_mode=O.MATH_CR, rs_cost=(kind == "rs_cost"))
            t = outs["cta"][q]
            assert t["n"] == ref["n"] and np.array_equal(t["parent"], ref["parent"]) and t["goal_index"] == ref["goal_index"]
            assert np.array_equal(t["x"], ref["x"]) and np.array_equal(t["cost"], ref["cost"], equal_nan=True)
