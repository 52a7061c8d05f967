"""GPU tests of RRT-Dubins (rrt_03:1402-1456, rrtk_rrt_dubins_run_dev / rrtk.RRTDubins) against the C oracle in cr mode
(bit for bit) and the unmodified reference's fixtures (same tree topology, poses and Dubins-length costs within 1e-9)."""
import json
import math
import random

import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu

RRT03 = golden_names("rrt03_")


def _run_gpu(m, stream):
    from rrtk import dubins_planner as DP
    return DP.run_rrt_batch([m["start"]], [m["goal"]], [m["obstacle_list"]], m["iters"], stream[None], m["robot_radius"],
                            m["curvature"], m["goal_yaw_th"], m["goal_xy_th"], m["search_until_max_iter"],
                            m.get("play_area"))[0]


@pytest.mark.parametrize("name", RRT03)
def test_kernel_bitwise_vs_oracle_cr_and_topology_vs_reference(name, oracle_lib):
    from rrtk import dubins_planner as DP
    O = oracle_lib
    g, m = load_golden(name)
    t = _run_gpu(m, g["stream"])
    ref = O.rrt_dubins_run(m["start"], m["goal"], m["obstacle_list"], m["iters"], m["robot_radius"], m["curvature"],
                           m["goal_yaw_th"], m["goal_xy_th"], m["search_until_max_iter"], g["stream"], m.get("play_area"),
                           math_mode=O.MATH_CR)
    assert t["status"] == 0 and t["n"] == ref["n"] and t["iters_done"] == ref["iters_done"]
    assert t["goal_index"] == ref["goal_index"]
    for k in ("x", "y", "yaw", "cost", "parent", "edge_from", "edge_to"):
        assert np.array_equal(t[k], ref[k]), k
    path = DP.final_course(t, m["start"], m["goal"], m["curvature"])
    assert (path is None and ref["path"] is None) or np.array_equal(np.array(path), np.array(ref["path"]))
    # the unmodified reference: same tree, values to 1e-9 (its libm is an ulp off the correctly rounded results here and there)
    assert t["n"] == len(g["x"]) and np.array_equal(t["parent"], g["parent"])
    for k in ("x", "y", "yaw", "cost"):
        assert np.allclose(t[k], g[k], rtol=0, atol=1e-9), k
    assert (path is None) == (len(g["path"]) == 0)
    if path is not None:
        assert len(path) == len(g["path"]) and np.allclose(np.array(path), g["path"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("name", ["rrt03_builtin_sobol_200", "rrt03_sobol_early_600", "rrt03_uniform_own_400"])
def test_class_draws_the_references_samples(name):
    """rrtk.RRTDubins with its own sampler after random.seed(k): the stream equals the one the reference drew (3-D Sobol
    points mapped by rrt_03:1545-1562, or the uniform draws), the tree has the reference's topology, and after an early
    exit sobol_inter_ stands where the reference's lazily drawing loop left it."""
    import rrtk
    g, m = load_golden(name)
    random.seed(m["seed"])
    rrt = rrtk.RRTDubins(m["start"], m["goal"], m["obstacle_list"], m["rand_area"], m["goal_sample_rate"], m["max_iter"],
                         m.get("play_area"), m["robot_radius"], m["sobol_sampler"], m["curvature"], m["goal_yaw_th"],
                         m["goal_xy_th"])
    path = rrt.planning(animation=False, search_until_max_iter=m["search_until_max_iter"])
    t = rrt.tree_arrays()
    assert t["iters_done"] == m["iters"] and rrt.sobol_inter_ == m["sobol_inter_"]
    assert np.array_equal(t["parent"], g["parent"]) and np.allclose(t["x"], g["x"], rtol=0, atol=1e-9)
    assert np.allclose(t["cost"], g["cost"], rtol=0, atol=1e-9)
    assert (path is None) == (len(g["path"]) == 0)
    assert len(rrt.node_list) == len(g["x"]) and rrt.node_list[-1].parent is rrt.node_list[int(g["parent"][-1])]
    if m["iters"] < m["max_iter"]:
        # the reference's next draw after its early exit
        ns_state = random.getstate()
        random.seed(m["seed"])
        for _ in range(m["iters"]):
            if random.randint(0, 100) > m["goal_sample_rate"] and not m["sobol_sampler"]:
                random.uniform(0, 1), random.uniform(0, 1), random.uniform(0, 1)
        assert random.getstate() == ns_state


def test_none_steer_with_play_area_raises_attribute_error():
    import rrtk
    start, goal = [0.0, 0.0, 0.0], [4.0, 0.0, 0.0]
    stream = np.array([goal, goal, [1.0, 1.0, 0.3]])
    rrt = rrtk.RRTDubins(start, goal, [], [-2, 6], max_iter=3, play_area=[-5.0, 5.0, -5.0, 5.0])
    with pytest.raises(AttributeError):
        rrt.planning(animation=False, sample_stream=stream)
    rrt = rrtk.RRTDubins(start, goal, [], [-2, 6], max_iter=3, goal_yaw_th=0.1)
    path = rrt.planning(animation=False, sample_stream=stream)      # no play area: the None steer is just skipped
    assert rrt.tree_arrays()["n"] == 3 and rrt.tree_arrays()["goal_index"] == 1 and path is not None


def test_batch_equals_single_queries(oracle_lib):
    """32 queries in one launch (ragged obstacle lists, one play area) == the oracle query by query."""
    from rrtk import dubins_planner as DP
    O = oracle_lib
    rng = np.random.default_rng(33)
    Q, iters = 32, 150
    starts = np.column_stack([rng.uniform(0, 2, (Q, 2)), rng.uniform(-math.pi, math.pi, Q)])
    goals = np.column_stack([rng.uniform(8, 12, (Q, 2)), rng.uniform(-math.pi, math.pi, Q)])
    obs = [[(float(x), float(y), float(r)) for x, y, r in zip(rng.uniform(2, 9, k), rng.uniform(2, 9, k), rng.uniform(0.3, 1.2, k))]
           for k in rng.integers(0, 9, Q)]
    st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
    play = [-3.0, 14.0, -2.5, 14.5]
    res = DP.run_rrt_batch(starts, goals, obs, iters, st, 0.25, 1.3, math.radians(40.0), 1.5, True, play)
    for q in range(Q):
        ref = O.rrt_dubins_run(starts[q], goals[q], obs[q] or np.zeros((0, 3)), iters, 0.25, 1.3, math.radians(40.0), 1.5, True,
                               st[q], play, math_mode=O.MATH_CR)
        t = res[q]
        assert t["n"] == ref["n"] and t["goal_index"] == ref["goal_index"], q
        for k in ("x", "y", "yaw", "cost", "parent"):
            assert np.array_equal(t[k], ref[k]), (q, k)
