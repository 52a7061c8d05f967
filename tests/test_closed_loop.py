"""Closed-loop RRT* (rrt_10): RRT*-Reeds-Shepp with Reeds-Shepp-length costs (rrt_10:1005-1207) and the pure-pursuit
feasibility filter (rrt_10:1215-1582).  Fixtures tests/golden/rrt10_cl_*.npz were produced by the unmodified reference
(oracle/make_golden.py run_rrt10)."""
import math

import numpy as np
import pytest

from conftest import load_golden

CASES = ["rrt10_cl_60", "rrt10_cl_builtin_150", "rrt10_cl_radius_100"]


def _courses(g):
    off = g["course_off"]
    return [g["course"][off[k]:off[k + 1]] for k in range(len(off) - 1)]


def _trajs(g):
    off = g["traj_off"]
    return [g["traj"][off[k]:off[k + 1]] for k in range(len(off) - 1)]


def _goal_indexes(x, y, yaw, goal, xy_th, yaw_th):
    """get_goal_indexes (rrt_10:1561-1580)."""
    return [i for i in range(len(x)) if math.hypot(x[i] - goal[0], y[i] - goal[1]) <= xy_th
            and abs(yaw[i] - goal[2]) <= yaw_th]


@pytest.mark.parametrize("name", CASES[:1] + CASES[2:])
def test_pyport_tree_matches_reference(name):
    import pyport as P
    g, m = load_golden(name)
    port = P.RRTStarRSPort(m["start"], m["goal"], m["obstacle_list"], expand_dis=float("inf"), max_iter=m["max_iter"],
                           robot_radius=m["robot_radius"], connect_circle_dist=m["connect_circle_dist"], curvature=1.0,
                           step_size=0.2, rs_cost=True)
    port.planning([tuple(map(float, r)) for r in g["stream"]])
    assert port.x == g["x"].tolist() and port.y == g["y"].tolist() and port.yaw == g["yaw"].tolist()
    assert port.cost == g["cost"].tolist() and port.parent == g["parent"].tolist()
    gi = _goal_indexes(port.x, port.y, port.yaw, m["goal"], m["xy_th"], m["yaw_th"])
    assert gi == g["goal_idx"].tolist()
    for k, i in enumerate(gi[:4]):
        assert np.array_equal(np.array(port.final_course(i)), _courses(g)[k])


@pytest.mark.parametrize("name", CASES)
def test_pyport_closed_loop_matches_reference(name):
    import pyport as P
    g, m = load_golden(name)
    courses, trajs = _courses(g), _trajs(g)
    best, info = P.cl_search_best_feasible([c.tolist() for c in courses], m["obstacle_list"], m["robot_radius"],
                                           m["target_speed"], m["yaw_th"], m["invalid_travel_ratio"])
    assert [b == 0 for b, _, _ in info] == g["found"].tolist()
    assert (best >= 0) == m["flag"]
    for k in (0, len(courses) // 2, len(courses) - 1):
        bits, t, x, y, yaw, v, a, d = P.cl_check_tracking(courses[k].tolist(), m["obstacle_list"], m["robot_radius"],
                                                          m["target_speed"], m["yaw_th"], m["invalid_travel_ratio"])
        assert np.array_equal(np.array([x, y, yaw, v, t, a, d]).T, trajs[k])
    if best >= 0:   # the winner (+ the goal pose appended by search_best_feasible_path :1513-1516)
        w = g["winner_xyyaw"]
        assert np.array_equal(w[:-1], trajs[best][:, 0:3]) and w[-1].tolist() == [float(v) for v in m["goal"]]


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_tree_matches_reference(name, oracle_lib):
    g, m = load_golden(name)
    r = oracle_lib.rrtstar_rs_run(m["start"], m["goal"], m["obstacle_list"], float("inf"), m["max_iter"], m["robot_radius"],
                                  m["connect_circle_dist"], 1.0, np.deg2rad(1.0), 0.5, True, g["stream"], step_size=0.2,
                                  rs_cost=True)
    assert np.array_equal(r["parent"], g["parent"])
    assert np.array_equal(r["x"], g["x"]) and np.array_equal(r["y"], g["y"]) and np.array_equal(r["yaw"], g["yaw"])
    assert np.array_equal(r["cost"], g["cost"])


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_closed_loop_matches_reference(name, oracle_lib):
    """libm mode against the reference's trajectories: same verdicts and lengths, values bit for bit (np.hypot == C hypot)."""
    g, m = load_golden(name)
    courses, trajs = _courses(g), _trajs(g)
    best, res = oracle_lib.closed_loop_best([c[::-1] for c in courses], m["obstacle_list"], m["robot_radius"],
                                            m["target_speed"], m["yaw_th"], m["invalid_travel_ratio"])
    assert [r["bits"] == 0 for r in res] == g["found"].tolist()
    for r, t in zip(res, trajs):
        assert r["traj"].shape == t.shape
        assert np.array_equal(r["traj"], t)
    if m["flag"]:
        assert np.array_equal(trajs[best][:, 0:3], g["winner_xyyaw"][:-1])


@pytest.mark.parametrize("name", CASES[:1])
def test_c_oracle_cr_mode_stays_on_the_reference_decisions(name, oracle_lib):
    """cr mode (the arithmetic the GPU uses) differs from libm only in rare last-bit roundings of the leaf functions:
    same verdicts, same step counts, trajectories equal to ~1e-9."""
    g, m = load_golden(name)
    courses, trajs = _courses(g), _trajs(g)
    _, res = oracle_lib.closed_loop_best([c[::-1] for c in courses], m["obstacle_list"], m["robot_radius"], m["target_speed"],
                                         m["yaw_th"], m["invalid_travel_ratio"], math_mode=oracle_lib.MATH_CR)
    assert [r["bits"] == 0 for r in res] == g["found"].tolist()
    for r, t in zip(res, trajs):
        assert r["traj"].shape == t.shape
        np.testing.assert_allclose(r["traj"], t, rtol=0, atol=1e-8)


def test_tan_correctly_rounded_identities(oracle_lib):
    L = oracle_lib.lib()
    assert L.orc_cr_tan(0.0) == 0.0 and math.copysign(1.0, L.orc_cr_tan(-0.0)) == -1.0
    assert L.orc_cr_tan(math.radians(40.0)) == 0.8390996311772799
    xs = np.random.default_rng(3).uniform(-1.2, 1.2, 4000)
    assert sum(L.orc_cr_tan(float(x)) != math.tan(float(x)) for x in xs) < 40   # glibc tan is correctly rounded ~always
