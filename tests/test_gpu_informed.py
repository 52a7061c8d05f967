"""GPU parity tests of Informed RRT* (rrt_07) against the oracle and the reference fixtures."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu
NAMES = golden_names("rrt07_")
CR_EXACT = [n for n in NAMES if n != "rrt07_builtin_2500"]


def _run(m, g):
    import rrtk
    r = rrtk.InformedRRTStar(m["start"], m["goal"], m["obstacle_list"], m["rand_area"], m["expand_dis"],
                             m["goal_sample_rate"], m["max_iter"])
    path = r.informed_rrt_star_search(animation=False, free_samples=g["free"], ball_draws=g["ball"])
    return r, path


@pytest.mark.parametrize("name", NAMES)
def test_informed_bitwise_vs_oracle_cr(name, oracle_lib):
    O = oracle_lib
    g, m = load_golden(name)
    ref = O.informed_run(m["start"], m["goal"], m["obstacle_list"], m["expand_dis"], m["max_iter"], m["rot"],
                         g["free"], g["ball"], O.MATH_CR)
    r, path = _run(m, g)
    a = r.tree_arrays()
    assert a["n"] == ref["n"]
    assert np.array_equal(a["parent"], ref["parent"])
    assert np.array_equal(a["x"], ref["x"]) and np.array_equal(a["y"], ref["y"])
    assert np.array_equal(a["cost"], ref["cost"])
    assert a["c_best"] == ref["c_best"]
    assert path == ref["path"]


@pytest.mark.parametrize("name", CR_EXACT)
def test_informed_bitwise_vs_reference_fixture(name):
    g, m = load_golden(name)
    r, path = _run(m, g)
    a = r.tree_arrays()
    assert np.array_equal(a["parent"], g["parent"])
    assert np.array_equal(a["x"], g["x"]) and np.array_equal(a["y"], g["y"])
    assert np.array_equal(a["cost"], g["cost"])
    assert np.array_equal(np.array(path, float), g["path"])
    nl = r.node_list
    assert nl[0].parent is None and all(isinstance(n.parent, int) for n in nl[1:])


def test_rotation_matches_fixture():
    from rrtk import informed
    g, m = load_golden("rrt07_alt_800")
    assert informed.rotation_to_world_frame(m["start"], m["goal"]) == m["rot"]


def test_informed_batch_random(oracle_lib):
    """A batch of random scenarios (more obstacles, some with no solution) against oracle[cr]."""
    from rrtk import informed
    O = oracle_lib
    Q, iters = 12, 500
    rng = np.random.default_rng(17)
    starts, goals, obs_lists, frees, balls = [], [], [], [], []
    for q in range(Q):
        obs = [(float(x), float(y), float(r)) for (x, y), r in
               zip(rng.uniform(1, 12, (20 + 5 * q, 2)), rng.uniform(0.2, 0.9, 20 + 5 * q))]
        obs = [o for o in obs if np.hypot(o[0], o[1]) > o[2] + 0.6 and np.hypot(o[0] - 10, o[1] - 9) > o[2] + 0.6]
        starts.append([0.0, 0.0]); goals.append([10.0, 9.0]); obs_lists.append(obs)
        f = rng.uniform(-2, 15, (iters, 2)); coin = rng.integers(0, 101, iters) <= 10; f[coin] = goals[-1]
        frees.append(f); balls.append(rng.random((iters, 2)))
    out = informed.run_batch(starts, goals, obs_lists, 0.7, iters, np.array(frees), np.array(balls))
    solved = 0
    for q in range(Q):
        rot = informed.rotation_to_world_frame(starts[q], goals[q])
        ref = O.informed_run(starts[q], goals[q], obs_lists[q], 0.7, iters, rot, frees[q], balls[q], O.MATH_CR)
        a = out[q]
        assert a["n"] == ref["n"] and np.array_equal(a["parent"], ref["parent"])
        assert np.array_equal(a["x"], ref["x"]) and np.array_equal(a["cost"], ref["cost"])
        assert a["c_best"] == ref["c_best"] and a["path"] == ref["path"]
        solved += a["path"] is not None
    assert solved >= 1


def test_informed_batch_equal_d2_and_long_near_lists(oracle_lib):
    """The `.index()` mapping of find_near_nodes in the batched kernel: the hash set that detects equal d2 among the hits
    must send the iteration to the exact mapping.  Repeated samples create coincident nodes (equal d2 at every later
    iteration that has both in its near list), a mirrored pair gives two DIFFERENT positions at bitwise-equal d2 from nodes
    on the axis; 1500 iterations push the near lists past 3/4 of the detector's 1024 slots (always-exact path)."""
    from rrtk import informed
    O = oracle_lib
    start, goal = [0.0, 0.0], [6.0, 0.25]
    obs = [(3.0, 2.5, 1.0), (3.0, -2.5, 1.0)]
    iters = 1500
    rng = np.random.default_rng(11)
    free = rng.uniform(-3.0, 8.0, (iters, 2))
    free[rng.integers(0, 101, iters) <= 10] = goal
    ball = rng.random((iters, 2))
    free[0] = (-1.0, 5.0); free[1] = (-1.0, -5.0)   # mirrored pair about the x axis, both grown from the root
    free[2:6] = (0.1, 0.0)                          # nearest stays the root: (0.5, 0) four times, on the axis
    free[200:260:2] = (0.1, 0.0)
    rot = informed.rotation_to_world_frame(start, goal)
    ref = O.informed_run(start, goal, obs, 0.5, iters, rot, free, ball, O.MATH_CR)
    xy = np.column_stack([ref["x"], ref["y"]])
    assert len(np.unique(xy, axis=0)) < len(xy), "scenario should contain coincident nodes"
    # the same query three times in one launch (several warps, one of them beside a plain random query)
    free2 = rng.uniform(-3.0, 8.0, (iters, 2))
    out = informed.run_batch([start] * 4, [goal] * 4, [obs] * 4, 0.5, iters, np.array([free, free, free2, free]),
                             np.array([ball] * 4))
    for q in (0, 1, 3):
        a = out[q]
        assert a["n"] == ref["n"] and np.array_equal(a["parent"], ref["parent"])
        assert np.array_equal(a["x"], ref["x"]) and np.array_equal(a["y"], ref["y"]) and np.array_equal(a["cost"], ref["cost"])
        assert a["c_best"] == ref["c_best"] and a["path"] == ref["path"]
    ref2 = O.informed_run(start, goal, obs, 0.5, iters, rot, free2, ball, O.MATH_CR)
    assert out[2]["n"] == ref2["n"] and np.array_equal(out[2]["parent"], ref2["parent"]) and np.array_equal(out[2]["cost"], ref2["cost"])


def test_informed_warp_and_cta_execution_agree(oracle_lib):
    """exec_mode: a warp per query and a CTA per query build the same trees (and the oracle's), with near lists of several
    rounds per warp and best-path updates along the way."""
    from rrtk import informed
    O = oracle_lib
    Q, iters = 6, 700
    rng = np.random.default_rng(23)
    obs = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]
    starts, goals = [[0.0, 0.0]] * Q, [[6.0, 10.0]] * (Q - 1) + [[12.0, 3.0]]
    free = rng.uniform(-2, 15, (Q, iters, 2))
    coin = rng.integers(0, 101, (Q, iters)) <= 10
    for q in range(Q):
        free[q][coin[q]] = goals[q]
    ball = rng.random((Q, iters, 2))
    out = {m: informed.run_batch(starts, goals, [obs] * Q, 0.5, iters, free, ball, exec_mode=m) for m in ("warp", "cta")}
    for q in range(Q):
        a, b = out["warp"][q], out["cta"][q]
        assert a["n"] == b["n"] and a["c_best"] == b["c_best"] and a["path"] == b["path"] and a["status"] == b["status"]
        for k in ("x", "y", "cost", "parent"):
            assert np.array_equal(a[k], b[k]), (q, k)
    for q in (0, Q - 1):
        rot = informed.rotation_to_world_frame(starts[q], goals[q])
        ref = O.informed_run(starts[q], goals[q], obs, 0.5, iters, rot, free[q], ball[q], O.MATH_CR)
        b = out["cta"][q]
        assert b["n"] == ref["n"] and np.array_equal(b["parent"], ref["parent"]) and np.array_equal(b["cost"], ref["cost"])
        assert b["c_best"] == ref["c_best"] and b["path"] == ref["path"]
