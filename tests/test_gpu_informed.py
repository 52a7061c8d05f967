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
