"""GPU parity tests of the single-large-tree Informed RRT* kernel (BASELINE config 3,
rrtk_informed_tree_run_dev) against the oracle, the reference fixtures and itself at other grid sizes."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu
NAMES = golden_names("rrt07_")
CR_EXACT = [n for n in NAMES if n != "rrt07_builtin_2500"]
BUILTIN_OBS = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]


def _same(a, ref):
    assert a["n"] == ref["n"]
    assert np.array_equal(a["parent"], ref["parent"])
    assert np.array_equal(a["x"], ref["x"]) and np.array_equal(a["y"], ref["y"])
    assert np.array_equal(a["cost"], ref["cost"])
    assert a["c_best"] == ref["c_best"]
    assert a["path"] == ref["path"]


@pytest.mark.parametrize("batch", [1, 8])
@pytest.mark.parametrize("grid", [0, 3])
@pytest.mark.parametrize("name", NAMES)
def test_tree_bitwise_vs_oracle_cr(name, grid, batch, oracle_lib):
    from rrtk import informed
    O = oracle_lib
    g, m = load_golden(name)
    ref = O.informed_run(m["start"], m["goal"], m["obstacle_list"], m["expand_dis"], m["max_iter"], m["rot"],
                         g["free"], g["ball"], O.MATH_CR)
    run = informed.run_tree(m["start"], m["goal"], m["obstacle_list"], m["expand_dis"], m["max_iter"],
                            g["free"], g["ball"], grid=grid, batch=batch)
    a = run.arrays()
    assert a["status"] == 0 and run.info["iters_done"] == m["max_iter"]
    _same(a, ref)


@pytest.mark.parametrize("name", CR_EXACT)
def test_tree_bitwise_vs_reference_fixture(name):
    from rrtk import informed
    g, m = load_golden(name)
    a = informed.run_tree(m["start"], m["goal"], m["obstacle_list"], m["expand_dis"], m["max_iter"],
                          g["free"], g["ball"]).arrays()
    assert np.array_equal(a["parent"], g["parent"])
    assert np.array_equal(a["x"], g["x"]) and np.array_equal(a["y"], g["y"])
    assert np.array_equal(a["cost"], g["cost"])
    assert np.array_equal(np.array(a["path"], float), g["path"])


def _draws(rng, iters, goal, lo=-2.0, hi=15.0, rate=10):
    free = rng.uniform(lo, hi, (iters, 2))
    coin = rng.integers(0, 101, iters) <= rate
    free[coin] = goal
    return free, rng.random((iters, 2))


@pytest.mark.parametrize("batch", [1, 3, 8])
@pytest.mark.parametrize("grid,iters,seed", [(0, 6000, 1), (5, 12000, 2), (2, 5000, 3)])
def test_tree_random_scenes_vs_oracle(grid, iters, seed, batch, oracle_lib):
    """More circles, multi-pass scans (small grids wrap the ownership chunks), thousands of near hits."""
    from rrtk import informed
    O = oracle_lib
    rng = np.random.default_rng(seed)
    start, goal = [0.0, 0.0], [10.0, 9.0]
    obs = [(float(x), float(y), float(r)) for (x, y), r in zip(rng.uniform(1, 12, (40, 2)), rng.uniform(0.2, 0.8, 40))]
    obs = [o for o in obs if np.hypot(o[0], o[1]) > o[2] + 0.6 and np.hypot(o[0] - 10, o[1] - 9) > o[2] + 0.6]
    free, ball = _draws(rng, iters, goal)
    rot = informed.rotation_to_world_frame(start, goal)
    ref = O.informed_run(start, goal, obs, 0.5, iters, rot, free, ball, O.MATH_CR)
    run = informed.run_tree(start, goal, obs, 0.5, iters, free, ball, grid=grid, batch=batch)
    a = run.arrays()
    assert a["status"] == 0
    _same(a, ref)
    assert ref["path"] is not None and run.info["goal_events"] > 0 and run.info["resamples"] > 0


def test_tree_25000_iterations_vs_oracle(oracle_lib):
    """The largest tree the C oracle replays in half a minute (O(n) per iteration): 25 000 iterations, ~20 000 nodes, batched
    kernel on the full grid -- every node, cost, parent and the path, bit for bit."""
    from rrtk import informed
    O = oracle_lib
    rng = np.random.default_rng(4)
    iters = 25000
    start, goal = [0.0, 0.0], [10.0, 9.0]
    obs = [(float(x), float(y), float(r)) for (x, y), r in zip(rng.uniform(1, 12, (40, 2)), rng.uniform(0.2, 0.8, 40))]
    obs = [o for o in obs if np.hypot(o[0], o[1]) > o[2] + 0.6 and np.hypot(o[0] - 10, o[1] - 9) > o[2] + 0.6]
    free, ball = _draws(rng, iters, goal)
    rot = informed.rotation_to_world_frame(start, goal)
    run = informed.run_tree(start, goal, obs, 0.5, iters, free, ball, grid=0, batch=8)
    ref = O.informed_run(start, goal, obs, 0.5, iters, rot, free, ball, O.MATH_CR)
    a = run.arrays()
    assert a["status"] == 0 and ref["n"] > 15000
    _same(a, ref)


def test_tree_duplicate_positions_and_equal_d2(oracle_lib):
    """Repeated samples create identical nodes (the `.index()` quirk shadows the later twin); a mirrored pair
    gives two DIFFERENT positions at bitwise-equal d^2 from a node on the axis (exact slow path)."""
    from rrtk import informed
    O = oracle_lib
    start, goal = [0.0, 0.0], [6.0, 0.25]
    obs = [(3.0, 2.5, 1.0), (3.0, -2.5, 1.0)]
    rng = np.random.default_rng(11)
    iters = 1500
    free, ball = _draws(rng, iters, goal, -3.0, 8.0)
    free[0] = (-1.0, 5.0); free[1] = (-1.0, -5.0)   # mirrored pair about the x axis, both grown from the root
    free[2:6] = (0.1, 0.0)                          # nearest stays the root: (0.5, 0) four times, on the axis
    free[200:260:2] = (0.1, 0.0)
    rot = informed.rotation_to_world_frame(start, goal)
    ref = O.informed_run(start, goal, obs, 0.5, iters, rot, free, ball, O.MATH_CR)
    for grid, batch in ((0, 1), (2, 1), (0, 8), (2, 4)):
        run = informed.run_tree(start, goal, obs, 0.5, iters, free, ball, grid=grid, batch=batch)
        _same(run.arrays(), ref)
        assert run.info["slow_paths"] >= 1
    xy = np.column_stack([ref["x"], ref["y"]])
    assert len(np.unique(xy, axis=0)) < len(xy), "scenario should contain coincident nodes"


def test_tree_matches_batched_kernel():
    """Same search through the warp-per-query kernel (rrtk_informed_run_dev)."""
    from rrtk import informed
    rng = np.random.default_rng(4)
    iters = 4000
    free, ball = _draws(rng, iters, [6.0, 10.0])
    b = informed.run_batch([[0.0, 0.0]], [[6.0, 10.0]], [BUILTIN_OBS], 0.5, iters, free[None], ball[None])[0]
    for batch in (1, 8):
        a = informed.run_tree([0.0, 0.0], [6.0, 10.0], BUILTIN_OBS, 0.5, iters, free, ball, batch=batch).arrays()
        _same(a, b)


def test_tree_large_grid_independence_and_invariants():
    """150 000 iterations (multi-chunk ownership at full grid): identical bits for 148 and 37 CTAs, plus the
    invariants the reference's tree satisfies."""
    import torch
    from rrtk import informed
    rng = np.random.default_rng(9)
    iters = 150_000
    start, goal = [0.0, 0.0], [6.0, 10.0]
    free, ball = _draws(rng, iters, goal)
    runs = [informed.run_tree(start, goal, BUILTIN_OBS, 0.5, iters, free, ball, grid=g, batch=bt)
            for g, bt in ((0, 1), (37, 1), (0, 8), (37, 5))]
    torch.cuda.synchronize()
    a, b = runs[0].arrays(), runs[1].arrays()
    _same(a, b)
    _same(runs[2].arrays(), a)
    _same(runs[3].arrays(), a)
    i = runs[0].info
    assert i["status"] == 0 and i["iters_done"] == iters and i["n_nodes"] > 100_000
    n, par = a["n"], a["parent"]
    assert par[0] == -1 and (par[1:] >= 0).all() and (par[1:] < n).all()
    px, py = a["x"][par[1:]], a["y"][par[1:]]
    d = np.hypot(a["x"][1:] - px, a["y"][1:] - py)
    assert (a["cost"][1:] >= a["cost"][par[1:]] + d - 1e-9).all()          # costs only go stale upwards
    for ox, oy, r in BUILTIN_OBS:                                          # every edge is collision free
        wx, wy = a["x"][1:] - px, a["y"][1:] - py
        l2 = np.maximum(wx * wx + wy * wy, 1e-300)
        t = np.clip(((ox - px) * wx + (oy - py) * wy) / l2, 0.0, 1.0)
        dd = (ox - px - t * wx) ** 2 + (oy - py - t * wy) ** 2
        assert (dd > r * r - 1e-9).all()
    path = np.array(a["path"])
    assert np.isclose(np.hypot(*(path[1:] - path[:-1]).T).sum(), a["c_best"], rtol=1e-12)
    assert np.hypot(6.0, 10.0) < a["c_best"] < 17.5                         # rrt_07's scenario: optimum ~16.9


def test_tree_argument_errors():
    from rrtk import _lib, informed
    with pytest.raises(_lib.RrtkError):
        informed.run_tree([0, 0], [1, 1], [(5.0, 5.0, 0.1)] * 600, 0.5, 10, np.zeros((10, 2)), np.zeros((10, 2)))
    with pytest.raises(_lib.RrtkError):
        informed.run_tree([0, 0], [1, 1], [], 0.5, 10, np.zeros((5, 2)), np.zeros((10, 2)))
