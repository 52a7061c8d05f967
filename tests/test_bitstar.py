"""BIT* (rrt_08).  Fixtures tests/golden/rrt08_*.npz hold every ordered container the unmodified reference ends with
(tree vertices and edges, g-scores, parent map, sample dict, both queues, final path) on a recorded stream of unit draws."""
import numpy as np
import pytest

from conftest import load_golden

CASES = ["rrt08_builtin_80", "rrt08_builtin_200", "rrt08_dense_120", "rrt08_far_goal_300"]
KEYS = ("vertices", "g_vertices", "edges", "parent_of", "sample_ids", "sample_xy", "vertex_queue", "edge_queue", "path")


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_matches_reference_bit_for_bit(name, oracle_lib):
    g, m = load_golden(name)
    r = oracle_lib.bitstar_plan(m["start"], m["goal"], m["obstacleList"], m["randArea"], m["maxIter"], g["draws"])
    assert r["status"] == 0 and r["draws_used"] == m["draws_used"]
    assert r["g_goal"] == m["g_goal"]
    for k in KEYS:
        assert r[k].shape == g[k].shape and np.array_equal(r[k], g[k]), k


def test_c_oracle_index_error_livelock_and_draw_exhaustion(oracle_lib):
    """Sparse samples (huge randArea): expanding the start finds no neighbour, both queues are empty inside the expansion
    loop, where the reference raises IndexError (status 1).  A start walled in by a circle: every edge is skipped,
    `iterations` stays 0 and the reference replays the same batch forever (status 2).  Too few draws: status -1."""
    d = np.random.default_rng(5).random(4000)
    r = oracle_lib.bitstar_plan([0.0, 0.0], [150.0, 150.0], [(50.0, 50.0, 3.0)], [-2, 400], 50, d)
    assert r["status"] == 1
    r = oracle_lib.bitstar_plan([0.0, 0.0], [10.0, 10.0], [(0.0, 0.0, 3.0)], [-2, 15], 50, d)
    assert r["status"] == 2
    r = oracle_lib.bitstar_plan([-1.0, 0.0], [3.0, 8.0], [(5, 5, 0.5)], [-2, 15], 50, d[:100])
    assert r["status"] == -1


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_matches_oracle_and_reference(name, oracle_lib):
    from rrtk import bitstar as B
    g, m = load_golden(name)
    r = B.run_batch([m["start"]], [m["goal"]], [m["obstacleList"]], m["randArea"], m["maxIter"], g["draws"][None], inspect=True)[0]
    o = oracle_lib.bitstar_plan(m["start"], m["goal"], m["obstacleList"], m["randArea"], m["maxIter"], g["draws"],
                                math_mode=oracle_lib.MATH_CR)
    assert r["status"] == 0 and r["draws_used"] == o["draws_used"] and r["g_goal"] == o["g_goal"]
    for k in ("batches", "resets", "skipped", "expansions"):
        assert r[k] == o[k], k
    for k in KEYS:
        assert r[k].shape == o[k].shape and np.array_equal(r[k], o[k]), k
    # the reference: every discrete container identical; the sample coordinates differ at most in the last bit of sin / cos
    for k in ("vertices", "edges", "parent_of", "sample_ids", "vertex_queue", "edge_queue", "path"):
        assert r[k].shape == g[k].shape and np.array_equal(r[k], g[k]), k
    np.testing.assert_allclose(r["g_vertices"], g["g_vertices"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(r["sample_xy"], g["sample_xy"], rtol=0, atol=1e-12)


@pytest.mark.gpu
def test_gpu_batch_of_random_queries_matches_oracle(oracle_lib):
    from rrtk import bitstar as B
    rng = np.random.default_rng(11)
    Q, it = 24, 120
    starts = rng.uniform(-1, 3, (Q, 2)); goals = rng.uniform(8, 14, (Q, 2))
    obs = [[(float(x), float(y), float(s)) for (x, y), s in zip(rng.uniform(3, 11, (5, 2)), rng.uniform(0.4, 1.5, 5))]
           for _ in range(Q)]
    obs[3] = [(float(starts[3, 0]), float(starts[3, 1]), 3.0)]          # walled-in start: the reference never returns
    draws = rng.random((Q, 8000))
    res = B.run_batch(starts, goals, obs, [-2, 15], it, draws, inspect=True)
    n_err = 0
    for i in range(Q):
        o = oracle_lib.bitstar_plan(starts[i], goals[i], obs[i], [-2, 15], it, draws[i], math_mode=oracle_lib.MATH_CR)
        if o["status"] == 2:
            assert res[i]["status"] == 64
            n_err += 1
            continue
        assert res[i]["status"] == 0 and o["status"] == 0
        assert res[i]["g_goal"] == o["g_goal"] and res[i]["draws_used"] == o["draws_used"]
        for k in KEYS:
            assert res[i][k].shape == o[k].shape and np.array_equal(res[i][k], o[k]), (i, k)
    assert n_err >= 1


@pytest.mark.gpu
def test_gpu_index_error_status(oracle_lib):
    from rrtk import bitstar as B
    d = np.random.default_rng(5).random(4000)
    r = B.run_batch([[0.0, 0.0]], [[150.0, 150.0]], [[(50.0, 50.0, 3.0)]], [-2, 400], 50, d[None])[0]
    assert r["status"] == 16


@pytest.mark.gpu
def test_gpu_bitstar_class_reproduces_the_reference_path():
    import random
    import rrtk
    g, m = load_golden("rrt08_builtin_80")
    b = rrtk.BITStar(m["start"], m["goal"], m["obstacleList"], m["randArea"], maxIter=m["maxIter"])
    path = b.plan(animation=False, draws=g["draws"])
    assert np.array_equal(np.array(path), g["path"])
    random.seed(4)
    assert isinstance(b.plan(animation=False), list)
