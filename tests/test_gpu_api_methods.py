"""The reference's per-method API on rrtk.RRT / rrtk.RRTStar (rrt_04:1086-1115, :1196-1238, :1242-1384):
Node.path_x / path_y against the unmodified reference's own, the device-backed per-step methods against the fused kernel,
and the refusal to plan with an overridden method the kernel would ignore."""
import math
import random

import numpy as np
import pytest

from conftest import golden_names, load_golden

PATHS = golden_names("paths04_")


def _planner(m, cls=None, **kw):
    import rrtk
    cls = cls or rrtk.RRTStar
    return cls(m["start"], m["goal"], m["obstacle_list"], m["rand_area"], m["expand_dis"], m["path_resolution"],
               m["goal_sample_rate"], kw.get("max_iter", m["max_iter"]), m["play_area"], m["robot_radius"], m["sobol_sampler"],
               m["connect_circle_dist"], m["search_until_max_iter"])


def test_overridden_fused_method_is_refused():
    """(no GPU needed) a subclass -- or an instance attribute -- that replaces a method the fused kernel stands in for."""
    import rrtk
    _, m = load_golden(PATHS[0])

    class MySteer(rrtk.RRTStar):
        def steer(self, from_node, to_node, extend_length=float("inf")):
            return None

    class MyRewire(rrtk.RRTStar):
        def rewire(self, new_node, near_inds):
            pass

    for cls, word in ((MySteer, "steer"), (MyRewire, "rewire")):
        with pytest.raises(rrtk.RrtkError, match=word):
            _planner(m, cls).planning(animation=False)
    r = _planner(m)
    r.check_collision = lambda node, obs, rr: True
    with pytest.raises(rrtk.RrtkError, match="check_collision"):
        r.planning(animation=False)

    class Harmless(rrtk.RRTStar):          # new names and the samplers are fine
        def draw_graph(self, rnd=None):
            pass
    Harmless._check_overrides(_planner(m, Harmless))


@pytest.mark.gpu
@pytest.mark.parametrize("name", PATHS)
def test_node_paths_match_the_reference(name):
    """Every node's path_x / path_y after planning(): same number of points as the reference's node, every point within
    1e-12, and bit-identical for all but the few edges whose libm cos / sin differ from the correctly rounded ones."""
    g, m = load_golden(name)
    rrt = _planner(m)
    rrt.planning(animation=False, sample_stream=g["stream"] if len(g["stream"]) >= m["max_iter"] else
                 np.vstack([g["stream"], np.zeros((m["max_iter"] - len(g["stream"]), 2))]))
    assert np.array_equal(rrt.tree_arrays()["parent"], g["parent"])
    off, ref = g["path_off"], g["path_xy"]
    exact = 0
    for k, nd in enumerate(rrt.node_list):
        want = ref[off[k]:off[k + 1]]
        got = np.column_stack([nd.path_x, nd.path_y]) if len(nd.path_x) else np.zeros((0, 2))
        assert got.shape == want.shape, (k, got.shape, want.shape)
        assert np.allclose(got, want, rtol=0.0, atol=1e-12), k
        exact += int(np.array_equal(got, want))
        if k:
            assert nd.path_x[-1] == nd.x and nd.path_y[-1] == nd.y
            assert nd.path_x[0] == nd.parent.x and nd.path_y[0] == nd.parent.y
    assert exact >= 0.95 * len(rrt.node_list), (exact, len(rrt.node_list))


@pytest.mark.gpu
def test_loop_built_from_the_per_step_methods_equals_the_fused_kernel():
    """The reference's planning() loop (rrt_04:1036-1084) written against rrtk's steer / check_collision /
    get_nearest_node_index / find_near_nodes / choose_parent / rewire gives the tree the kernel gives."""
    import rrtk
    g, m = load_golden("paths04_c1_sobol_500")
    iters = 220
    fused = _planner(m, max_iter=iters)
    fused.planning(animation=False, sample_stream=g["stream"][:iters])
    r = _planner(m, max_iter=iters)
    r.node_list = [r.start]
    for i in range(iters):
        rnd = r.Node(float(g["stream"][i, 0]), float(g["stream"][i, 1]))
        ni = r.get_nearest_node_index(r.node_list, rnd)
        new_node = r.steer(r.node_list[ni], rnd, r.expand_dis)
        near_node = r.node_list[ni]
        new_node.cost = near_node.cost + math.hypot(new_node.x - near_node.x, new_node.y - near_node.y)
        if r.check_if_outside_play_area(new_node, r.play_area) and r.check_collision(new_node, r.obstacle_list, r.robot_radius):
            near_inds = r.find_near_nodes(new_node)
            upd = r.choose_parent(new_node, near_inds)
            if upd:
                r.rewire(upd, near_inds)
                r.node_list.append(upd)
            else:
                r.node_list.append(new_node)
    a = fused.tree_arrays()
    assert len(r.node_list) == len(a["x"])
    idx = {id(n): i for i, n in enumerate(r.node_list)}
    assert [(-1 if n.parent is None else idx[id(n.parent)]) for n in r.node_list] == a["parent"].tolist()
    assert np.array_equal([n.x for n in r.node_list], a["x"]) and np.array_equal([n.y for n in r.node_list], a["y"])
    assert np.array_equal([n.cost for n in r.node_list], a["cost"])
    for mine, theirs in zip(r.node_list, fused.node_list):
        assert mine.path_x == theirs.path_x and mine.path_y == theirs.path_y
    assert r.search_best_goal_node() == fused.search_best_goal_node()


@pytest.mark.gpu
def test_own_sampler_method_feeds_the_kernel():
    """get_random_node_sobol replaced on the instance (how the reference's users inject samples): called once per iteration."""
    g, m = load_golden("paths04_c1_sobol_500")
    it = iter(g["stream"])
    r = _planner(m)
    calls = []

    def sampler():
        x, y = next(it)
        calls.append(1)
        return r.Node(float(x), float(y))
    r.get_random_node_sobol = sampler
    r.planning(animation=False)
    assert len(calls) == m["max_iter"] and np.array_equal(r.tree_arrays()["parent"], g["parent"])
    # the stock samplers draw what the reference draws (rrt_04:1132-1155)
    random.seed(m["seed"])
    s = _planner(m)
    got = np.array([[n.x, n.y] for n in (s.get_random_node_sobol() for _ in range(40))])
    assert np.array_equal(got, g["stream"][:40]) and s.sobol_inter_ == int((got != [m["goal"][0], m["goal"][1]]).any(axis=1).sum())


@pytest.mark.gpu
def test_nearest_and_near_primitives_follow_the_reference_lists():
    from rrtk import engine
    rng = np.random.default_rng(3)
    xy = rng.uniform(-2, 15, (3000, 2))
    xy[100] = xy[7]; xy[2500] = xy[7]; xy[40] = xy[2999]          # equal squared distances (the `.index()` quirk)
    for _ in range(6):
        c = rng.uniform(0, 13, 2)
        d = [(x - c[0]) ** 2 + (y - c[1]) ** 2 for x, y in xy.tolist()]
        assert engine.nearest_index(xy, [c])[0] == d.index(min(d))
        r2 = float(rng.uniform(0.5, 9.0))
        assert engine.near_indices(xy, c[0], c[1], r2) == [d.index(v) for v in d if v <= r2]
    c = xy[7]                                                     # three nodes at distance 0, and a tie for the minimum
    d = [(x - c[0]) ** 2 + (y - c[1]) ** 2 for x, y in xy.tolist()]
    assert engine.nearest_index(xy, [c])[0] == 7
    assert engine.near_indices(xy, c[0], c[1], 4.0) == [d.index(v) for v in d if v <= 4.0]


def test_dubins_overridden_method_is_refused():
    """(no GPU needed) rrt_05 / rrt_03 classes: same rule as rrtk.RRTStar."""
    import rrtk

    class MySteer(rrtk.RRTStarDubins):
        def steer(self, from_node, to_node):
            return None

    class MyCC(rrtk.RRTDubins):
        @staticmethod
        def check_collision(node, obstacleList, robot_radius):
            return True
    with pytest.raises(rrtk.RrtkError, match="steer"):
        MySteer([0, 0, 0], [10, 10, 0], [(5, 5, 1)], [-2, 15], max_iter=10).planning(animation=False)
    with pytest.raises(rrtk.RrtkError, match="check_collision"):
        MyCC([0, 0, 0], [10, 10, 0], [(5, 5, 1)], [-2, 15], max_iter=10).planning(animation=False)


@pytest.mark.gpu
def test_dubins_loop_built_from_the_per_step_methods_equals_the_fused_kernel():
    """rrt_05's planning() loop (:1416-1456) written against rrtk.RRTStarDubins' steer / check_collision /
    get_nearest_node_index / find_near_nodes / choose_parent / rewire gives the tree the kernel gives."""
    import rrtk
    g, m = load_golden("rrt05_builtin_500")
    iters = 120
    stream = g["stream"][:iters]
    kw = dict(expand_dis=m["expand_dis"], goal_sample_rate=m["goal_sample_rate"], max_iter=iters, robot_radius=m["robot_radius"],
              connect_circle_dist=m["connect_circle_dist"], curvature=m["curvature"], goal_yaw_th=m["goal_yaw_th"],
              goal_xy_th=m["goal_xy_th"])
    fused = rrtk.RRTStarDubins(m["start"], m["goal"], m["obstacle_list"], m["rand_area"], **kw)
    fused.planning(animation=False, search_until_max_iter=True, sample_stream=stream)
    r = rrtk.RRTStarDubins(m["start"], m["goal"], m["obstacle_list"], m["rand_area"], **kw)
    r.node_list = [r.start]
    for i in range(iters):
        rnd = r.Node(*[float(v) for v in stream[i]])
        ni = r.get_nearest_node_index(r.node_list, rnd)
        new_node = r.steer(r.node_list[ni], rnd)
        if r.check_collision(new_node, r.obstacle_list, r.robot_radius):
            near = r.find_near_nodes(new_node)
            new_node = r.choose_parent(new_node, near)
            if new_node:
                r.node_list.append(new_node)
                r.rewire(new_node, near)
    t = fused.tree_arrays()
    assert len(r.node_list) == len(t["x"])
    idx = {id(n): i for i, n in enumerate(r.node_list)}
    assert [(-1 if n.parent is None else idx[id(n.parent)]) for n in r.node_list] == t["parent"].tolist()
    assert np.array_equal([n.x for n in r.node_list], t["x"]) and np.array_equal([n.yaw for n in r.node_list], t["yaw"])
    assert np.array_equal([n.cost for n in r.node_list], t["cost"])
