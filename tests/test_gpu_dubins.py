"""GPU tests of batched Dubins steering against the reference fixture and the oracle."""
import math

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _fixture():
    g, _ = load_golden("dubins_pairs_150")
    return g["cases"], g["pts"], g["offsets"]


def test_steer_batch_bitwise_vs_oracle_cr(oracle_lib):
    from rrtk import dubins
    O = oracle_lib
    cases, _, _ = _fixture()
    for kappa in (1.0, 0.5, 2.0):
        sel = cases[cases[:, 6] == kappa]
        r = dubins.steer_batch(sel[:, 0:3], sel[:, 3:6], kappa, 0.1, max_pts=1200)
        for i, row in enumerate(sel):
            ref = O.dubins_plan(row[0:3], row[3:6], kappa, 0.1, O.MATH_CR)
            assert r["mode"][i] == ref["mode"] and r["n_pts"][i] == ref["n"]
            assert np.array_equal(r["lengths"][i], ref["lengths"])
            assert np.array_equal(r["pts"][i, :ref["n"]], ref["pts"])
            assert np.array_equal(r["end"][i], ref["pts"][-1])


def test_steer_batch_vs_reference_fixture():
    """Against the unmodified reference: same word, same number of course points, lengths and points equal
    to 1e-12 (the reference's libm differs from correctly rounded results by an ulp here and there)."""
    from rrtk import dubins
    cases, pts, off = _fixture()
    n_bit = 0
    for kappa in (1.0, 0.5, 2.0):
        idx = np.flatnonzero(cases[:, 6] == kappa)
        sel = cases[idx]
        r = dubins.steer_batch(sel[:, 0:3], sel[:, 3:6], kappa, 0.1, max_pts=1200)
        for i, (ci, row) in enumerate(zip(idx, sel)):
            ref_pts = pts[off[ci]:off[ci + 1]]
            assert r["mode"][i] == int(row[7]) and r["n_pts"][i] == int(row[11]) == len(ref_pts)
            assert np.allclose(r["lengths"][i], row[8:11], rtol=1e-13, atol=1e-13)
            got = r["pts"][i, :len(ref_pts)]
            assert np.allclose(got, ref_pts, rtol=0, atol=1e-12)
            n_bit += np.array_equal(got, ref_pts)
    assert n_bit >= len(cases) // 2     # most edges are bit-identical to the reference's


def test_docstring_example():
    """rrt_05:1067-1075: (1, 1, 45 deg) -> (-3, -3, -45 deg), curvature 1: LSL, 94 points."""
    import rrtk
    x, y, yaw, mode, lengths = rrtk.plan_dubins_path(1.0, 1.0, math.radians(45.0), -3.0, -3.0, math.radians(-45.0), 1.0)
    assert "".join(mode) == "LSL" and len(x) == 94
    assert np.allclose(lengths, [3.353118, 4.763013, 1.359271], atol=1e-6)
    assert abs(x[-1] + 3) < 1e-12 and abs(y[-1] + 3) < 1e-12 and abs(yaw[-1] + math.pi / 4) < 1e-12


def test_collision_verdicts_vs_oracle(oracle_lib):
    """check_collision over the sampled course (rrt_05:1625-1638) for edges among random circles."""
    from rrtk import dubins
    O = oracle_lib
    rng = np.random.default_rng(31)
    n = 300
    f = np.column_stack([rng.uniform(0, 12, (n, 2)), rng.uniform(-math.pi, math.pi, n)])
    t = np.column_stack([f[:, 0:2] + rng.uniform(-4, 4, (n, 2)), rng.uniform(-math.pi, math.pi, n)])
    sets = [[(float(x), float(y), float(r)) for (x, y), r in zip(rng.uniform(0, 12, (k, 2)), rng.uniform(0.2, 0.8, k))]
            for k in (6, 25, 0)]
    sid = rng.integers(0, 3, n).astype(np.int32)
    rr = 0.2
    r = dubins.steer_batch(f, t, 1.0, 0.1, obstacle_sets=sets, obs_set=sid, robot_radius=rr)
    n_hit = 0
    for i in range(n):
        ref = O.dubins_plan(f[i], t[i], 1.0, 0.1, O.MATH_CR)
        ok = True
        for (ox, oy, size) in sets[sid[i]]:
            d2 = (ox - ref["pts"][:, 0]) * (ox - ref["pts"][:, 0]) + (oy - ref["pts"][:, 1]) * (oy - ref["pts"][:, 1])
            if d2.min() <= (size + rr) ** 2:
                ok = False
                break
        assert bool(r["free"][i]) == ok, i
        n_hit += not ok
    assert 20 < n_hit < n - 20


def test_degenerate_same_pose_and_acos(oracle_lib):
    from rrtk import dubins, _lib
    import torch
    r = dubins.steer_batch([[1.0, 2.0, 0.3]], [[1.0, 2.0, 0.3]], 1.0, 0.1, max_pts=8)
    ref = oracle_lib.dubins_plan([1.0, 2.0, 0.3], [1.0, 2.0, 0.3], 1.0, 0.1, oracle_lib.MATH_CR)
    assert r["n_pts"][0] == ref["n"] and r["mode"][0] == ref["mode"]
    # correctly rounded acos on the device (probe kind 6) against the CPU build of the same header
    x = np.concatenate([np.random.default_rng(1).uniform(-1, 1, 4000), [1.0, -1.0, 0.0, 1 - 1e-16, -1 + 1e-16]])
    tx = torch.from_numpy(x).cuda()
    out = torch.empty_like(tx)
    _lib.check(_lib.lib().rrtk_crmath_probe_dev(6, tx.numel(), tx.data_ptr(), None, out.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
    want = np.array([oracle_lib.lib().orc_cr_acos(float(v)) for v in x])
    assert np.array_equal(out.cpu().numpy(), want)


# ---- RRT*-Dubins planning loop (rrt_05:1416-1779) ----
from conftest import golden_names  # noqa: E402

RRT05 = golden_names("rrt05_")


def _dub_planner(m):
    import rrtk
    return rrtk.RRTStarDubins(m["start"], m["goal"], m["obstacle_list"], m["rand_area"], m["expand_dis"],
                              goal_sample_rate=m["goal_sample_rate"], max_iter=m["max_iter"],
                              robot_radius=m["robot_radius"], connect_circle_dist=m["connect_circle_dist"],
                              curvature=m["curvature"], goal_yaw_th=m["goal_yaw_th"], goal_xy_th=m["goal_xy_th"])


@pytest.mark.parametrize("name", RRT05)
def test_rrtstar_dubins_bitwise_vs_oracle_cr(name, oracle_lib):
    O = oracle_lib
    g, m = load_golden(name)
    ref = O.rrtstar_dubins_run(m["start"], m["goal"], m["obstacle_list"], m["expand_dis"], m["max_iter"],
                               m["robot_radius"], m["connect_circle_dist"], m["curvature"], m["goal_yaw_th"],
                               m["goal_xy_th"], m["search_until_max_iter"], g["stream"], O.MATH_CR)
    r = _dub_planner(m)
    path = r.planning(animation=False, search_until_max_iter=m["search_until_max_iter"], sample_stream=g["stream"])
    t = r.tree_arrays()
    assert t["n"] == ref["n"] and t["iters_done"] == ref["iters_done"] and t["goal_index"] == ref["goal_index"]
    assert np.array_equal(t["parent"], ref["parent"])
    for k in ("x", "y", "yaw", "cost"):
        assert np.array_equal(t[k], ref[k]), k
    assert np.array_equal(t["edge_from"][1:], ref["edge_from"][1:]) and np.array_equal(t["edge_to"][1:], ref["edge_to"][1:])
    assert (path is None) == (ref["path"] is None)
    if path is not None:
        assert path == ref["path"]


@pytest.mark.parametrize("name", RRT05)
def test_rrtstar_dubins_vs_reference_fixture(name):
    """Against the unmodified reference: identical node count, parents and path length; poses, costs and path
    points within 1e-9 (Dubins poses carry ulp-level libm differences; north_star allows 1e-5 relative)."""
    g, m = load_golden(name)
    r = _dub_planner(m)
    path = r.planning(animation=False, search_until_max_iter=m["search_until_max_iter"], sample_stream=g["stream"])
    t = r.tree_arrays()
    assert t["n"] == len(g["x"]) and np.array_equal(t["parent"], g["parent"])
    for k in ("x", "y", "yaw", "cost"):
        assert np.allclose(t[k], g[k], rtol=0, atol=1e-9), k
    if len(g["path"]) == 0:
        assert path is None
    else:
        assert len(path) == len(g["path"]) and np.allclose(np.array(path), g["path"], rtol=0, atol=1e-9)
        # Node objects expose the sampled edges like the reference's
        nd = r.node_list[t["goal_index"]]
        assert len(nd.path_x) == len(nd.path_y) == len(nd.path_yaw) > 1 and nd.parent is not None
