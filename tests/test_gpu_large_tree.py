"""Large-tree mode (rrtk/large_tree.py): the HBM-streaming FP32 scans as a FILTER in front of the reference's FP64 decision.
The checker is the reference's own expression -- dlist = [(x - sx)**2 + (y - sy)**2 ...]; dlist.index(min(dlist)) and
[dlist.index(d) for d in dlist if d <= r**2] (rrt_04:1196-1202, :1314-1338) -- evaluated in FP64 with numpy on the CPU."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _ref_nearest(xy, s):
    d = (xy[:, 0] - s[0]) ** 2 + (xy[:, 1] - s[1]) ** 2
    return int(np.argmin(d)), d


def _ref_near(d, r2):
    hits = np.flatnonzero(d <= r2)
    first = {}
    for i in hits:                       # dist_list.index(v): first index holding the same value
        first.setdefault(d[i], i)
    return [int(first[d[i]]) for i in hits]


def test_decisions_equal_the_fp64_reference_where_fp32_cannot_tell():
    import rrtk
    rng = np.random.default_rng(17)
    n = 3_000_000
    xy = rng.uniform(-2.0, 15.0, (n, 2))
    c = np.array([6.25, 7.5])
    # nodes the FP32 mirror cannot tell apart: a ring of 200 nodes whose distances to c differ by ~1e-10, exact twins
    # (equal d2: the lowest index must win), and a node exactly on a query point
    ang = rng.uniform(0, 2 * np.pi, 200)
    ring = c + (0.003 + 1e-10 * rng.standard_normal(200))[:, None] * np.column_stack([np.cos(ang), np.sin(ang)])
    where = rng.choice(n, 200, replace=False)
    xy[where] = ring
    xy[123456] = xy[77]; xy[2_900_000] = xy[77]
    xy[500_000] = (3.0, 4.0)
    T = rrtk.LargeTree(n)
    T.extend(xy)
    queries = [c, xy[77], np.array([3.0, 4.0]), xy[where[5]]] + [rng.uniform(-2, 15, 2) for _ in range(12)]
    for s in queries:
        want, d = _ref_nearest(xy, s)
        assert T.nearest(float(s[0]), float(s[1])) == want
    # a sparse neighbourhood too: the band scales with the distance
    far = np.array([40.0, -30.0])
    assert T.nearest(*far) == _ref_nearest(xy, far)[0]
    for s, r in ((c, 0.003), (c, 0.0030000001), (xy[77], 0.02), (np.array([3.0, 4.0]), 0.0), (queries[6], 0.05), (queries[7], 0.2)):
        _, d = _ref_nearest(xy, s)
        assert T.near(float(s[0]), float(s[1]), float(r)) == _ref_near(d, r ** 2)
    st = T.stats
    assert st["nearest_queries"] == 17 and st["candidates"] >= st["nearest_queries"]
    # the filter is tight: away from the constructed ring a query re-checks a handful of nodes
    assert st["max_candidates"] < 5000


def test_rrt_over_the_large_tree_equals_the_fused_kernel():
    """Basic RRT grown node by node through LargeTree.nearest + rrtk_steer_collide_dev == the one-launch kernel (which is
    pinned to the reference): same nodes, same parents, same course."""
    import rrtk
    g, m = load_golden("rrt04_c1_uniform_500")
    iters = 400
    stream = g["stream"][:iters]
    a = rrtk.RRT(m["start"], m["goal"], m["obstacle_list"], m["rand_area"], m["expand_dis"], m["path_resolution"],
                 m["goal_sample_rate"], iters, m["play_area"], m["robot_radius"])
    pa = a.planning(animation=False, sample_stream=stream)
    b = rrtk.RRTLarge(m["start"], m["goal"], m["obstacle_list"], m["rand_area"], m["expand_dis"], m["path_resolution"],
                      m["goal_sample_rate"], iters, m["play_area"], m["robot_radius"])
    pb = b.planning(stream)
    ta = a.tree_arrays()
    n = len(ta["x"])
    assert b.tree.n == n and b.iters_done == a.iters_done
    assert np.array_equal(b.tree.xy64[:n].cpu().numpy(), np.column_stack([ta["x"], ta["y"]]))
    assert np.array_equal(b.parent[:n].cpu().numpy(), ta["parent"])
    assert pa == pb


def test_a_tree_past_twenty_million_nodes_scans_at_hbm_speed():
    """6.7 x 10^7 seeded nodes (far beyond the 126 MB L2), then the planner grows the tree: every get_nearest_node_index is two
    streaming passes over the FP32 mirror (8 B / node each), timed with CUDA events inside the planner loop."""
    import torch
    import rrtk
    n0, iters = 1 << 26, 40          # 67 M nodes: 537 MB per pass over the FP32 mirror, 1.6 GB of tree
    gen = torch.Generator(device="cuda").manual_seed(5)
    seed = torch.rand((n0, 2), dtype=torch.float64, device="cuda", generator=gen) * 17.0 - 2.0
    obs = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (7, 5, 2), (9, 5, 2)]
    parent = torch.arange(-1, n0 - 1, dtype=torch.int32, device="cuda")
    p = rrtk.RRTLarge([0.0, 0.0], [60.0, 60.0], obs, [-2, 15], 1.0, 0.1, 5, iters, None, 0.0, capacity=n0 + iters + 1,
                      seed_xy=seed, seed_parent=parent)
    rng = np.random.default_rng(9)
    stream = rng.uniform(-2, 15, (iters, 2))
    p.max_iter = 8
    p.planning(stream[:8])                       # warm-up
    p.tree.time_scans = True
    p.tree.stats.update(scans=0, scan_ms=0.0)
    n_before = p.tree.n
    p.max_iter = iters - 8
    p.planning(stream[8:])
    st = p.tree.stats
    assert p.tree.n > 20_000_000 and p.tree.n > n_before          # it grew past the seed
    gbs = 8.0 * n0 * st["scans"] / (st["scan_ms"] / 1e3) / 1e9
    print(f"large tree: {p.tree.n} nodes, {st['scans']} scans, {st['scan_ms'] / st['scans'] * 1e3:.1f} us per scan, {gbs:.0f} GB/s, "
          f"{st['candidates'] / max(st['nearest_queries'], 1):.1f} FP64 re-checks per query")
    assert gbs > 0.70 * 6531.6                                    # north_star: NN search >= 70 % of HBM bandwidth
    # the decisions are the FP64 reference's (numpy on the CPU over all 6.7 x 10^7 nodes)
    xy = p.tree.xy64[:p.tree.n].cpu().numpy()
    for s in stream[-3:]:
        assert p.tree.nearest(float(s[0]), float(s[1])) == _ref_nearest(xy, s)[0]
