"""CPU tests of the C-ABI boundary: the library loads without a GPU, exports every symbol that
include/rrtk.h declares, and rejects bad arguments with a status + message (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "rrtk.h")).read()
    return sorted(set(re.findall(r"RRTK_API\s+(?:const\s+)?\w+\s*\*?\s*(rrtk_\w+)\s*\(", src)))


def test_header_symbols_exported():
    import rrtk
    from rrtk import _lib
    handle = C.CDLL(rrtk.LIB_PATH)
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/rrtk.h but not exported"
    # the ctypes binding covers every declared symbol
    assert sorted(_lib.EXPORTED) == names


def test_version_and_error_string():
    import rrtk
    L = rrtk.lib()
    assert L.rrtk_version() == 100
    assert L.rrtk_sobol_table(0, None) == -1
    assert b"dim" in L.rrtk_last_error()
    assert L.rrtk_sobol_fill_dev(41, 0, 10, None, None) == -1
    assert L.rrtk_crmath_probe_dev(9, 1, None, None, None, None) == -1


def test_params_struct_layout_matches_header():
    """sizeof(rrtk_rrtstar_params) = 10 int32 + 8 double + uint64 + 2 int32 + 3 double + 2 int32 = 152 bytes on LP64."""
    from rrtk import _lib
    assert C.sizeof(_lib.RRTStarParams) == 168


def test_every_struct_layout_matches_the_compiled_library():
    from rrtk import _lib
    L = _lib.lib()
    for which, cls in enumerate((_lib.RRTStarParams, _lib.InformedParams, _lib.InformedTreeParams, _lib.InformedTreeResult,
                                 _lib.DubinsParams, _lib.ClosedLoopParams, _lib.BitStarParams)):
        assert L.rrtk_sizeof(which) == C.sizeof(cls), cls.__name__
    assert L.rrtk_sizeof(99) == -1


def test_bad_params_rejected_without_gpu():
    from rrtk import _lib, engine
    L = _lib.lib()
    p = engine.make_params(4, 10, 11, 1, 1.0, 0.1, near_cap=33)
    rc = L.rrtk_rrtstar_run_dev(C.byref(p), *([None] * 16))
    assert rc == -1 and b"near_cap" in L.rrtk_last_error()
    p = engine.make_params(4, 10, 11, 1, 1.0, 0.0)
    assert L.rrtk_rrtstar_run_dev(C.byref(p), *([None] * 16)) == -1
    p = engine.make_params(4, 10, 11, 1, 1.0, 0.1)
    assert L.rrtk_rrtstar_run_dev(C.byref(p), *([None] * 16)) == -1
    assert b"NULL" in L.rrtk_last_error()
    p = engine.make_params(0, 10, 11, 1, 1.0, 0.1)
    assert L.rrtk_rrtstar_run_dev(C.byref(p), *([None] * 16)) == 0   # empty batch is a no-op


def test_sobol_table_matches_oracle(oracle_lib):
    import rrtk
    v = np.zeros((40, 30), dtype=np.uint32)
    assert rrtk.lib().rrtk_sobol_table(40, v.ctypes.data) == 0
    assert np.array_equal(v, oracle_lib.sobol_table(40))


def test_no_cpu_fallback():
    """Without a GPU every compute entry point of the Python API raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rrtk
    r = rrtk.RRTStar([0, 0], [6.0, 10.0], [(5, 5, 1)], [-2, 15], max_iter=10)
    with pytest.raises(rrtk.RrtkError):
        r.planning(animation=False)


def test_host_sampler_consumes_random_like_the_reference():
    """draw_stream draws randint/uniform in the reference's order (rrt_04:1132-1139); checked on the
    uniform fixture whose stream the reference produced from random.seed(3)."""
    import random
    from conftest import load_golden
    from rrtk import sampling
    g, m = load_golden("rrt04_c1_uniform_500")
    rng = random.Random(m["seed"])
    stream, is_goal, _ = sampling.draw_stream(m["max_iter"], m["goal"], m["rand_area"][0], m["rand_area"][1],
                                              m["goal_sample_rate"], False, 0, rng)
    assert np.array_equal(stream, g["stream"])


def test_shard_range_partitions():
    from rrtk import shard_range
    for n in (0, 1, 7, 4096, 1000):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_new_entry_points_reject_bad_arguments_without_gpu():
    """Argument validation of the round-1 additions happens before any CUDA call."""
    from rrtk import _lib
    L = _lib.lib()
    p = _lib.ClosedLoopParams()
    p.n_courses, p.course_cap, p.traj_cap = 4, 2, 100                      # extend_path needs 3 course points
    assert L.rrtk_closed_loop_dev(C.byref(p), *([None] * 10)) == -1 and b"course_cap" in L.rrtk_last_error()
    p.course_cap = 16
    assert L.rrtk_closed_loop_dev(C.byref(p), *([None] * 10)) == -1 and b"NULL" in L.rrtk_last_error()
    p.n_courses = 0
    assert L.rrtk_closed_loop_dev(C.byref(p), *([None] * 10)) == 0
    b = _lib.BitStarParams()
    b.n_queries, b.max_iter, b.vertex_cap, b.sample_cap, b.edge_cap, b.path_cap, b.n_draws = 2, 10, 12, 64, 64, 14, 100
    b.min_rand, b.max_rand, b.num_cells = 15.0, -2.0, 1700.0               # inverted randArea
    assert L.rrtk_bitstar_run_dev(C.byref(b), *([None] * 12)) == -1 and b"randArea" in L.rrtk_last_error()
    b.min_rand, b.max_rand = -2.0, 15.0
    assert L.rrtk_bitstar_run_dev(C.byref(b), *([None] * 12)) == -1 and b"NULL" in L.rrtk_last_error()
    assert L.rrtk_steer_collide_dev(4, None, None, 1.0, 0.0, None, None, 0, *([None] * 8)) == -1 and b"path_resolution" in L.rrtk_last_error()
    assert L.rrtk_steer_collide_dev(0, None, None, 1.0, 0.1, None, None, 0, *([None] * 8)) == 0
    d = _lib.DubinsParams()
    d.n_queries, d.max_iter, d.node_cap, d.near_cap, d.curvature, d.step_size, d.rs_cost = 1, 5, 11, 32, 1.0, 0.2, 7
    assert L.rrtk_rrtstar_rs_run_dev(C.byref(d), *([1] * 16), None) == -1 and b"rs_cost" in L.rrtk_last_error()
    r = _lib.RRTStarParams()
    r.n_queries, r.max_iter, r.node_cap, r.near_cap, r.path_resolution, r.expand_dis, r.resume = 1, 5, 6, 32, 0.1, 1.0, 3
    assert L.rrtk_rrtstar_run_dev(C.byref(r), *([None] * 16)) == -1 and b"resume" in L.rrtk_last_error()
