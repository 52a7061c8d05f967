"""Reeds-Shepp steering (rs00:73-515): oracle ports against a fixture made by the unmodified reference (CPU), the CUDA
kernel against the oracle (bit-exact in `cr` mode) and against the reference fixture (GPU)."""
import math

import numpy as np
import pytest

from conftest import load_golden


def _cases():
    g, _ = load_golden("rs_pairs_400")
    off = np.concatenate([[0], np.cumsum(g["n_pts"])])
    return g, off


def test_python_port_matches_reference():
    import pyport
    g, off = _cases()
    for k in range(0, 400, 3):
        r = pyport.reeds_shepp_path_planning(*g["start"][k], *g["goal"][k], float(g["maxc"][k]), float(g["step"][k]))
        if g["n_pts"][k] == 0:
            assert r[0] is None
            continue
        m = int((g["types"][k] >= 0).sum())
        assert ["LSR".index(c) for c in r[3]] == g["types"][k, :m].tolist()
        assert r[4] == g["lengths"][k, :m].tolist()
        sl = slice(off[k], off[k + 1])
        assert r[0] == g["x"][sl].tolist() and r[1] == g["y"][sl].tolist() and r[2] == g["yaw"][sl].tolist()


def test_c_oracle_libm_matches_reference_bitwise(oracle_lib):
    O = oracle_lib
    g, off = _cases()
    for k in range(400):
        r = O.reeds_shepp(g["start"][k], g["goal"][k], float(g["maxc"][k]), float(g["step"][k]), O.MATH_LIBM)
        if g["n_pts"][k] == 0:
            assert r is None
            continue
        m = int((g["types"][k] >= 0).sum())
        assert ["LSR".index(c) for c in r["types"]] == g["types"][k, :m].tolist() and r["lengths"] == g["lengths"][k, :m].tolist()
        sl = slice(off[k], off[k + 1])
        assert np.array_equal(r["pts"][:, 0], g["x"][sl]) and np.array_equal(r["pts"][:, 1], g["y"][sl])
        assert np.array_equal(r["pts"][:, 2], g["yaw"][sl])


def test_c_oracle_cr_mode_agrees_with_reference(oracle_lib):
    """Correctly rounded leaf functions: same word, same point count, points within 1e-9 of the reference's."""
    O = oracle_lib
    g, off = _cases()
    same_bits = 0
    for k in range(400):
        r = O.reeds_shepp(g["start"][k], g["goal"][k], float(g["maxc"][k]), float(g["step"][k]), O.MATH_CR)
        if g["n_pts"][k] == 0:
            assert r is None
            continue
        m = int((g["types"][k] >= 0).sum())
        assert ["LSR".index(c) for c in r["types"]] == g["types"][k, :m].tolist() and r["n"] == g["n_pts"][k]
        sl = slice(off[k], off[k + 1])
        assert np.allclose(r["pts"][:, 0], g["x"][sl], atol=1e-9, rtol=0) and np.allclose(r["pts"][:, 1], g["y"][sl], atol=1e-9, rtol=0)
        same_bits += np.array_equal(r["pts"][:, 0], g["x"][sl])
    assert same_bits > 200


def test_asin_correctly_rounded_identities(oracle_lib):
    L = oracle_lib.lib()
    assert L.orc_cr_asin(0.0) == 0.0 and L.orc_cr_asin(1.0) == math.pi / 2 and L.orc_cr_asin(-1.0) == -math.pi / 2
    assert L.orc_cr_asin(0.5) == 0.5235987755982989 and math.isnan(L.orc_cr_asin(1.5))
    xs = np.random.default_rng(2).uniform(-1, 1, 20000)
    assert sum(L.orc_cr_asin(float(x)) != math.asin(float(x)) for x in xs) < 40   # glibc asin is correctly rounded ~always


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_bitwise_vs_oracle_cr_and_reference_fixture(oracle_lib):
    from rrtk import reeds_shepp as RS
    O = oracle_lib
    g, off = _cases()
    for maxc in (1.0, 0.5, 2.0, 0.1):
        for step in (0.2, 0.1, 0.05):
            idx = [k for k in range(400) if g["maxc"][k] == maxc and g["step"][k] == step]
            if not idx:
                continue
            out = RS.steer_batch(g["start"][idx], g["goal"][idx], maxc, step, max_pts=1200)
            ty, ln, npts, pts, npaths = (out[k].cpu().numpy() for k in ("types", "lengths", "n_pts", "pts", "n_paths"))
            for j, k in enumerate(idx):
                ref = O.reeds_shepp(g["start"][k], g["goal"][k], maxc, step, O.MATH_CR)
                if ref is None:
                    assert npts[j] == 0 and (ty[j] == -1).all() and g["n_pts"][k] == 0
                    continue
                m = len(ref["types"])
                assert ty[j, :m].tolist() == ["LSR".index(c) for c in ref["types"]] and (ty[j, m:] == -1).all()
                assert ln[j, :m].tolist() == ref["lengths"] and npts[j] == ref["n"] and npaths[j] == ref["n_paths"]
                assert np.array_equal(pts[j, :npts[j]], ref["pts"])                      # bit-exact vs oracle[cr]
                sl = slice(off[k], off[k + 1])                                           # and the reference itself
                assert ty[j, :m].tolist() == g["types"][k, :m].tolist() and npts[j] == g["n_pts"][k]
                assert np.allclose(pts[j, :npts[j], 0], g["x"][sl], atol=1e-9, rtol=0)
                assert np.allclose(pts[j, :npts[j], 1], g["y"][sl], atol=1e-9, rtol=0)


@pytest.mark.gpu
def test_gpu_drop_in_function_and_collision_flag(oracle_lib):
    import rrtk
    from rrtk import reeds_shepp as RS
    x, y, yaw, modes, lengths = rrtk.reeds_shepp_path_planning(-1.0, -4.0, np.deg2rad(-20.0), 5.0, 5.0, np.deg2rad(25.0), 0.1, 0.05)
    ref = oracle_lib.reeds_shepp([-1.0, -4.0, np.deg2rad(-20.0)], [5.0, 5.0, np.deg2rad(25.0)], 0.1, 0.05, oracle_lib.MATH_CR)
    assert modes == ref["types"] and lengths == ref["lengths"] and x == ref["pts"][:, 0].tolist() and yaw == ref["pts"][:, 2].tolist()
    same = rrtk.reeds_shepp_path_planning(0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0)             # identical poses: a closed loop
    ref0 = oracle_lib.reeds_shepp([0.0, 0.0, 0.0], [0.0, 0.0, 0.0], 1.0, 0.2, oracle_lib.MATH_CR)
    assert same[3] == ref0["types"] and same[0] == ref0["pts"][:, 0].tolist()
    # "Step size too large for Reeds-Shepp paths." -> the reference returns five Nones
    assert oracle_lib.reeds_shepp([0.0, 0.0, 0.0], [0.3, 0.0, 0.0], 1.0, 0.35, oracle_lib.MATH_CR) is None
    assert rrtk.reeds_shepp_path_planning(0.0, 0.0, 0.0, 0.3, 0.0, 0.0, 1.0, 0.35) == (None,) * 5
    # collision flag: a circle on the course blocks it, a far one does not
    mid = ref["pts"][len(ref["pts"]) // 2]
    out = RS.steer_batch([[-1.0, -4.0, np.deg2rad(-20.0)]] * 2, [[5.0, 5.0, np.deg2rad(25.0)]] * 2, 0.1, 0.05,
                         obstacle_sets=[[(float(mid[0]), float(mid[1]), 0.3)], [(50.0, 50.0, 1.0)]], obs_set=[0, 1])
    assert out["free"].cpu().numpy().tolist() == [0, 1]


@pytest.mark.gpu
def test_gpu_asin_probe(oracle_lib):
    import torch
    from rrtk import _lib
    x = np.concatenate([np.random.default_rng(3).uniform(-1, 1, 4000), [0.0, 1.0, -1.0, 0.5, 1e-300, 0.9999999999999999]])
    d_x = torch.from_numpy(x).cuda(); d_o = torch.empty_like(d_x)
    _lib.check(_lib.lib().rrtk_crmath_probe_dev(7, len(x), d_x.data_ptr(), None, d_o.data_ptr(), None))
    want = np.array([oracle_lib.lib().orc_cr_asin(float(v)) for v in x])
    assert np.array_equal(d_o.cpu().numpy(), want)
