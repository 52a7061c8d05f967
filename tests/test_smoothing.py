"""path_smoothing (rrt_04:1390-1479): oracle ports against fixtures made by the unmodified reference (CPU), and the
CUDA kernel against both (GPU)."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

NAMES = golden_names("smooth_")


def _obs(m):
    return [tuple(o) for o in m["obstacle_list"]]


@pytest.mark.parametrize("name", NAMES)
def test_python_port_matches_reference(name):
    import pyport
    g, m = load_golden(name)
    out = pyport.path_smoothing(g["path_in"].tolist(), g["draws"].tolist(), _obs(m))
    assert np.array_equal(np.array(out), g["path_out"])
    assert pyport.get_path_length(out) == float(g["length_out"])


@pytest.mark.parametrize("name", NAMES)
def test_c_oracle_matches_reference(name, oracle_lib):
    g, m = load_golden(name)
    out, rc = oracle_lib.path_smoothing(g["path_in"], g["draws"], m["obstacle_list"])
    assert rc == 0 and np.array_equal(np.array(out), g["path_out"])


def test_fixtures_cover_the_branches():
    lens = {n: (load_golden(n)[0]["path_in"].shape[0], load_golden(n)[0]["path_out"].shape[0]) for n in NAMES}
    assert any(a == b for a, b in lens.values()), "a case where every shortcut is rejected"
    assert any(b < a for a, b in lens.values())


def test_zero_division_is_reported(oracle_lib):
    """A zero-length pair under the pick: the reference raises ZeroDivisionError; the ports report it."""
    import pyport
    path = [[0.0, 0.0], [1.0, 0.0], [1.0, 0.0], [2.0, 0.0], [3.0, 1.0]]
    draws = [[0.30, 0.9]]   # 0.30 * le lands exactly at the end of the first pair or inside the zero-length one
    le = pyport.get_path_length(path)
    draws = [[1.0 / le, 0.9]]   # target = 1.0: reached at i = 0 -> fine; make it hit the zero pair instead
    path2 = [[0.0, 0.0], [0.0, 0.0], [1.0, 0.0], [2.0, 1.0]]
    with pytest.raises(ZeroDivisionError):
        pyport.path_smoothing(path2, [[0.0, 0.5]], [])
    out, rc = oracle_lib.path_smoothing(path2, [[0.0, 0.5]], [])
    assert rc == 1


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_reference_fixture(name):
    import rrtk
    g, m = load_golden(name)
    out = rrtk.path_smoothing(g["path_in"].tolist(), m["max_iter"], _obs(m), draws=g["draws"])
    assert np.array_equal(np.array(out), g["path_out"])
    assert rrtk.get_path_length(out) == float(g["length_out"])


@pytest.mark.gpu
def test_gpu_batch_matches_oracle_on_planner_output(oracle_lib):
    """Plan a batch, extract the courses on the device, smooth them there; compare every path with the C oracle."""
    import torch
    import rrtk
    from rrtk import smoothing, workloads as W
    cfg = W.C2
    Q, iters, n_obs, sm_iters = 48, 700, 48, 400
    qids = list(range(Q))
    rows = W.c2_rows(qids, n_obs)
    starts = np.tile(np.array(cfg["start"]), (Q, 1)); goals = np.tile(np.array(cfg["goal"]), (Q, 1))
    b = rrtk.RRTStarBatch(starts, goals, rows, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"],
                          cfg["goal_sample_rate"], iters, None, cfg["robot_radius"], "sobol", cfg["connect_circle_dist"],
                          True, seed=5, sobol_offset=np.asarray(qids, dtype=np.int64) * iters)
    res = b.run()
    cap = 64 + sm_iters
    path, plen = res.paths_device(cap)
    before = path.cpu().numpy(); n_before = plen.cpu().numpy()
    assert (n_before > 0).sum() >= Q // 2 and n_before.max() <= 64
    obs3 = torch.from_numpy(np.ascontiguousarray(rows[:, :, :3])).cuda()     # robot_radius = 0: column 2 is the size
    cnt = torch.full((Q,), n_obs, dtype=torch.int32, device="cuda")
    draws = np.random.default_rng(3).random((Q, sm_iters, 2))
    status, done = smoothing.smooth_batch(path, plen, sm_iters, obs3, cnt, torch.from_numpy(draws).cuda())
    after = path.cpu().numpy(); n_after = plen.cpu().numpy(); st = status.cpu().numpy()
    shortened = 0
    for q in range(Q):
        if n_before[q] == 0:
            assert n_after[q] == 0
            continue
        ref, rc = oracle_lib.path_smoothing(before[q, :n_before[q]], draws[q], rows[q, :, :3])
        assert (st[q] != 0) == (rc != 0)
        assert n_after[q] == len(ref) and np.array_equal(after[q, :n_after[q]], np.array(ref))
        shortened += n_after[q] < n_before[q]
    assert shortened >= 1


@pytest.mark.gpu
def test_gpu_zero_division_and_empty_paths():
    import rrtk
    with pytest.raises(ZeroDivisionError):
        rrtk.path_smoothing([[0.0, 0.0], [0.0, 0.0], [1.0, 0.0], [2.0, 1.0]], 1, [], draws=[[0.0, 0.5]])
    out = rrtk.path_smoothing([[0.0, 0.0], [1.0, 0.0], [2.0, 1.0]], 0, [(5.0, 5.0, 1.0)])
    assert out == [[0.0, 0.0], [1.0, 0.0], [2.0, 1.0]]
