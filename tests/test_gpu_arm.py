"""GPU tests of the arm C-space grid against the reference fixtures and the oracle."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_names

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names("arm02_"))
def test_grid_equals_reference_fixture(name):
    import rrtk
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = json.loads(str(g["meta"]))
    M = m["M"]
    want = np.unpackbits(g["grid_bits"])[:M * M].reshape(M, M)
    arm = rrtk.NLinkArm(m["link_length"], [0.0] * len(m["link_length"]))
    got = rrtk.get_occupancy_grid(arm, m["obstacles"], M)
    assert got.dtype == np.int64 and got.shape == (M, M)
    assert np.array_equal(got, want)
    assert int(got.sum()) == m["occupied"]


def test_survey_anchor_counts():
    """script arm + script obstacles: M=100 -> 5793, M=200 -> 23206; [1,1] arm M=100 -> 5010."""
    import rrtk
    obs = [[1.75, 0.75, 0.6], [0.55, 1.5, 0.5], [0, -1, 0.7], [0, -0.6, 0.4], [-1, 1., 0.3]]
    arm = rrtk.NLinkArm([0.5, 0.5, 0.3, 0.5, 0.1], [0.0] * 5)
    assert int(rrtk.get_occupancy_grid(arm, obs, 100).sum()) == 5793
    assert int(rrtk.get_occupancy_grid(arm, obs, 200).sum()) == 23206
    assert int(rrtk.get_occupancy_grid(rrtk.NLinkArm([1.0, 1.0], [0.0, 0.0]), obs, 100).sum()) == 5010


@pytest.mark.parametrize("M,n_links,S,O", [(33, 1, 3, 4), (128, 5, 8, 5), (257, 2, 5, 9), (96, 7, 2, 1), (64, 3, 4, 0)])
def test_batched_grids_vs_oracle(M, n_links, S, O, oracle_lib):
    from rrtk import arm as A
    Or = oracle_lib
    rng = np.random.default_rng(M * 7 + S)
    links = rng.uniform(0.1, 0.9, n_links)
    sets = np.concatenate([rng.uniform(-2, 2, (S, O, 2)), rng.uniform(0.1, 0.7, (S, O, 1))], axis=2)
    got = A.occupancy_grids_device(links, sets, M).cpu().numpy()
    for s in range(S):
        want = Or.arm_grid(M, links, sets[s], Or.MATH_CR) if O > 0 else np.zeros((M, M), np.uint8)
        assert np.array_equal(got[s], want), (s, int((got[s] != want).sum()))


def test_row_sharding_is_invariant(oracle_lib):
    """Rows computed as two shards (the multi-GPU partition) equal the whole grid."""
    from rrtk import arm as A
    rng = np.random.default_rng(3)
    links = [0.5, 0.5, 0.3, 0.5, 0.1]
    sets = np.concatenate([rng.uniform(-2, 2, (6, 5, 2)), rng.uniform(0.2, 0.7, (6, 5, 1))], axis=2)
    M = 200
    whole = A.occupancy_grids_device(links, sets, M).cpu().numpy()
    a = A.occupancy_grids_device(links, sets, M, 0, 77).cpu().numpy()
    b = A.occupancy_grids_device(links, sets, M, 77, M - 77).cpu().numpy()
    assert np.array_equal(np.concatenate([a, b], axis=1), whole)


def test_degenerate_zero_length_link(oracle_lib):
    """A zero-length link makes the reference's projection NaN, which it reports as a collision."""
    from rrtk import arm as A
    Or = oracle_lib
    links = [0.7, 0.0, 0.4]
    sets = np.array([[[1.0, 0.5, 0.3]]])
    got = A.occupancy_grids_device(links, sets, 40).cpu().numpy()[0]
    assert np.array_equal(got, Or.arm_grid(40, links, sets[0], Or.MATH_CR))
    assert got.all()


def _adversarial_sets(rng, link, S=12, O=5):
    """Random circles plus the cases the row rasteriser has to get right: circles tangent to the reach of links 2..n from
    some joint-1 position, joint 1 / the base exactly on a circle, radius 0, far away, everything inside, the far end of
    the arm reaching exactly the tangent point."""
    sets = np.concatenate([rng.uniform(-2.5, 2.5, (S, O, 2)), rng.uniform(0.05, 0.9, (S, O, 1))], axis=2)
    Ls, l0 = float(np.sum(link[1:])), link[0]
    ang = rng.uniform(-np.pi, np.pi)
    p1 = np.array([l0 * np.cos(ang), l0 * np.sin(ang)])
    u = rng.normal(size=2)
    u /= np.linalg.norm(u)
    sets[1, 0] = [*(p1 + u * (Ls + 0.4)), 0.4]
    sets[2, 0] = [*(p1 + u * (Ls + 0.4)), 0.4 * (1 + 1e-12)]
    sets[3, 0] = [*(p1 + u * 0.3), 0.3]
    sets[4, 0] = [0.4, 0.0, 0.4]
    sets[5, :, 2] = 0.0
    sets[6, :, :2] += 50.0
    sets[7, 0] = [0.0, 0.0, 10.0]
    sets[8, 0] = [*(p1 + u * np.sqrt(Ls * Ls + 0.25)), 0.5]
    return sets


@pytest.mark.parametrize("M", [64, 100, 101, 257, 1000, 2048])
@pytest.mark.parametrize("link", [[0.5, 0.5, 0.3, 0.5, 0.1], [1.0, 1.0], [0.7], [0.3, 1.2, 0.2]])
def test_row_rasteriser_equals_cell_by_cell(M, link):
    """rrtk_arm_grid_dev (runs of columns per circle, undecided cells evaluated exactly) against rrtk_arm_grid_cells_dev
    (every cell in the reference's order), bit for bit, on adversarial circles; odd M and M not a multiple of 32 take the
    unaligned store path and the partial last bitmap word."""
    from rrtk import arm as A
    rng = np.random.default_rng(M + len(link))
    sets = _adversarial_sets(rng, link)
    a = A.occupancy_grids_device(link, sets, M)
    b = A.occupancy_grids_device(link, sets, M, cell_by_cell=True)
    assert bool((a == b).all()), int((a != b).sum().item())


def test_row_rasteriser_fallbacks_and_shards():
    """A theta list that is not the reference's, a zero-length link, a row shard: still equal to the cell-by-cell grid."""
    from rrtk import arm as A
    rng = np.random.default_rng(9)
    link, M = [0.5, 0.5, 0.3, 0.5, 0.1], 500
    sets = _adversarial_sets(rng, link)
    th = A.theta_list(M).copy()
    th[250] += 1e-3
    for kw in (dict(theta=th), dict(row0=123, n_rows=200)):
        a = A.occupancy_grids_device(link, sets, M, **kw)
        b = A.occupancy_grids_device(link, sets, M, cell_by_cell=True, **kw)
        assert bool((a == b).all()), kw.keys()
    a = A.occupancy_grids_device([0.5, 0.0, 0.3], sets, M)
    b = A.occupancy_grids_device([0.5, 0.0, 0.3], sets, M, cell_by_cell=True)
    assert bool((a == b).all())


def test_row_rasteriser_full_size_rows_equal_cell_by_cell():
    """Config 5's size (M = 8192, the script's arm, 64 obstacle sets): 512 rows of it, both ways."""
    from rrtk import arm as A
    rng = np.random.default_rng(5)
    S, M = 64, 8192
    sets = np.concatenate([rng.uniform(-2, 2, (S, 5, 2)), rng.uniform(0.2, 0.7, (S, 5, 1))], axis=2)
    sets[0] = [[1.75, 0.75, 0.6], [0.55, 1.5, 0.5], [0, -1, 0.7], [0, -0.6, 0.4], [-1, 1., 0.3]]
    link = [0.5, 0.5, 0.3, 0.5, 0.1]
    for row0 in (0, 3840, 7680):
        a = A.occupancy_grids_device(link, sets, M, row0, 512)
        b = A.occupancy_grids_device(link, sets, M, row0, 512, cell_by_cell=True)
        assert bool((a == b).all()), row0


def test_arm_grid_argument_checks_and_empty_inputs():
    """Both entry points: bad row ranges / link counts / NULL pointers are RRTK_ERR_INVALID with a message; zero rows or
    zero sets is a no-op; zero circles per set gives an all-free grid; M = 1 works."""
    import torch
    from rrtk import _lib, arm as A
    L = _lib.lib()
    dev = torch.device("cuda")
    s = torch.cuda.current_stream().cuda_stream
    M = 48
    theta = torch.from_numpy(A.theta_list(M)).to(dev)
    link = np.array([0.6, 0.5], dtype=np.float64)
    obs = torch.from_numpy(np.array([[[1.0, 0.2, 0.3]]], dtype=np.float64)).to(dev)
    grid = torch.full((1, M, M), 9, dtype=torch.uint8, device=dev)
    for fn in (L.rrtk_arm_grid_dev, L.rrtk_arm_grid_cells_dev):
        call = lambda M_=M, th=theta.data_ptr(), r0=0, nr=M, nl=2, lk=link.ctypes.data, ob=obs.data_ptr(), S=1, O=1, g=grid.data_ptr(): \
            fn(M_, th, r0, nr, nl, lk, ob, S, O, g, s)  # noqa: E731
        assert call() == 0
        assert call(M_=0) < 0 and b"M" in L.rrtk_last_error()
        assert call(r0=40, nr=20) < 0                      # rows past the grid
        assert call(nl=0) < 0 and call(nl=17) < 0          # 1 <= n_links <= 16
        assert call(th=None) < 0 and call(g=None) < 0 and call(ob=None) < 0 and call(lk=None) < 0
        assert call(S=-1) < 0 and call(O=-1) < 0
        grid.fill_(9)
        assert call(nr=0) == 0 and call(S=0) == 0          # nothing to do, nothing written
        torch.cuda.synchronize()
        assert bool((grid == 9).all())
        assert call(O=0, ob=None) == 0                     # no circles: every cell free
        torch.cuda.synchronize()
        assert bool((grid == 0).all())
    one = A.occupancy_grids_device([0.5, 0.5], np.array([[[0.0, 0.0, 0.1]]]), 1)
    assert one.shape == (1, 1, 1) and int(one.item()) == 1  # the base sits inside the circle


def test_negative_radius_touches_nothing(oracle_lib):
    """detect_collision returns False whenever dist > radius (arm02:72): a circle of negative radius never collides, in
    either kernel (the squared-radius filter must not see it as |r|)."""
    from rrtk import arm as A
    Or = oracle_lib
    links = [0.5, 0.5, 0.3]
    sets = np.array([[[0.3, 0.2, -0.5], [1.0, 0.4, 0.3]], [[0.1, 0.1, -0.2], [0.0, 0.0, -1.0]]])
    for cbc in (False, True):
        got = A.occupancy_grids_device(links, sets, 40, cell_by_cell=cbc).cpu().numpy()
        for s in range(2):
            assert np.array_equal(got[s], Or.arm_grid(40, links, sets[s], Or.MATH_CR)), (cbc, s)
        assert int(got[1].sum()) == 0
