"""astar_torus (arm02:113-233): oracle ports against fixtures made by the unmodified reference (CPU), and the CUDA
kernels against both (GPU)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_names, load_golden

NAMES = golden_names("astar_")


def _grid(m):
    src = np.load(os.path.join(GOLDEN, m["source"] + ".npz"))
    M = m["M"]
    return np.unpackbits(src["grid_bits"])[:M * M].reshape(M, M)


@pytest.mark.parametrize("name", NAMES)
def test_python_port_matches_reference(name):
    import pyport
    g, m = load_golden(name)
    grid = _grid(m).astype(np.int64).tolist()
    route = pyport.astar_torus(grid, m["start"], m["goal"])
    assert np.array_equal(np.array(route, dtype=np.int64).reshape(-1, 2), g["route"])
    assert np.array_equal(np.array(grid), g["grid_after"])
    assert np.array_equal(np.array(pyport.calc_heuristic_map(m["M"], m["goal"])), g["heuristic"])


@pytest.mark.parametrize("name", NAMES)
def test_c_oracle_matches_reference(name, oracle_lib):
    g, m = load_golden(name)
    route, after = oracle_lib.astar_torus(_grid(m), m["start"], m["goal"])
    assert np.array_equal(route, g["route"]) and np.array_equal(after, g["grid_after"])
    assert np.array_equal(oracle_lib.astar_heuristic(m["M"], m["goal"]), g["heuristic"])


def test_heuristic_closed_form_equals_in_place_loop(oracle_lib):
    """The closed form the GPU evaluates per cell against the sequential in-place loop, small and odd sizes."""
    for M, goal in ((1, (0, 0)), (2, (1, 0)), (3, (0, 2)), (7, (6, 6)), (16, (0, 0)), (33, (20, 5))):
        assert np.array_equal(oracle_lib.astar_heuristic_closed(M, goal), oracle_lib.astar_heuristic(M, goal)), (M, goal)


def test_fixtures_cover_the_branches():
    metas = {n: load_golden(n) for n in NAMES}
    assert any(g["route"].shape[0] == 0 for g, _ in metas.values()), "an unreachable goal"
    assert any(g["route"].shape[0] == 1 for g, _ in metas.values()), "start == goal"
    wraps = 0
    for g, m in metas.values():
        r = g["route"]
        wraps += int((np.abs(np.diff(r, axis=0)).max(axis=1) > 1).sum()) if len(r) > 1 else 0
    assert wraps >= 2, "routes that cross the torus seam"


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_reference_fixture(name):
    import rrtk
    g, m = load_golden(name)
    grid = _grid(m).astype(np.int64)
    route = rrtk.astar_torus(grid, tuple(m["start"]), tuple(m["goal"]))
    assert np.array_equal(np.array(route, dtype=np.int64).reshape(-1, 2), g["route"])
    assert np.array_equal(grid, g["grid_after"])


@pytest.mark.gpu
def test_gpu_grid_to_route_pipeline_matches_oracle(oracle_lib):
    """Occupancy grids of 12 obstacle sets computed on the GPU, 3 start/goal pairs each searched on the GPU; every
    route and final grid against the C oracle."""
    import torch
    from rrtk import arm as A
    M, S = 96, 12
    rng = np.random.default_rng(12)
    ang, rad = rng.uniform(0, 2 * np.pi, (S, 5)), rng.uniform(0.9, 2.0, (S, 5))   # circles that leave the base free
    sets = np.stack([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(0.15, 0.45, (S, 5))], axis=2)
    grids = A.occupancy_grids_device([1.0, 1.0], sets, M)
    host = grids.cpu().numpy()
    starts, goals, which = [], [], []
    for s in range(S):
        free = np.argwhere(host[s] == 0)
        for _ in range(3):
            a, b = free[rng.integers(len(free))], free[rng.integers(len(free))]
            starts.append(a); goals.append(b); which.append(s)
    work = grids[torch.tensor(which, device=grids.device)].contiguous()
    routes, rlen, expanded = A.astar_torus_batch(work, np.array(starts), np.array(goals))
    routes, rlen, after = routes.cpu().numpy(), rlen.cpu().numpy(), work.cpu().numpy()
    found = 0
    for k, s in enumerate(which):
        ref_route, ref_after = oracle_lib.astar_torus(host[s], starts[k], goals[k])
        assert rlen[k] == len(ref_route)
        assert np.array_equal(routes[k, :rlen[k]], ref_route) and np.array_equal(after[k], ref_after)
        found += rlen[k] > 0
    assert found >= len(which) // 2 and (expanded.cpu().numpy() > 0).any()


@pytest.mark.gpu
def test_gpu_rejects_cells_outside_the_grid():
    import torch
    from rrtk import arm as A
    g = torch.zeros((1, 8, 8), dtype=torch.uint8, device="cuda")
    with pytest.raises(IndexError):
        A.astar_torus_batch(g, [[0, 8]], [[1, 1]])
