"""Host side of the samplers: the goal-bias coin and the uniform draws come from Python's `random`
exactly as the reference draws them (rrt_04:1132-1153), so a script that seeds `random` sees the same
stream; the Sobol points come from the device generator (rrtk_sobol_fill, the closed form of
rrt_04:230-503).  The result is the per-iteration sample stream the planning kernel consumes."""
from __future__ import annotations

import random

import numpy as np

from . import _lib


def sobol_points(dim: int, first: int, count: int) -> np.ndarray:
    """`[i4_sobol(dim, first + i)[0] for i in range(count)]` as an FP64 array [count, dim] (GPU)."""
    torch = _lib.require_cuda()
    out = torch.empty((max(count, 0), dim), dtype=torch.float64, device="cuda")
    if count > 0:
        _lib.check(_lib.lib().rrtk_sobol_fill_dev(dim, first, count, out.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream),
                   "rrtk_sobol_fill_dev")
    return out.cpu().numpy()


def draw_stream(max_iter, goal_xy, min_rand, max_rand, goal_sample_rate, sobol_sampler,
                sobol_first=0, rng=random):
    """Pre-draw `max_iter` samples the way `get_random_node[_sobol]` would, one per iteration.

    Returns (stream [max_iter, 2] float64, is_goal [max_iter] bool, next_sobol_index).
    Consumes `rng` draw-for-draw like the reference: one `randint(0, 100)` per iteration and, for the
    uniform sampler, two `uniform(min_rand, max_rand)` on non-goal iterations."""
    stream = np.empty((max_iter, 2), dtype=np.float64)
    is_goal = np.zeros(max_iter, dtype=bool)
    if sobol_sampler:
        for i in range(max_iter):
            is_goal[i] = not (rng.randint(0, 100) > goal_sample_rate)
        n_pts = int((~is_goal).sum())
        pts = sobol_points(2, sobol_first, n_pts)
        # rrt_04:1146-1147: min_rand + q * (max_rand - min_rand), numpy float64
        mapped = min_rand + pts * (max_rand - min_rand)
        stream[~is_goal] = mapped
        stream[is_goal] = (goal_xy[0], goal_xy[1])
        return stream, is_goal, sobol_first + n_pts
    for i in range(max_iter):
        if rng.randint(0, 100) > goal_sample_rate:
            stream[i, 0] = rng.uniform(min_rand, max_rand)
            stream[i, 1] = rng.uniform(min_rand, max_rand)
        else:
            is_goal[i] = True
            stream[i] = (goal_xy[0], goal_xy[1])
    return stream, is_goal, sobol_first


def consume_draws(iters, goal_sample_rate, sobol_sampler, min_rand, max_rand, rng=random):
    """Advance `rng` by exactly the draws `iters` iterations of the reference's loop make (rrt_04:1132-1153) and return the
    number of non-goal iterations among them (= how far `sobol_inter_` moves).  Used after an early exit: the stream is
    pre-drawn for max_iter iterations, the reference draws lazily, so the RNG is rewound and re-advanced to where the
    reference's would stand."""
    non_goal = 0
    for _ in range(iters):
        if rng.randint(0, 100) > goal_sample_rate:
            non_goal += 1
            if not sobol_sampler:
                rng.uniform(min_rand, max_rand)
                rng.uniform(min_rand, max_rand)
    return non_goal


# ---- the in-kernel counter-based sampler, restated on the host (tests, CPU baselines) ----
_M64 = (1 << 64) - 1


def _splitmix64(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def _rng_key(seed: int, q: int, it: int) -> int:
    inner = (q * 0x100000001B3 + it * 0x9E3779B1 + 0x51) & _M64
    return _splitmix64((seed ^ _splitmix64(inner)) & _M64)


def kernel_coins(seed: int, q: int, max_iter: int, goal_sample_rate: int) -> np.ndarray:
    """is_goal[it] of the in-kernel sampler (RRTK_SAMPLER_SOBOL / _UNIFORM) for query q."""
    out = np.zeros(max_iter, dtype=bool)
    for it in range(max_iter):
        coin = _splitmix64(_rng_key(seed, q, it)) % 101
        out[it] = not (coin > goal_sample_rate)
    return out
