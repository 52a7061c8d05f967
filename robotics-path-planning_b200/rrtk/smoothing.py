"""Path smoothing (rrt_04:1390-1479) behind the reference's function names, computed on the GPU.

`path_smoothing(path, max_iter, obstacle_list)` is the reference's signature; `smooth_batch` takes device tensors
straight from `BatchResult.paths_device()` so the courses never leave the GPU."""
from __future__ import annotations

import math
import random

import numpy as np

from . import _lib


def get_path_length(path):
    """rrt_04:1390-1398 (host helper, same arithmetic)."""
    le = 0
    for i in range(len(path) - 1):
        le += math.hypot(path[i + 1][0] - path[i][0], path[i + 1][1] - path[i][1])
    return le


def smooth_batch(path, path_len, max_iter, obstacles3, n_obs, draws=None, seed=None):
    """In-place smoothing of Q device paths.  path [Q, cap, 2] float64 CUDA, path_len [Q] int32 CUDA,
    obstacles3 [Q, O, 3] (x, y, size) CUDA, n_obs [Q] int32 CUDA, draws [Q, max_iter, 2] unit uniforms (drawn with
    torch if None).  Returns (status [Q], iters_done [Q]) device tensors."""
    torch = _lib.require_cuda()
    q, cap = path.shape[0], path.shape[1]
    dev = path.device
    if draws is None:
        gen = torch.Generator(device=dev)
        if seed is not None:
            gen.manual_seed(int(seed))
        draws = torch.rand((q, max_iter, 2), dtype=torch.float64, device=dev, generator=gen)
    for t, dt in ((path, torch.float64), (path_len, torch.int32), (draws, torch.float64), (obstacles3, torch.float64),
                  (n_obs, torch.int32)):
        if t.dtype != dt or not t.is_contiguous() or not t.is_cuda:
            raise _lib.RrtkError("smooth_batch: tensors must be contiguous CUDA tensors of the documented dtype")
    status = torch.empty((q,), dtype=torch.int32, device=dev)
    iters = torch.empty((q,), dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().rrtk_path_smoothing_dev(
        q, cap, int(max_iter), path.data_ptr(), path_len.data_ptr(), draws.data_ptr(), obstacles3.data_ptr(),
        obstacles3.shape[1], n_obs.data_ptr(), status.data_ptr(), iters.data_ptr(),
        torch.cuda.current_stream().cuda_stream), "rrtk_path_smoothing_dev")
    return status, iters


def path_smoothing(path, max_iter, obstacle_list, draws=None):
    """The reference's `path_smoothing` (rrt_04:1447-1479): returns the smoothed path as a list of [x, y].
    `draws` ([max_iter, 2] unit uniforms) stands in for the reference's random.uniform calls; by default they are
    drawn from Python's `random`, two per iteration, in the reference's order."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda")
    if draws is None:
        draws = [[random.random(), random.random()] for _ in range(max_iter)]
    draws = np.ascontiguousarray(np.asarray(draws, dtype=np.float64).reshape(1, -1, 2))[:, :max_iter]
    pts = np.asarray(path, dtype=np.float64).reshape(-1, 2)
    cap = pts.shape[0] + max_iter + 2
    buf = np.zeros((1, cap, 2))
    buf[0, :pts.shape[0]] = pts
    obs = np.asarray([list(o) for o in obstacle_list], dtype=np.float64).reshape(1, -1, 3)
    if obs.shape[1] == 0:
        obs = np.zeros((1, 1, 3))
    d_path = torch.from_numpy(buf).to(dev)
    d_len = torch.tensor([pts.shape[0]], dtype=torch.int32, device=dev)
    status, _ = smooth_batch(d_path, d_len, max_iter, torch.from_numpy(np.ascontiguousarray(obs)).to(dev),
                             torch.tensor([len(obstacle_list)], dtype=torch.int32, device=dev),
                             torch.from_numpy(draws).to(dev))
    st = int(status[0].item())
    if st & _lib.Q_DIV_ZERO:
        raise ZeroDivisionError("float division by zero")   # what the reference raises here
    if st & _lib.Q_PATH_OVERFLOW:
        raise _lib.RrtkError("path longer than 512 points")
    n = int(d_len[0].item())
    return d_path[0, :n].cpu().numpy().tolist()
