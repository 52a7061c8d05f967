"""ctypes binding of librrtk.so (include/rrtk.h).  There is no CPU fallback: if the CUDA extension
is missing or no GPU is present, every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RRTK_LIB", os.path.join(HERE, "librrtk.so"))  # override: tuning builds only

SAMPLER_STREAM, SAMPLER_SOBOL, SAMPLER_UNIFORM = 0, 1, 2
EXEC_AUTO, EXEC_WARP, EXEC_CTA = 0, 1, 2   # rrtk_rrtstar_params.exec_mode / rrtk_dubins_params.exec_mode
WS_TAIL_INTS = 4                           # RRTK_WS_TAIL_INTS: every planner workspace ends with the work-queue counter
Q_OK, Q_NEAR_OVERFLOW, Q_NODE_OVERFLOW, Q_PATH_OVERFLOW, Q_DIV_ZERO, Q_NONE_STEER = 0, 1, 2, 4, 8, 16


class RrtkError(RuntimeError):
    pass


class RRTStarParams(C.Structure):
    """struct rrtk_rrtstar_params (include/rrtk.h)."""
    _fields_ = [("n_queries", C.c_int32), ("max_iter", C.c_int32), ("node_cap", C.c_int32),
                ("obs_stride", C.c_int32), ("near_cap", C.c_int32),
                ("search_until_max_iter", C.c_int32), ("sampler", C.c_int32),
                ("goal_sample_rate", C.c_int32), ("has_play_area", C.c_int32),
                ("rrt_only", C.c_int32), ("expand_dis", C.c_double),
                ("path_resolution", C.c_double), ("min_rand", C.c_double),
                ("max_rand", C.c_double), ("play_area", C.c_double * 4), ("seed", C.c_uint64),
                ("grid_nx", C.c_int32), ("grid_ny", C.c_int32), ("grid_x0", C.c_double),
                ("grid_y0", C.c_double), ("grid_cell", C.c_double), ("resume", C.c_int32), ("iter_offset", C.c_int32),
                ("near_r_max", C.c_double), ("exec_mode", C.c_int32), ("query_base", C.c_int32)]


class InformedParams(C.Structure):
    """struct rrtk_informed_params (include/rrtk.h)."""
    _fields_ = [("n_queries", C.c_int32), ("max_iter", C.c_int32), ("node_cap", C.c_int32),
                ("obs_stride", C.c_int32), ("path_cap", C.c_int32), ("exec_mode", C.c_int32),
                ("expand_dis", C.c_double), ("coord_bound", C.c_double)]


class InformedTreeParams(C.Structure):
    """struct rrtk_informed_tree_params (include/rrtk.h)."""
    _fields_ = [("max_iter", C.c_int32), ("node_cap", C.c_int32), ("n_obs", C.c_int32), ("path_cap", C.c_int32),
                ("grid", C.c_int32), ("batch", C.c_int32), ("expand_dis", C.c_double),
                ("start_goal", C.c_double * 4), ("rot", C.c_double * 4), ("coord_bound", C.c_double)]


class InformedTreeResult(C.Structure):
    """struct rrtk_informed_tree_result (include/rrtk.h)."""
    _fields_ = [("n_nodes", C.c_int32), ("path_len", C.c_int32), ("status", C.c_int32), ("iters_done", C.c_int32),
                ("c_best", C.c_double), ("total_hits", C.c_int64), ("slow_paths", C.c_int32),
                ("goal_events", C.c_int32), ("resamples", C.c_int32), ("grid", C.c_int32),
                ("reextends", C.c_int32), ("pad_", C.c_int32), ("cycles", C.c_int64 * 6),
                ("cycles_max", C.c_int64 * 6), ("cycles_negmin", C.c_int64 * 6)]


class DubinsParams(C.Structure):
    """struct rrtk_dubins_params (include/rrtk.h)."""
    _fields_ = [("n_queries", C.c_int32), ("max_iter", C.c_int32), ("node_cap", C.c_int32),
                ("obs_stride", C.c_int32), ("near_cap", C.c_int32), ("search_until_max_iter", C.c_int32),
                ("curvature", C.c_double), ("step_size", C.c_double), ("goal_xy_th", C.c_double),
                ("goal_yaw_th", C.c_double), ("rs_cost", C.c_int32), ("exec_mode", C.c_int32)]


class ClosedLoopParams(C.Structure):
    """struct rrtk_closed_loop_params (include/rrtk.h)."""
    _fields_ = [("n_courses", C.c_int32), ("course_cap", C.c_int32), ("traj_cap", C.c_int32), ("pad_", C.c_int32),
                ("target_speed", C.c_double), ("yaw_th", C.c_double), ("invalid_travel_ratio", C.c_double)]


class BitStarParams(C.Structure):
    """struct rrtk_bitstar_params (include/rrtk.h)."""
    _fields_ = [(k, C.c_int32) for k in ("n_queries", "max_iter", "vertex_cap", "sample_cap", "edge_cap", "path_cap",
                                         "obs_stride", "n_draws")] + \
               [(k, C.c_double) for k in ("min_rand", "max_rand", "num_cells")]


BIT_SAMPLE_OVERFLOW, BIT_EDGE_OVERFLOW, BIT_VERTEX_OVERFLOW, BIT_DRAWS_EXHAUSTED, BIT_INDEX_ERROR, BIT_PATH_OVERFLOW, \
    BIT_LIVELOCK = 1, 2, 4, 8, 16, 32, 64
CL_NOT_REACHED, CL_BAD_ANGLE, CL_TOO_LONG, CL_COLLISION, CL_TRAJ_OVERFLOW = 1, 2, 4, 8, 16

_lib = None

_VP = C.c_void_p
_SIGS = {
    "rrtk_version": (C.c_int, []),
    "rrtk_last_error": (C.c_char_p, []),
    "rrtk_device_count": (C.c_int, []),
    "rrtk_sizeof": (C.c_int, [C.c_int]),
    "rrtk_sobol_fill_dev": (C.c_int, [C.c_int, C.c_int64, C.c_int64, _VP, _VP]),
    "rrtk_sobol_fill_host": (C.c_int, [C.c_int, C.c_int64, C.c_int64, _VP]),
    "rrtk_sobol_table": (C.c_int, [C.c_int, _VP]),
    "rrtk_rrtstar_run_dev": (C.c_int, [C.POINTER(RRTStarParams)] + [_VP] * 16),
    "rrtk_rrtstar_run_host": (C.c_int, [C.POINTER(RRTStarParams)] + [_VP] * 14),
    "rrtk_informed_run_dev": (C.c_int, [C.POINTER(InformedParams)] + [_VP] * 18),
    "rrtk_informed_tree_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "rrtk_informed_tree_run_dev": (C.c_int, [C.POINTER(InformedTreeParams)] + [_VP] * 10 + [C.c_int64, _VP]),
    "rrtk_tree_exchange_probe_dev": (C.c_int, [C.c_int32, C.c_int32, _VP, _VP, C.c_int64, _VP]),
    "rrtk_rrtstar_dubins_run_dev": (C.c_int, [C.POINTER(DubinsParams)] + [_VP] * 17),
    "rrtk_rrtstar_rs_run_dev": (C.c_int, [C.POINTER(DubinsParams)] + [_VP] * 17),
    "rrtk_rrt_dubins_run_dev": (C.c_int, [C.POINTER(DubinsParams)] + [_VP] * 17),
    "rrtk_dubins_steer_dev": (C.c_int, [C.c_int32, C.c_double, C.c_double, _VP, _VP, _VP, _VP, C.c_int32, _VP,
                                        _VP, _VP, _VP, _VP, _VP, _VP, C.c_int32, _VP]),
    "rrtk_reeds_shepp_steer_dev": (C.c_int, [C.c_int32, C.c_double, C.c_double, _VP, _VP, _VP, _VP, C.c_int32, _VP,
                                             _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_int32, _VP]),
    "rrtk_extract_paths_dev": (C.c_int, [C.c_int32, C.c_int32, C.c_int32] + [_VP] * 7),
    "rrtk_path_smoothing_dev": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _VP, _VP, _VP, _VP, C.c_int32, _VP, _VP, _VP, _VP]),
    "rrtk_closed_loop_dev": (C.c_int, [C.POINTER(ClosedLoopParams), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "rrtk_bitstar_run_dev": (C.c_int, [C.POINTER(BitStarParams)] + [_VP] * 12),
    "rrtk_steer_collide_dev": (C.c_int, [C.c_int64, _VP, _VP, C.c_double, C.c_double, _VP, _VP, C.c_int32, _VP, _VP, _VP, _VP, _VP,
                                         _VP, _VP, _VP]),
    "rrtk_steer_points_dev": (C.c_int, [C.c_int64, _VP, _VP, _VP, C.c_double, C.c_double, C.c_int32, _VP, _VP, _VP]),
    "rrtk_points_collide_dev": (C.c_int, [C.c_int32, _VP, _VP, C.c_int32, _VP, _VP, C.c_int32, _VP, _VP, _VP]),
    "rrtk_nearest_f64_dev": (C.c_int, [_VP, C.c_int64, _VP, C.c_int32, _VP, _VP, _VP]),
    "rrtk_near_f64_dev": (C.c_int, [_VP, C.c_int32, C.c_double, C.c_double, C.c_double, _VP, _VP, C.c_int32, _VP, _VP]),
    "rrtk_sample_stream_dev": (C.c_int, [C.POINTER(RRTStarParams), _VP, _VP, _VP, _VP]),
    "rrtk_crmath_probe_dev": (C.c_int, [C.c_int, C.c_int64, _VP, _VP, _VP, _VP]),
    "rrtk_nearest_f32_dev": (C.c_int, [_VP, C.c_int64, _VP, C.c_int32, _VP, _VP, _VP, _VP]),
    "rrtk_near_f32_dev": (C.c_int, [_VP, C.c_int64, C.c_float, C.c_float, C.c_float, _VP, C.c_int32, _VP, _VP]),
    "rrtk_astar_torus_dev": (C.c_int, [C.c_int32, C.c_int32, _VP, _VP, _VP, C.c_int32, _VP, _VP, _VP, _VP, _VP, _VP]),
    "rrtk_fma_peak_dev": (C.c_int, [C.c_int, C.c_int32, C.c_int32, _VP, _VP]),
    "rrtk_arm_grid_dev": (C.c_int, [C.c_int32, _VP, C.c_int32, C.c_int32, C.c_int32, _VP, _VP, C.c_int32,
                                    C.c_int32, _VP, _VP]),
    "rrtk_arm_grid_cells_dev": (C.c_int, [C.c_int32, _VP, C.c_int32, C.c_int32, C.c_int32, _VP, _VP, C.c_int32,
                                    C.c_int32, _VP, _VP]),
}

EXPORTED = tuple(_SIGS)


def lib():
    """Load librrtk.so (built by `__graft_entry__.build()` / csrc/Makefile).  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RrtkError(f"CUDA extension not built: {LIB_PATH} is missing "
                            "(run `python -c 'import __graft_entry__ as g; g.build()'`); "
                            "rrtk has no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            if "RRTK_LIB" in os.environ and not hasattr(handle, name):
                continue                      # an older tuning build selected through RRTK_LIB (A/B timing only)
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = "rrtk") -> None:
    if rc != 0:
        msg = lib().rrtk_last_error().decode("utf-8", "replace")
        raise RrtkError(f"{what} failed (status {rc}): {msg}")


def require_cuda():
    """torch + a visible GPU, or raise (no CPU fallback)."""
    import torch
    if not torch.cuda.is_available():
        raise RrtkError("rrtk needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
    return torch
