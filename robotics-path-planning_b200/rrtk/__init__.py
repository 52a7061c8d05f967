"""rrtk -- B200-native RRT-family planning kernels behind the reference's Python class API."""
from ._lib import RrtkError, lib, LIB_PATH  # noqa: F401
from .planners import RRT, RRTStar, Node, AreaBounds  # noqa: F401
from .batch import RRTStarBatch, shard_range  # noqa: F401
from .informed import InformedRRTStar  # noqa: F401
from . import dubins  # noqa: F401
from .dubins import plan_dubins_path  # noqa: F401
from .dubins_planner import RRTStarDubins, RRTDubins  # noqa: F401
from .rs_planner import RRTStarReedsShepp  # noqa: F401
from .arm import NLinkArm, get_occupancy_grid, get_occupancy_grids, astar_torus, astar_torus_batch  # noqa: F401
from . import smoothing, reeds_shepp  # noqa: F401
from .reeds_shepp import reeds_shepp_path_planning  # noqa: F401
from .smoothing import path_smoothing, get_path_length  # noqa: F401
from . import closed_loop, bitstar  # noqa: F401
from .bitstar import BITStar  # noqa: F401
from .closed_loop import ClosedLoopRRTStar  # noqa: F401
from .large_tree import LargeTree, RRTLarge  # noqa: F401

__all__ = ["LargeTree", "RRTLarge", "RRT", "RRTStar", "Node", "AreaBounds", "InformedRRTStar", "RRTStarDubins", "RRTDubins", "RRTStarReedsShepp", "ClosedLoopRRTStar", "closed_loop", "BITStar", "bitstar", "plan_dubins_path", "dubins", "RRTStarBatch", "shard_range", "NLinkArm", "get_occupancy_grid", "get_occupancy_grids", "astar_torus", "astar_torus_batch", "path_smoothing", "get_path_length", "smoothing", "reeds_shepp", "reeds_shepp_path_planning", "RrtkError",
           "lib", "LIB_PATH"]
