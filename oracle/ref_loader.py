"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference scripts.

The reference (`/root/reference/src_path_planning/*.py`) is a set of notebook-style
scripts that run a matplotlib demo at import time and matplotlib is not installed.
This loader registers stub ``matplotlib`` modules, reads a reference file and
``exec``s only its definition block, returning the namespace.

It is used only by ``oracle/make_golden.py`` (which runs in the build container,
where ``/root/reference`` exists) to generate the fixtures under ``tests/golden/``.
Nothing in the product, in ``-m gpu`` tests, in ``smoke()`` or in ``bench.py`` imports
this module: ``/root/reference`` does not exist on the GPU box.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

REF_DIR = os.environ.get("RRTK_REFERENCE_DIR", "/root/reference/src_path_planning")

# file alias -> (file name, last line of the definition block)   [SURVEY.md 8c]
FILES = {
    "rrt_01": ("10_path_planning_01_rrt_01_simple.py", 337),
    "rrt_03": ("10_path_planning_01_rrt_03_dubins_path.py", 1650),
    "rrt_04": ("10_path_planning_01_rrt_04_rrt_star.py", 1482),
    "rrt_05": ("10_path_planning_01_rrt_05_rrt_star_dubins_path.py", 1797),
    "rrt_07": ("10_path_planning_01_rrt_07_informed_rrt_star.py", 1330),
    "arm02": ("02_arm_obstacle_navigation.py", 283),
    "dub00": ("10_path_planning_00_dubins_path.py", 423),
    "rrt_06": ("10_path_planning_01_rrt_06_rrt_star_reeds_shepp_path.py", 2005),
    "rs00": ("10_path_planning_00_reeds_shepp_path.py", 515),
    "rrt_08": ("10_path_planning_01_rrt_08_batch_informed_rrt_star.py", 611),
    "rrt_10": ("10_path_planning_01_rrt_10_closed_loop_rrt_star.py", 1608),   # classes + the constants :1592-1607
}


def available() -> bool:
    return os.path.isdir(REF_DIR)


def _install_matplotlib_stub() -> None:
    if "matplotlib" in sys.modules:
        return

    class _Anything(types.ModuleType):
        def __getattr__(self, name):  # any attribute is a no-op callable
            def _noop(*a, **k):
                return None
            return _noop

    mpl = _Anything("matplotlib")
    plt = _Anything("matplotlib.pyplot")
    colors = _Anything("matplotlib.colors")
    mpl.pyplot = plt
    mpl.colors = colors
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    sys.modules["matplotlib.colors"] = colors


def load(alias: str) -> dict:
    """Exec the definition block of reference file `alias`; return its globals."""
    fname, last = FILES[alias]
    _install_matplotlib_stub()
    path = os.path.join(REF_DIR, fname)
    with open(path, "r") as fh:
        lines = fh.readlines()
    if alias == "rrt_03":
        # rrt_03 builds _PATH_TYPE_MAP (:1030-1031) BEFORE the word functions it names are defined (:1138-1212), so the
        # file raises NameError as shipped; the statement is executed after the definitions instead
        assert lines[1029].startswith("_PATH_TYPE_MAP"), "rrt_03 layout changed"
        moved = lines[1029:1031]
        lines[1029:1031] = ["\n", "\n"]
        lines = lines[:last] + moved
        last += 2
    src = "".join(lines[:last])
    ns: dict = {"__name__": "ref_" + alias, "__file__": path}
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(src, path, "exec"), ns)
    if alias == "rrt_10":
        ns["animation"] = False        # module global read by closed_loop_prediction (:1350), set at :1637 in the script
    return ns


def reset_sobol(ns: dict) -> None:
    """Reset the module-global Sobol state (rrt_04:41-51)."""
    for k in ("initialized", "seed_save", "dim_num_save"):
        if k in ns:
            ns[k] = None
    if "initialized" in ns:
        ns["initialized"] = 0


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield
